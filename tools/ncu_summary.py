"""Summarises .ncu-rep captures (ncu --set full) into the JSON committed under profiles/.

    python tools/ncu_summary.py out.json "note" file.ncu-rep=command-description [...]
"""
import csv
import io
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__cycles_elapsed.max"]
STALLS = ["barrier", "long_scoreboard", "short_scoreboard", "wait", "mio_throttle", "math_pipe_throttle", "not_selected",
          "selected", "branch_resolving", "lg_throttle", "membar", "sleeping", "tex_throttle", "dispatch_stall", "no_instruction"]


def summarise(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units, data = rows[0], rows[1], rows[2:]
    out = []
    for r in data:
        d = dict(zip(head, r))
        u = dict(zip(head, units))
        rec = {"Kernel Name": d.get("Kernel Name", "")[:160]}
        for k in KEYS:
            if k in d:
                rec[k] = d[k] + (" " + u[k] if u.get(k) else "")
        for st in STALLS:
            k = f"smsp__average_warps_issue_stalled_{st}_per_issue_active.ratio"
            if k in d:
                rec["stall_" + st] = d[k]
        out.append(rec)
    return out


def main():
    out_path, note = sys.argv[1], sys.argv[2]
    res = {"note": note, "captures": {}}
    for spec in sys.argv[3:]:
        path, _, cmd = spec.partition("=")
        res["captures"][path.split("/")[-1]] = {"command": cmd, "launches": summarise(path)}
    with open(out_path, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res, indent=1)[:3000])


if __name__ == "__main__":
    main()
