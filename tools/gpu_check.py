"""Run every GPU parity case (tests/gpu_cases.py) in its own subprocess with a timeout.

    python tools/gpu_check.py [--timeout 180] [case ...]      -> gpurun_out/gpu_check.json

One faulting kernel (illegal address, mbarrier trap, hang) then costs one case, not the whole run.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time
import traceback

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def run_one(name: str) -> int:
    from tests.gpu_cases import CASES

    try:
        t = time.time()
        res = CASES[name]()
        print("RESULT " + json.dumps({"case": name, "ok": True, "seconds": round(time.time() - t, 2), "metrics": res},
                                     default=str))
        return 0
    except Exception as e:  # noqa: BLE001
        traceback.print_exc()
        print("RESULT " + json.dumps({"case": name, "ok": False, "error": f"{type(e).__name__}: {e}"[:2000]}))
        return 1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--one")
    ap.add_argument("--timeout", type=int, default=240)
    ap.add_argument("--out", default=os.path.join(REPO, "gpurun_out", "gpu_check.json"))
    ap.add_argument("cases", nargs="*")
    a = ap.parse_args()
    if a.one:
        sys.exit(run_one(a.one))
    from tests.gpu_cases import CASES

    names = a.cases or list(CASES)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    report = []
    for n in names:
        t = time.time()
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", n], capture_output=True, text=True,
                               timeout=a.timeout, cwd=REPO)
            line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
            rec = json.loads(line[-1][7:]) if line else {"case": n, "ok": False, "error": "no result line"}
            if not rec.get("ok"):
                rec["stderr_tail"] = p.stderr[-3000:]
                rec["stdout_tail"] = p.stdout[-1500:]
            rec["returncode"] = p.returncode
        except subprocess.TimeoutExpired as e:
            rec = {"case": n, "ok": False, "error": f"timeout after {a.timeout}s",
                   "stderr_tail": (e.stderr or b"")[-2000:].decode(errors="replace") if isinstance(e.stderr, bytes) else str(e.stderr)[-2000:]}
        rec["wall"] = round(time.time() - t, 1)
        report.append(rec)
        print(json.dumps({k: v for k, v in rec.items() if k not in ("stderr_tail", "stdout_tail")}, default=str), flush=True)
        with open(a.out, "w") as f:
            json.dump(report, f, indent=1, default=str)
    n_ok = sum(1 for r in report if r.get("ok"))
    print(f"{n_ok}/{len(report)} cases ok")
    sys.exit(0 if n_ok == len(report) else 1)


if __name__ == "__main__":
    main()
