"""Times (CUDA events) one decoder step (CUDA-graph replay) and one encoder pass at the bench shapes;
with --eager N runs N eager decoder steps so that `ncu --metrics gpu__time_duration.sum` lists every launch.

    python tools/profile_step.py [--model large-v3] [--windows 120] [--enc-windows 40] [--eager 0] [--skip-encoder]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="large-v3")
    ap.add_argument("--windows", type=int, default=120)
    ap.add_argument("--enc-windows", type=int, default=40)
    ap.add_argument("--eager", type=int, default=0)
    ap.add_argument("--skip-encoder", action="store_true")
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--trace", action="store_true", help="time every 16-step block up to 220 steps, then profile one eager step there")
    a = ap.parse_args()

    from bench import build_model
    from whisper_mlx_b200.decoding import DecodingOptions, DecodingTask, DecodeSession

    model, _ = build_model(a.model, 0, "cuda:0")
    dm = model.dims
    out = {}

    def ev():
        return torch.cuda.Event(enable_timing=True)

    if not a.skip_encoder:
        slabs = (torch.randn(a.enc_windows, 3002, dm.n_mels, device="cuda") * 0.3).bfloat16()
        slabs[:, 0] = 0
        slabs[:, -1] = 0
        model.encode_slabs(slabs)
        torch.cuda.synchronize()
        e0, e1 = ev(), ev()
        e0.record()
        for _ in range(3):
            xa = model.encode_slabs(slabs)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        out["encoder_ms_per_window"] = ms / a.enc_windows
        flops = a.enc_windows * 2273.8e9 if a.model.startswith("large") else None
        out["encoder_tflops"] = flops / (ms * 1e-3) / 1e12 if flops else None
        from whisper_mlx_b200._lib import kernel_profile

        with kernel_profile() as prof:
            model.encode_slabs(slabs)
        out["encoder_kernels_ms_per_window"] = {k: {"launches": v["launches"], "ms": v["total_ms"] / a.enc_windows,
                                                    "avg_us": v["total_ms"] / v["launches"] * 1e3} for k, v in prof.result.items()}
        e0, e1 = ev(), ev()
        e0.record()
        ckv = model.cross_kv(xa)
        e1.record()
        torch.cuda.synchronize()
        out["cross_kv_ms_per_window"] = e0.elapsed_time(e1) / a.enc_windows
        del ckv, slabs

    B = a.windows
    xa = torch.randn(B, dm.n_audio_ctx, dm.n_audio_state, device="cuda").bfloat16()
    task = DecodingTask(model, DecodingOptions(language="en"))
    sess = DecodeSession(model, xa, 1, max_tokens=3 + 224)
    init = torch.tensor(task.initial_tokens, dtype=torch.int32).repeat(B, 1)
    sess.set_tokens(init)
    sess.set_filter(task._filter_params(sess), task._get_suppress_tokens())
    e0, e1 = ev(), ev()
    e0.record()
    sess.prompt_step(len(task.initial_tokens), task.sot_index)
    e1.record()
    torch.cuda.synchronize()
    out["prompt_step_ms"] = e0.elapsed_time(e1)
    if a.eager:
        from whisper_mlx_b200._lib import kernel_profile

        sess._step(1, -1, True)
        with kernel_profile() as prof:
            for _ in range(a.eager):
                sess._step(1, -1, True)
        out["eager_kernels_ms_per_step"] = {k: {"launches": v["launches"] // a.eager, "ms": v["total_ms"] / a.eager,
                                                "avg_us": v["total_ms"] / v["launches"] * 1e3} for k, v in prof.result.items()}
        out["eager_sum_ms_per_step"] = sum(v["total_ms"] for v in prof.result.values()) / a.eager
    else:
        sess.sample_step()
        torch.cuda.synchronize()
        e0, e1 = ev(), ev()
        e0.record()
        for _ in range(a.steps):
            sess.sample_step()
        e1.record()
        torch.cuda.synchronize()
        out["decode_step_ms"] = e0.elapsed_time(e1) / a.steps
        out["graph_kernels_per_step"] = sess._graph_kernels
        if a.trace:
            from whisper_mlx_b200._lib import kernel_profile

            done = 1 + a.steps
            evs = [ev()]
            evs[0].record()
            marks = [done]
            while done + 16 <= 220:
                for _ in range(16):
                    sess.sample_step()
                done += 16
                evs.append(ev())
                evs[-1].record()
                marks.append(done)
            torch.cuda.synchronize()
            out["step_ms_by_position"] = {str(marks[i + 1]): evs[i].elapsed_time(evs[i + 1]) / 16 for i in range(len(evs) - 1)}
            with kernel_profile() as prof:
                sess._step(1, -1, True)
            out["eager_kernels_at_end"] = {k: {"launches": v["launches"], "ms": v["total_ms"],
                                               "avg_us": v["total_ms"] / v["launches"] * 1e3} for k, v in prof.result.items()}
    print(json.dumps(out))
    os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
    with open(os.path.join(REPO, "gpurun_out", "profile_step.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
