#!/usr/bin/env python
"""Benchmark of the Whisper transcription hot path on B200 (contract: see DESIGN.md, "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--model large-v3] [--hours 1.0]

Workload (BASELINE.json config 4): whisper-large-v3 (128 mel, 32+32 layers, random-init, bf16) greedy
transcription of 1 h of synthetic 16 kHz audio = 120 fixed 30 s windows, through the public
`transcribe(audio, ...)` API in fixed-window batched mode.  One step = one pass over the whole hour.
  value : RTFx (audio seconds per wall second) with the audio already resident in HBM.
  e2e   : the same call with the audio in pinned HOST memory: H2D copy of the samples and D2H read of the
          decoded tokens are inside the timed region.
Multi-GPU: one process per GPU, each rank transcribes its own hour (windows are independent units; no
collective on the data path), whole-job RTFx = N * 3600 / max-over-ranks time -> "scaling": "weak".
`--impl reference` times the CPU restatement of the reference algorithm (oracle/, the only other place
that may execute it) on the host cores for a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

METRIC = "large-v3 RTFx (audio-s/s)"
UNIT = "audio-s/s"


def load_peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained"),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except FileNotFoundError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ======================================================================================== product arm
def build_model(name: str, seed: int, device: str):
    import torch
    from tools import synth
    from whisper_mlx_b200.whisper import ModelDimensions, Whisper

    dims = synth.DIMS[name]
    weights = dict(synth.random_weights(dims, seed, device=device))
    model = Whisper(ModelDimensions(**dims), weights, device=device)
    return model, weights


def make_audio(hours: float, seed: int) -> np.ndarray:
    from tools import synth

    return synth.long_audio(hours * 3600.0, seed)


def kernel_rooflines(model, peaks, n_windows: int):
    """Per-kernel achieved vs. roofline at the workload's shapes, timed with CUDA events on the launch stream."""
    import torch
    from whisper_mlx_b200 import _lib as L

    lib = L.load()
    dm = model.dims
    d, T, H = dm.n_text_state, dm.n_audio_ctx, dm.n_text_head
    out = {}

    def timed(fn, n):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(n):
            fn(i)
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n * 1e-3

    # K8 cross-attention: one launch per decoder layer per step; each launch streams its layer's K|V once
    Lr = min(dm.n_text_layer, 8)
    ckv = torch.empty((Lr, n_windows, T, 2 * d), dtype=torch.bfloat16, device=model.device).normal_()
    q = torch.randn(n_windows, 1, d, device=model.device).bfloat16()
    o = torch.empty_like(q)
    slot = torch.arange(n_windows, dtype=torch.int32, device=model.device)
    t = timed(lambda i=0: L.check(lib.b200w_decoder_cross_attention(L.ptr(q), n_windows, 1, H, L.ptr(ckv[i % Lr]), T * 2 * d, T,
                                                                     L.ptr(slot), L.ptr(o), L.stream())), 4 * Lr)
    bytes_alg = n_windows * T * 2 * d * 2  # K and V rows of every window, bf16, read once
    # DRAM traffic per launch from the ncu --set full capture of this kernel at 120 windows
    # (profiles/r01_ncu_full_summaries.json: 921.93 MB read + 4.60 MB written), scaled to this launch's windows
    traffic = (921_940_480 + 4_274_944) * n_windows / 120.0 if d == 1280 else None  # profiles/r01_ncu_full_summaries_v2.json
    out["cross_attention_decode"] = {"bound": "hbm", "achieved": bytes_alg / t / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                     "frac": bytes_alg / t / 1e9 / peaks["hbm_gbs"], "traffic": traffic, "launch_ms": t * 1e3,
                                     "algorithmic_bytes_per_launch": bytes_alg}
    del ckv
    # K5 encoder GEMM (fused QKV projection shape): M = windows * 1500, N = 3d, K = d
    M = min(n_windows, 32) * T
    a = torch.randn(M, d, device=model.device).bfloat16()
    w = (torch.randn(3 * d, d, device=model.device) / d ** 0.5).bfloat16()
    bias = torch.zeros(3 * d, device=model.device)
    c = torch.empty((M, 3 * d), dtype=torch.bfloat16, device=model.device)
    t = timed(lambda i=0: L.check(lib.b200w_gemm_bf16(L.ptr(a), d, L.ptr(w), L.ptr(c), 3 * d, L.ptr(bias), None, M, 3 * d, d, 0,
                                                       L.stream())), 10)
    fl = 2.0 * M * 3 * d * d
    out["encoder_gemm_qkv"] = {"bound": "tensor", "achieved": fl / t / 1e12, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                               "frac": fl / t / 1e12 / peaks["bf16_tflops"], "traffic": None, "launch_ms": t * 1e3,
                               "algorithmic_flops_per_launch": fl}
    del a, w, c
    # K1 log-mel at BASELINE config 2 (batch 1024 x 30 s), f32 in / f32 out incl. the clamp pass
    from whisper_mlx_b200.audio import log_mel_spectrogram

    x = torch.randn(1024, 480000, device=model.device) * 0.1
    for n_mels in (80, 128):
        t = timed(lambda i=0: log_mel_spectrogram(x, n_mels=n_mels), 5)
        bytes_alg = 1024 * (4 * 480000 + 4 * 3000 * n_mels)
        out[f"logmel_{n_mels}"] = {"bound": "hbm", "achieved": bytes_alg / t / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                   "frac": bytes_alg / t / 1e9 / peaks["hbm_gbs"], "traffic": None, "launch_ms": t * 1e3,
                                   "frames_per_s": 1024 * 3000 / t, "algorithmic_bytes_per_launch": bytes_alg}
    return out


def run_product(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (B200); there is no CPU fallback for the product arm")
    torch.cuda.set_device(local)
    device = f"cuda:{local}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(device))

    from whisper_mlx_b200 import transcribe
    from whisper_mlx_b200.decoding import total_kernel_launches

    peaks = load_peaks()
    model, weights = build_model(args.model, 0, device)
    audio_host = torch.from_numpy(make_audio(args.hours, 100 + rank)).pin_memory()
    audio_dev = audio_host.to(device)
    n_windows = int(np.ceil(args.hours * 3600 / 30))
    kw = dict(model=model, temperature=0.0, condition_on_previous_text=False, language="en",
              window_batch=args.window_batch, encoder_batch=args.encoder_batch)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run(audio, steps):
        res = None
        barrier()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            res = transcribe(audio, **kw)
        e1.record()
        torch.cuda.synchronize()
        dev_s = e0.elapsed_time(e1) * 1e-3
        barrier()
        wall = time.perf_counter() - t0
        t = torch.tensor([dev_s, wall], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t[0].item(), t[1].item(), res

    for _ in range(args.warmup):
        transcribe(audio_dev, **kw)
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = total_kernel_launches()
    dev_s, wall_s, res = run(audio_dev, args.steps)
    launches = total_kernel_launches() - launches0
    e2e_dev_s, e2e_wall_s, res_h = run(audio_host, args.steps)
    clocks = sampler.stop()
    audio_s = args.hours * 3600.0 * args.steps * world
    n_tokens = sum(len(s["tokens"]) for s in res["segments"])

    line = None
    if rank == 0:
        roof = kernel_rooflines(model, peaks, min(n_windows, args.window_batch))
        # the CPU baseline is timed on rank 0 at N = 1 only (torchrun pins OMP_NUM_THREADS=1 for N > 1)
        cpu = cpu_baseline(args, weights) if (not args.no_cpu_baseline and world == 1) else None
        dom = roof["cross_attention_decode"]
        line = {
            "metric": METRIC, "value": audio_s / dev_s, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_s / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"whisper-{args.model} (random-init) greedy transcription of {args.hours:g} h synthetic 16 kHz audio per GPU "
                                   f"= {n_windows} fixed 30 s windows, transcribe(window_batch={args.window_batch}), sample_len 224, no fallback",
                       "windows_per_gpu": n_windows, "window_batch": args.window_batch, "encoder_batch": args.encoder_batch,
                       "tokens_decoded_per_step": n_tokens, "l2": "working set (weights 3.1 GB + cross-KV 29.5 GB) >> 126 MB L2",
                       "parallelism": f"windows sharded, {world} replica(s), no collective"},
            "e2e": {"value": audio_s / e2e_dev_s, "unit": UNIT, "h2d_bytes_per_step": int(audio_host.numel() * 4),
                    "d2h_bytes_per_step": int(n_windows * (456 * 4 + 12)), "wall_value": audio_s / e2e_wall_s},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {k: dom[k] for k in ("bound", "achieved", "peak", "unit", "frac", "traffic")},
            "roofline_kernel": "decoder_cross_attention_kernel (K8)",
            "roofline_peak_source": peaks["source"],
            "rooflines": roof,
            "cpu_baseline": cpu,
            "wall_value": audio_s / wall_s,
        }
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line), flush=True)


# ======================================================================================== CPU arm
def _cpu_sample(weights_f32, dims_dict, sample_len: int, threads: int):
    """One 30 s window through the oracle: log-mel + encoder + `sample_len` greedy decode steps (fp32)."""
    import torch
    from oracle import audio as OA, decoding as OD, model as OM
    from tools import synth

    torch.set_num_threads(threads)
    dims = OM.ModelDimensions(**dims_dict)
    x = synth.white_noise(480000, 7)
    t0 = time.perf_counter()
    mel = torch.from_numpy(OA.log_mel_spectrogram(x, dims.n_mels))[None]
    t1 = time.perf_counter()
    xa = OM.encoder_forward(weights_f32, dims, mel)
    t2 = time.perf_counter()
    OD.decode(weights_f32, dims, mel, language="en", sample_len=sample_len, audio_features=xa)
    t3 = time.perf_counter()
    return {"logmel_s": t1 - t0, "encoder_s": t2 - t1, "decode_s": t3 - t2, "decode_steps": sample_len}


def cpu_baseline(args, weights=None, steps: int = 1):
    import torch
    from tools import synth

    threads = os.cpu_count() or 1
    dims = synth.DIMS[args.model]
    if weights is None:
        weights = dict(synth.random_weights(dims, 0, device="cpu"))
    w32 = {k: v.detach().to("cpu", torch.float32) for k, v in weights.items()}
    sample_len = args.cpu_sample_len
    best = None
    for _ in range(steps):
        r = _cpu_sample(w32, dims, sample_len, threads)
        if best is None or sum(r[k] for k in ("logmel_s", "encoder_s", "decode_s")) < sum(best[k] for k in ("logmel_s", "encoder_s", "decode_s")):
            best = r
    # scale the measured decode steps to the 224 the GPU arm runs per window
    per_window = best["logmel_s"] + best["encoder_s"] + best["decode_s"] * (224.0 / sample_len)
    return {"value": 30.0 / per_window, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"one 30 s window of the same workload through oracle/ (PyTorch-CPU fp32 restatement, not MLX): log-mel "
                      f"{best['logmel_s']:.2f} s + encoder {best['encoder_s']:.2f} s + {sample_len} greedy decode steps {best['decode_s']:.2f} s, "
                      f"decode scaled x{224.0 / sample_len:.1f} to 224 steps",
            "detail": best}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.perf_counter()
    from tools import synth
    import torch

    dims = synth.DIMS[args.model]
    weights = dict(synth.random_weights(dims, 0, device="cpu"))
    w32 = {k: v.to(torch.float32) for k, v in weights.items()}
    threads = os.cpu_count() or 1
    vals = []
    for i in range(args.warmup + args.steps):
        r = _cpu_sample(w32, dims, args.cpu_sample_len, threads)
        per_window = r["logmel_s"] + r["encoder_s"] + r["decode_s"] * (224.0 / args.cpu_sample_len)
        if i >= args.warmup:
            vals.append((per_window, r))
    per_window = float(np.mean([v[0] for v in vals]))
    value = 30.0 / per_window
    n_windows = int(np.ceil(args.hours * 3600 / 30))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_window * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"whisper-{args.model} (random-init) greedy transcription, CPU restatement of the reference algorithm "
                               f"(mlx-whisper is not installable here); each step = one 30 s window sample of the {n_windows}-window job"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"per step: one 30 s window, log-mel + encoder + {args.cpu_sample_len} greedy decode steps scaled to 224; "
                                   f"torch threads = {threads}"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "total_wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default="large-v3", choices=["tiny", "small", "large-v3", "large-v3-turbo", "micro"])
    ap.add_argument("--hours", type=float, default=1.0)
    ap.add_argument("--window-batch", type=int, default=120)
    ap.add_argument("--encoder-batch", type=int, default=40)
    ap.add_argument("--cpu-sample-len", type=int, default=16)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_product(args)


if __name__ == "__main__":
    main()
