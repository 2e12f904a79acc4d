"""Kernel-level timings (CUDA events) for the hot kernels at BASELINE sizes -> gpurun_out/perf_probe.json.

Development aid; bench.py is the judged measurement.  Each probe runs in-process; run after gpu_check.
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from whisper_mlx_b200 import _lib as L  # noqa: E402

lib = L.load()
PEAK_TF = 1645.3
PEAK_GBS = 6540.2


def timed(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts) // 2], ts[0]


def probe_gemm(out):
    for (M, N, K, gelu, f32, resid) in [(180000, 3840, 1280, 0, 0, 0), (180000, 1280, 1280, 0, 1, 1), (180000, 5120, 1280, 1, 0, 0),
                                         (180000, 1280, 5120, 0, 1, 1), (24000, 3840, 1280, 0, 0, 0), (8192, 8192, 8192, 0, 0, 0),
                                         (120, 3840, 1280, 0, 0, 0), (120, 1280, 1280, 0, 1, 1), (120, 5120, 1280, 1, 0, 0),
                                         (120, 1280, 5120, 0, 1, 1), (120, 51866, 1280, 0, 1, 0)]:
        a = torch.randn(M, K, device="cuda").bfloat16()
        w = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
        b = torch.randn(N, device="cuda") if N % 32 == 0 else None
        ld = (N + 127) // 128 * 128
        c = torch.empty((M, ld), dtype=torch.float32 if f32 else torch.bfloat16, device="cuda")
        r = c if resid else None
        flags = gelu | (2 if f32 else 0)

        def run():
            L.check(lib.b200w_gemm_bf16(L.ptr(a), K, L.ptr(w), L.ptr(c), ld, L.ptr(b), L.ptr(r), M, N, K, flags, L.stream()))

        med, best = timed(run)
        tf = 2.0 * M * N * K / (med * 1e-3) / 1e12
        gbs = (N * K * 2 + M * K * 2 + M * N * (4 if f32 else 2)) / (med * 1e-3) / 1e9
        torch_ms, _ = timed(lambda: torch.matmul(a, w.T))
        out[f"gemm_{M}x{N}x{K}_g{gelu}f{f32}r{resid}"] = {"ms": med, "best_ms": best, "tflops": tf, "frac_tensor": tf / PEAK_TF,
                                                         "gbs": gbs, "cublas_ms": torch_ms}
        print(f"gemm {M}x{N}x{K} gelu{gelu} f32{f32} resid{resid}: {med:.3f} ms {tf:.0f} TF/s ({tf / PEAK_TF:.2%}) {gbs:.0f} GB/s | cublas {torch_ms:.3f} ms", flush=True)
        del a, w, c


def probe_attention(out):
    for (B, T, H) in [(16, 1500, 20), (120, 1500, 20)]:
        d = 64 * H
        qkv = torch.randn(B, T, 3 * d, device="cuda").bfloat16()
        o = torch.empty((B, T, d), dtype=torch.bfloat16, device="cuda")
        med, best = timed(lambda: L.check(lib.b200w_encoder_attention(L.ptr(qkv), B, T, H, L.ptr(o), L.stream())))
        fl = 4.0 * B * H * T * T * 64
        out[f"enc_attn_{B}"] = {"ms": med, "tflops": fl / (med * 1e-3) / 1e12}
        print(f"enc attention B={B}: {med:.3f} ms {fl / (med * 1e-3) / 1e12:.0f} TF/s", flush=True)
    T, H, d = 1500, 20, 1280
    for B in (16, 120):
        ckv = torch.randn(B, T, 2 * d, device="cuda").bfloat16()
        q = torch.randn(B, 1, d, device="cuda").bfloat16()
        o = torch.empty((B, 1, d), dtype=torch.bfloat16, device="cuda")
        slot = torch.arange(B, dtype=torch.int32, device="cuda")
        med, best = timed(lambda: L.check(lib.b200w_decoder_cross_attention(L.ptr(q), B, 1, H, L.ptr(ckv), T * 2 * d, T,
                                                                           L.ptr(slot), L.ptr(o), L.stream())), iters=20)
        gbs = ckv.numel() * 2 / (med * 1e-3) / 1e9
        out[f"cross_attn_{B}"] = {"ms": med, "gbs": gbs, "frac_hbm": gbs / PEAK_GBS}
        print(f"cross attention B={B}: {med:.4f} ms {gbs:.0f} GB/s ({gbs / PEAK_GBS:.2%})", flush=True)


def probe_logmel(out):
    from whisper_mlx_b200.audio import log_mel_unclamped, log_mel_spectrogram

    x = (torch.randn(1024, 480000, device="cuda") * 0.1)
    for n_mels in (80, 128):
        med, best = timed(lambda: log_mel_unclamped(x, n_mels), iters=5)
        bytes_alg = 1024 * (4 * 480000 + 4 * 3000 * n_mels)
        out[f"logmel_unclamped_{n_mels}"] = {"ms": med, "gbs": bytes_alg / (med * 1e-3) / 1e9, "gframes_s": 1024 * 3000 / (med * 1e-3) / 1e9}
        med2, _ = timed(lambda: log_mel_spectrogram(x, n_mels), iters=5)
        out[f"logmel_full_{n_mels}"] = {"ms": med2, "gbs": bytes_alg / (med2 * 1e-3) / 1e9, "frac_hbm": bytes_alg / (med2 * 1e-3) / 1e9 / PEAK_GBS}
        print(f"logmel {n_mels}: kernel {med:.3f} ms, with finalize {med2:.3f} ms -> {bytes_alg / (med2 * 1e-3) / 1e9:.0f} GB/s algorithmic", flush=True)


def main():
    out = {}
    which = sys.argv[1:] or ["gemm", "attention", "logmel"]
    for name in which:
        try:
            {"gemm": probe_gemm, "attention": probe_attention, "logmel": probe_logmel}[name](out)
        except Exception as e:  # noqa: BLE001
            out[name + "_error"] = repr(e)
            print("ERROR", name, repr(e), flush=True)
    os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
    with open(os.path.join(REPO, "gpurun_out", "perf_probe.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
