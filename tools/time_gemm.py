"""CUDA-event timing of one GEMM shape with the three epilogues (bf16 / f32 / f32 + residual), L2 flushed between runs."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from whisper_mlx_b200 import _lib as L
lib = L.load()
M, N, K = (int(x) for x in sys.argv[1:4])
a = torch.randn(M, K, device="cuda").bfloat16()
w = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
b = torch.randn(N, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name, fl in (("bf16", 0), ("f32", 2), ("f32+resid", 6), ("f32+resid (other buffer)", 14)):
    c = torch.randn(M, N, device="cuda") if fl & 2 else torch.empty((M, N), dtype=torch.bfloat16, device="cuda")
    r = (torch.randn(M, N, device="cuda") if fl & 8 else c) if fl & 4 else None
    ts = []
    for i in range(8):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.check(lib.b200w_gemm_bf16(L.ptr(a), K, L.ptr(w), L.ptr(c), N, L.ptr(b), L.ptr(r), M, N, K, fl & 3, L.stream()))
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = sorted(ts)[len(ts) // 2]
    print(f"{name:28s} {t*1e3:8.1f} us  {2.0*M*N*K/t/1e9:7.1f} TF/s  frac {2.0*M*N*K/t/1e9/1645.3:.3f}")
