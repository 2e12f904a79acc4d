"""Epilogue cost probe: same GEMM shape with different epilogues."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from whisper_mlx_b200 import _lib as L
from tools.perf_probe import timed
lib = L.load()
for (M, N, K) in [(180000, 1280, 1280), (180000, 1280, 5120), (180000, 5120, 1280)]:
    a = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
    b = torch.randn(N, device="cuda")
    c16 = torch.empty((M, N), dtype=torch.bfloat16, device="cuda")
    c32 = torch.empty((M, N), dtype=torch.float32, device="cuda")
    r32 = torch.randn(M, N, device="cuda")
    for name, c, bias, resid, flags in [("bf16", c16, None, None, 0), ("bf16+bias", c16, b, None, 0), ("bf16+bias+gelu", c16, b, None, 1),
                                        ("f32", c32, None, None, 2), ("f32+bias", c32, b, None, 2), ("f32+bias+resid(inplace)", c32, b, c32, 2),
                                        ("f32+bias+resid(other)", c32, b, r32, 2)]:
        med, best = timed(lambda: L.check(lib.b200w_gemm_bf16(L.ptr(a), K, L.ptr(w), L.ptr(c), N, L.ptr(bias), L.ptr(resid), M, N, K, flags, L.stream())))
        print(f"{M}x{N}x{K} {name:26s} {med:.3f} ms  {2.0*M*N*K/med/1e9:.0f} TF/s", flush=True)
    del a, w, c16, c32, r32
