"""A/B of the 1-CTA and 2-CTA GEMM kernels on the encoder shapes (run with B200W_GEMM2=0 / 1)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from whisper_mlx_b200 import _lib as L
from tools.perf_probe import timed
lib = L.load()
print("GEMM2 =", os.environ.get("B200W_GEMM2", "1"))
for M in (12000, 24000, 60000, 180000):
    for (N, K, flags, resid) in ((3840, 1280, 0, 0), (1280, 1280, 2, 1), (5120, 1280, 1, 0), (1280, 5120, 2, 1), (2560, 1280, 0, 0)):
        a = torch.randn(M, K, device="cuda").bfloat16()
        w = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
        b = torch.randn(N, device="cuda")
        c = torch.randn(M, N, device="cuda") if flags & 2 else torch.empty((M, N), dtype=torch.bfloat16, device="cuda")
        r = c if resid else None
        med, _ = timed(lambda: L.check(lib.b200w_gemm_bf16(L.ptr(a), K, L.ptr(w), L.ptr(c), N, L.ptr(b), L.ptr(r), M, N, K, flags, L.stream())))
        print(f"M={M:6d} N={N} K={K} flags={flags} resid={resid}: {med:.3f} ms {2.0*M*N*K/med/1e9:.0f} TF/s", flush=True)
        del a, w, c
