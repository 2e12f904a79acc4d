"""Driver for `ncu --set full` on one GEMM launch: python tools/ncu_gemm.py M N K flags(1=gelu,2=f32,4=resid)"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from whisper_mlx_b200 import _lib as L
lib = L.load()
M, N, K, fl = (int(x) for x in sys.argv[1:5])
a = torch.randn(M, K, device="cuda").bfloat16()
w = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
b = torch.randn(N, device="cuda")
c = torch.randn(M, N, device="cuda") if fl & 2 else torch.empty((M, N), dtype=torch.bfloat16, device="cuda")
r = c if fl & 4 else None
for _ in range(3):
    L.check(lib.b200w_gemm_bf16(L.ptr(a), K, L.ptr(w), L.ptr(c), N, L.ptr(b), L.ptr(r), M, N, K, fl & 3, L.stream()))
torch.cuda.synchronize()
print("ok")
