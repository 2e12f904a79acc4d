"""Driver for `ncu --set full` on the decode cross-attention kernel (K8) at the bench shape (120 windows, large-v3)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from whisper_mlx_b200 import _lib as L
lib = L.load()
B, T, H, d = 120, 1500, 20, 1280
ckv = torch.randn(3, B, T, 2 * d, device="cuda").bfloat16()
q = torch.randn(B, 1, d, device="cuda").bfloat16()
o = torch.empty_like(q)
slot = torch.arange(B, dtype=torch.int32, device="cuda")
for i in range(3):
    L.check(lib.b200w_decoder_cross_attention(L.ptr(q), B, 1, H, L.ptr(ckv[i]), T * 2 * d, T, L.ptr(slot), L.ptr(o), L.stream()))
torch.cuda.synchronize()
print("ok")
