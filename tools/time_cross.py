"""CUDA-event timing of the decode cross-attention kernel (K8) at the bench shape; B200W_LIB selects an A/B build."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from whisper_mlx_b200 import _lib as L
lib = L.load()
B, T, H, d = (int(sys.argv[1]) if len(sys.argv) > 1 else 120), 1500, 20, 1280
ckv = torch.randn(3, B, T, 2 * d, device="cuda").bfloat16()
q = torch.randn(B, 1, d, device="cuda").bfloat16()
o = torch.empty_like(q)
slot = torch.arange(B, dtype=torch.int32, device="cuda")
ts = []
for i in range(15):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    L.check(lib.b200w_decoder_cross_attention(L.ptr(q), B, 1, H, L.ptr(ckv[i % 3]), T * 2 * d, T, L.ptr(slot), L.ptr(o), L.stream()))
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
print(os.environ.get("B200W_LIB", "default"), "B", B, "CA us", [round(t, 1) for t in sorted(ts)[:5]])
