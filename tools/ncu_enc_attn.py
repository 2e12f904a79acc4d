"""Driver for timing / `ncu --set full` on the encoder attention kernel (K6) at the large-v3 shape."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from whisper_mlx_b200 import _lib as L
lib = L.load()
B, T, H, d = int(sys.argv[1]) if len(sys.argv) > 1 else 40, 1500, 20, 1280
qkv = (torch.randn(B, T, 3 * d, device="cuda") * 0.5).bfloat16()
o = torch.empty(B, T, d, device="cuda", dtype=torch.bfloat16)
ts = []
for i in range(6):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    L.check(lib.b200w_encoder_attention(L.ptr(qkv), B, T, H, L.ptr(o), L.stream()))
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
flops = 4.0 * B * H * T * T * 64
print("enc attention ms", [round(t, 3) for t in ts], "TF/s", round(flops / (min(ts) * 1e-3) / 1e12, 1))
