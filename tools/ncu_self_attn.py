"""Driver for timing / `ncu --set full` on the decode self-attention kernel (K7) at the bench shape
(120 sequences, large-v3 heads) with `--pos` cached keys per sequence.  Prints CUDA-event times per launch."""
import argparse, json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from whisper_mlx_b200 import _lib as L

ap = argparse.ArgumentParser()
ap.add_argument("--pos", type=int, default=220)
ap.add_argument("--layers", type=int, default=4)
ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()
lib = L.load()
B, H, d, ps, max_pages = 120, 20, 1280, 16, 15
n_pages = B * max_pages
kp = torch.randn(a.layers, n_pages, ps, d, device="cuda").bfloat16()
vp = torch.randn(a.layers, n_pages, ps, d, device="cuda").bfloat16()
bt = torch.arange(n_pages, dtype=torch.int32, device="cuda").view(B, max_pages).contiguous()
pos = torch.full((B,), a.pos, dtype=torch.int32, device="cuda")
qkv = torch.randn(B, 1, 3 * d, device="cuda").bfloat16()
o = torch.empty(B, 1, d, device="cuda", dtype=torch.bfloat16)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
times = []
for r in range(a.reps):
    for l in range(a.layers):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.check(lib.b200w_decoder_self_attention(L.ptr(qkv), B, 1, H, L.ptr(pos), L.ptr(kp[l]), L.ptr(vp[l]), L.ptr(bt),
                                                 max_pages, ps, L.ptr(o), L.stream()))
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1) * 1e3)
times = sorted(times[a.layers:])
nbytes = B * (a.pos + 1) * d * 2 * 2
print(json.dumps({"pos": a.pos, "us_median": times[len(times) // 2], "us_min": times[0], "MB": nbytes / 1e6,
                  "GBps_median": nbytes / times[len(times) // 2] / 1e3}))
