"""Times the absorbed cross-attention (K14a + K14b + K14c) against K8 on the K/V cache it replaces.

    python tools/time_absorb.py [B] [H] -> JSON (ms per launch, CUDA events over `iters` launches; inputs >> L2)
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from whisper_mlx_b200 import _lib as L  # noqa: E402

lib = L.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 120
H = int(sys.argv[2]) if len(sys.argv) > 2 else 20
T, d = 1500, 64 * H
dev = "cuda:0"
torch.manual_seed(0)
xa = torch.randn(B, T, d, device=dev).to(torch.bfloat16)
w = (torch.randn(2 * d, d, device=dev) / d ** 0.5).to(torch.bfloat16)
bias = torch.cat([torch.zeros(d), torch.randn(d)]).to(dev)
q = torch.randn(B, d, device=dev).to(torch.bfloat16)
slot = torch.arange(B, dtype=torch.int32, device=dev)
ws = torch.empty(lib.b200w_absorbed_cross_attention_workspace_bytes(B, H), dtype=torch.uint8, device=dev)
o = torch.empty((B, d), dtype=torch.bfloat16, device=dev)
kv = torch.empty((B, T, 2 * d), dtype=torch.bfloat16, device=dev)
for b0 in range(0, B, 8):
    kv[b0:b0 + 8] = (xa[b0:b0 + 8].float() @ w.float().T + bias).to(torch.bfloat16)
o8 = torch.empty((B, 1, d), dtype=torch.bfloat16, device=dev)


def absorbed():
    L.check(lib.b200w_absorbed_cross_attention(L.ptr(q), B, H, L.ptr(w), L.ptr(bias), L.ptr(xa), B, T, L.ptr(slot), None,
                                               L.ptr(ws), ws.numel(), L.ptr(o), L.stream()))


def k8():
    L.check(lib.b200w_decoder_cross_attention(L.ptr(q), B, 1, H, L.ptr(kv), T * 2 * d, T, L.ptr(slot), L.ptr(o8), L.stream()))


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


res = {"B": B, "H": H, "absorbed_ms": timeit(absorbed), "k8_ms": timeit(k8)}
res["xa_bytes"] = B * T * d * 2
res["absorbed_GBps_on_xa"] = res["xa_bytes"] / res["absorbed_ms"] / 1e6
res["k8_GBps_on_kv"] = 2 * res["xa_bytes"] / res["k8_ms"] / 1e6
L.check(lib.b200w_profile_begin())
for _ in range(5):
    absorbed()
torch.cuda.synchronize()
buf = (__import__("ctypes").c_char * 65536)()
lib.b200w_profile_end(buf, 65536)
res["profile"] = json.loads(buf.value.decode())
tl = torch.zeros(64 * 8, dtype=torch.int64, device=dev)
lib.b200w_debug_absorb_timeline.argtypes = [__import__("ctypes").c_void_p]
lib.b200w_debug_absorb_timeline(tl.data_ptr())
absorbed()
torch.cuda.synchronize()
lib.b200w_debug_absorb_timeline(None)
t = tl.view(64, 8).cpu()
t0 = int(t[0, 0])
res["timeline_cycles"] = [[int(v - t0) if v else None for v in row[:7]] for row in t[:20]]
res["err_vs_k8"] = float((o.float() - o8[:, 0].float()).abs().max())
print(json.dumps(res))
