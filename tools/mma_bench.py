"""Cycles per tcgen05.mma (kind::f16, K = 16) for the shapes the absorbed cross-attention could use."""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from whisper_mlx_b200 import _lib as L  # noqa: E402

lib = L.load()
fn = lib.b200w_debug_mma_bench
fn.argtypes = [C.c_int] * 5 + [C.c_void_p, C.c_void_p]
out = {}
cyc = torch.zeros(2, dtype=torch.int64, device="cuda:0")
reps = 64
for (m, n, a_mn, ts) in ((64, 24, 0, 0), (64, 32, 0, 0), (64, 64, 0, 0), (64, 128, 0, 0), (64, 256, 0, 0),
                         (128, 16, 0, 0), (128, 32, 0, 0), (128, 64, 0, 0), (128, 128, 0, 0), (128, 256, 0, 0),
                         (128, 32, 1, 0), (128, 64, 1, 0), (128, 128, 1, 0), (128, 256, 1, 0),
                         (128, 32, 0, 1), (128, 64, 0, 1), (128, 128, 0, 1), (128, 256, 0, 1)):
    for _ in range(2):
        L.check(fn(m, n, a_mn, ts, reps, cyc.data_ptr(), None))
        torch.cuda.synchronize()
    c = cyc.cpu()
    out[f"M{m}_N{n}_{'mnA' if a_mn else 'kA'}{'_ts' if ts else ''}"] = {"issue_per_mma": int(c[0]) / (4 * reps),
                                                                         "total_per_mma": int(c[1]) / (4 * reps)}
print(json.dumps(out, indent=1))
