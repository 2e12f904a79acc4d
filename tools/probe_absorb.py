"""Development probe for K14 (absorb.cu): which tensor-memory lanes hold the rows of an M = 64 accumulator, and which
(LBO, SBO) pair makes a 128-row MN-major A operand out of two 64-feature TMA boxes.  Prints JSON."""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from whisper_mlx_b200 import _lib  # noqa: E402

lib = _lib.load()
fn = lib.b200w_debug_absorb_probe
fn.restype = C.c_int
fn.argtypes = [C.c_void_p] * 3 + [C.c_uint, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p]
torch.manual_seed(0)
dev = "cuda:0"
x = torch.randn(64, 128, device=dev).to(torch.bfloat16)
q = torch.randn(24, 128, device=dev).to(torch.bfloat16)
p = torch.rand(32, 64, device=dev).to(torch.bfloat16)
s_ref = x.float() @ q.float().T          # (64 keys, 24)
o_ref = x.float().T @ p.float().T        # (128 features, 32)
out = {}
for lbo, sbo in ((8192, 1024), (1024, 8192), (8192, 8192), (16, 1024)):
    ds = torch.zeros(128, 32, device=dev)
    do = torch.zeros(128, 32, device=dev)
    rc = fn(x.data_ptr(), q.data_ptr(), p.data_ptr(), lbo, sbo, ds.data_ptr(), do.data_ptr(), None)
    torch.cuda.synchronize()
    assert rc == 0, lib.b200w_last_error()
    lanes = []
    for r in range(64):
        err = (ds[:, :24] - s_ref[r][None]).abs().max(dim=1).values
        lanes.append(int(err.argmin()) if float(err.min()) < 0.05 else -1)
    out[f"lbo{lbo}_sbo{sbo}"] = {"s_row_to_lane": lanes, "o_max_err": float((do - o_ref).abs().max()),
                                 "o_ref_absmax": float(o_ref.abs().max())}
print(json.dumps(out))
