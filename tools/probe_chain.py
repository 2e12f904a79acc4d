"""Cost of the decode chain's grid barrier: chains of 1..6 empty phases (development probe)."""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from whisper_mlx_b200 import _lib as L
lib = L.load()
fn = lib.b200w_debug_chain_barriers
fn.restype = C.c_int
fn.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
cnt = torch.zeros(64, dtype=torch.int32, device="cuda")
x = torch.zeros(128 * 128, device="cuda")
h = torch.zeros(128 * 128, dtype=torch.bfloat16, device="cuda")
g = torch.ones(128, device="cuda")
res = {}
side = torch.cuda.Stream()
for n in (1, 2, 4, 6):
    g_ = torch.cuda.CUDAGraph()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        g_.capture_begin()
        cnt.zero_()
        for k in range(32):  # 32 launches back to back, one counter each
            L.check(fn(n, cnt[k:].data_ptr(), x.data_ptr(), h.data_ptr(), g.data_ptr(), L.stream()))
        g_.capture_end()
    torch.cuda.current_stream().wait_stream(side)
    ts = []
    for it in range(20):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g_.replay()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / 32)
    res[n] = sorted(ts)[len(ts) // 2]
print({k: round(v, 2) for k, v in res.items()}, "us per launch in a graph; barrier ~", round((res[6] - res[2]) / 4, 2), "us")
