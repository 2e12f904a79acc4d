"""Do the persistent cross-attention (K8p) and the decode chain kernel (K11) share the SMs?  K8p launches on one stream,
chain launches (barrier-only phases: same kernel, same footprint) on another; alone vs together."""
import ctypes as C, os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from whisper_mlx_b200 import _lib as L
lib = L.load()
fn = lib.b200w_debug_chain_barriers
fn.restype, fn.argtypes = C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
B, T, H, d = 60, 1500, 20, 1280
ckv = torch.randn(3, B, T, 2 * d, device="cuda").bfloat16()
q = torch.randn(B, 1, d, device="cuda").bfloat16()
o = torch.empty_like(q)
slot = torch.arange(B, dtype=torch.int32, device="cuda")
cnt = torch.zeros(4096, dtype=torch.int32, device="cuda")
x = torch.zeros(128 * 128, device="cuda")
h = torch.zeros(128 * 128, dtype=torch.bfloat16, device="cuda")
g = torch.ones(128, device="cuda")
sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
N = 40


def cross():
    with torch.cuda.stream(sa):
        for i in range(N):
            L.check(lib.b200w_decoder_cross_attention(L.ptr(q), B, 1, H, L.ptr(ckv[i % 3]), T * 2 * d, T, L.ptr(slot), L.ptr(o), L.stream()))


def chain(k0):
    with torch.cuda.stream(sb):
        for i in range(N):
            L.check(fn(6, cnt[k0 + i:].data_ptr(), x.data_ptr(), h.data_ptr(), g.data_ptr(), L.stream()))


def timed(f):
    torch.cuda.synchronize(); t0 = time.perf_counter(); f(); torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3


cross(); chain(0); cnt.zero_(); torch.cuda.synchronize()
ta = timed(cross)
tb = timed(lambda: chain(100)); cnt.zero_()
tab = timed(lambda: (cross(), chain(200))); cnt.zero_()
tba = timed(lambda: (chain(300), cross()))
print({"cross_alone_ms": round(ta, 3), "chain_alone_ms": round(tb, 3), "together_ms": round(tab, 3), "together_chain_first_ms": round(tba, 3)})
