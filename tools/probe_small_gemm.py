"""Small-M (decode) GEMM latency probe: CUDA-graph of 32 launches per shape, cold (rotating weights) and hot."""
import os, sys, json
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from whisper_mlx_b200 import _lib as L
lib = L.load()
M = int(sys.argv[1]) if len(sys.argv) > 1 else 120
out = {}
for (N, K) in [(1280, 64), (1280, 320), (1280, 1280), (1280, 5120), (3840, 1280), (5120, 1280), (3840, 320), (5120, 320), (1280, 640), (51866, 1280)]:
    n_copies = 32 if N * K * 2 * 32 < 8e9 else 4
    a = torch.randn(M, K, device="cuda").bfloat16()
    ws = [(torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16() for _ in range(n_copies)]
    ld = (N + 127) // 128 * 128
    c = torch.empty((M, ld), dtype=torch.float32, device="cuda")
    for mode in ("cold", "hot"):
        def launch(i):
            w = ws[i % n_copies] if mode == "cold" else ws[0]
            L.check(lib.b200w_gemm_bf16(L.ptr(a), K, L.ptr(w), L.ptr(c), ld, None, None, M, N, K, 2, L.stream()))
        launch(0); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(32):
                launch(i)
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            g.replay()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / (5 * 32) * 1e3
        out[f"{M}x{N}x{K}_{mode}"] = us
        print(f"M={M} N={N} K={K} {mode}: {us:.2f} us/launch  ({N*K*2/us/1e3:.0f} GB/s weights)", flush=True)
json.dump(out, open(os.path.join(REPO, "gpurun_out", f"probe_small_gemm_{M}.json"), "w"), indent=1)
