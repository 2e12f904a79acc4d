"""Tiny driver for `ncu --set full` on the log-mel kernel (64 windows, one launch after a warm-up)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from whisper_mlx_b200.audio import log_mel_unclamped
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
x = torch.randn(n, 480000, device="cuda") * 0.1
for _ in range(2):
    out, g = log_mel_unclamped(x, 128)
torch.cuda.synchronize()
print(out.shape, float(g.max()))
