"""Times the log-mel front-end at BASELINE config 2 (1024 x 30 s) in its one-pass (fused clamp) and two-pass forms."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from whisper_mlx_b200.audio import log_mel_spectrogram, log_mel_unclamped

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
x = torch.randn(n, 480000, device="cuda") * 0.1
out = {}


def timed(fn, reps=10):
    fn(); fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


for n_mels in (80, 128):
    os.environ.pop("B200W_LOGMEL_TWO_PASS", None)
    one = log_mel_spectrogram(x, n_mels=n_mels)
    t1 = timed(lambda: log_mel_spectrogram(x, n_mels=n_mels))
    t0 = timed(lambda: log_mel_unclamped(x, n_mels))
    os.environ["B200W_LOGMEL_TWO_PASS"] = "1"
    two = log_mel_spectrogram(x, n_mels=n_mels)
    t2 = timed(lambda: log_mel_spectrogram(x, n_mels=n_mels))
    os.environ.pop("B200W_LOGMEL_TWO_PASS", None)
    bytes_alg = n * (4 * 480000 + 4 * 3000 * n_mels)
    out[n_mels] = {"one_pass_ms": t1, "two_pass_ms": t2, "kernel_only_unclamped_ms": t0, "identical": bool(torch.equal(one, two)),
                   "one_pass_frac_of_hbm_6540": bytes_alg / (t1 * 1e-3) / 6540.2e9, "frames_per_s": n * 3000 / (t1 * 1e-3)}
    del one, two
print(json.dumps(out))
