"""Decode-step time of one batch as ONE stream vs TWO half-batches on two streams (B200W_DECODE_STREAMS), through
DecodingTask.run_features: wall time of `--steps` greedy steps, per step.

    python tools/time_dual.py [--windows 120] [--steps 96]
"""
import argparse, json, os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import build_model
from whisper_mlx_b200.decoding import DecodingOptions, DecodingTask

ap = argparse.ArgumentParser()
ap.add_argument("--windows", type=int, default=120)
ap.add_argument("--steps", type=int, default=96)
ap.add_argument("--model", default="large-v3")
a = ap.parse_args()
model, _ = build_model(a.model, 0, "cuda:0")
dm = model.dims
xa = torch.randn(a.windows, dm.n_audio_ctx, dm.n_audio_state, device="cuda").bfloat16()
out = {}
toks = {}


def run(steps):
    task = DecodingTask(model, DecodingOptions(language="en", sample_len=steps))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = task.run_features(xa)
    torch.cuda.synchronize()
    return time.perf_counter() - t0, res


for n_streams in (1, 2):  # sessions, graphs
    model.decode_streams = n_streams
    run(8), run(a.steps + 8)
for n_streams in (1, 2, 1, 2):
    model.decode_streams = n_streams
    t8, _ = run(8)
    tn, res = run(a.steps + 8)
    toks[n_streams] = [r.tokens for r in res]
    out.setdefault(str(n_streams), []).append(round((tn - t8) / a.steps * 1e3, 4))
print(json.dumps({"windows": a.windows, "steps": a.steps, "ms_per_step": out, "same_tokens": toks[1] == toks[2]}))
