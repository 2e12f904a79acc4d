"""Exact-mode throughput of several files: one after the other (transcribe) against in lockstep (transcribe_many).

    python tools/time_many.py [--files 3] [--seconds 90]
"""
import argparse, json, os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import build_model, make_audio
import whisper_mlx_b200.transcribe  # noqa: F401
T = sys.modules["whisper_mlx_b200.transcribe"]

ap = argparse.ArgumentParser()
ap.add_argument("--files", type=int, default=3)
ap.add_argument("--seconds", type=float, default=90.0)
a = ap.parse_args()
model, _ = build_model("large-v3", 0, "cuda:0")
audios = [torch.from_numpy(make_audio(a.seconds / 3600.0, 300 + i)).cuda() for i in range(a.files)]
kw = dict(model=model, temperature=0.0, condition_on_previous_text=False, language="en")
T.transcribe(audios[0][: 16000 * 31], **kw)
T.transcribe_many([x[: 16000 * 31] for x in audios], **kw)
torch.cuda.synchronize(); t0 = time.perf_counter()
for x in audios:
    T.transcribe(x, **kw)
torch.cuda.synchronize(); t_seq = time.perf_counter() - t0
t0 = time.perf_counter()
res = T.transcribe_many(audios, **kw)
torch.cuda.synchronize(); t_many = time.perf_counter() - t0
assert all(isinstance(r, dict) for r in res), res
print(json.dumps({"files": a.files, "seconds_each": a.seconds, "one_after_the_other_s": round(t_seq, 3), "lockstep_s": round(t_many, 3),
                  "rtfx_sequential": round(a.files * a.seconds / t_seq, 1), "rtfx_lockstep": round(a.files * a.seconds / t_many, 1)}))
