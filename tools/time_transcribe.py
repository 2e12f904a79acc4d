"""Phase breakdown of one transcribe() call at the bench shapes (development aid): wraps the phase entry
points with synchronising timers."""
import os, sys, time, json
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from bench import build_model, make_audio
import whisper_mlx_b200.transcribe  # noqa: F401 (the package attribute of that name is the function)
T = sys.modules['whisper_mlx_b200.transcribe']
import whisper_mlx_b200.decoding as D
import whisper_mlx_b200.audio as A

model, _ = build_model(sys.argv[1] if len(sys.argv) > 1 else "large-v3", 0, "cuda:0")
audio = torch.from_numpy(make_audio(1.0, 100)).cuda()
acc = {}
def wrap(obj, name, label):
    f = getattr(obj, name)
    def g(*a, **k):
        torch.cuda.synchronize(); t = time.perf_counter()
        r = f(*a, **k)
        torch.cuda.synchronize(); acc[label] = acc.get(label, 0.0) + time.perf_counter() - t
        return r
    setattr(obj, name, g)
wrap(T, "log_mel_unclamped", "logmel")
wrap(model, "encode_slabs", "encoder")
wrap(model, "mel_windows", "mel_windows")
wrap(model, "cross_kv", "cross_kv")
orig_init = D.DecodeSession.__init__
def init(self, *a, **k):
    torch.cuda.synchronize(); t = time.perf_counter(); orig_init(self, *a, **k); torch.cuda.synchronize()
    acc["session_init(incl cross_kv)"] = acc.get("session_init(incl cross_kv)", 0.0) + time.perf_counter() - t
D.DecodeSession.__init__ = init
wrap(D.DecodeSession, "prompt_step", "prompt_step")
orig_rf = D.DecodingTask.run_features
def rf(self, *a, **k):
    torch.cuda.synchronize(); t = time.perf_counter(); r = orig_rf(self, *a, **k); torch.cuda.synchronize()
    acc["run_features_total"] = acc.get("run_features_total", 0.0) + time.perf_counter() - t
    return r
D.DecodingTask.run_features = rf
kw = dict(model=model, temperature=0.0, condition_on_previous_text=False, language="en", window_batch=120, encoder_batch=40)
for i in range(3):
    acc.clear()
    torch.cuda.synchronize(); t = time.perf_counter()
    r = T.transcribe(audio, **kw)
    torch.cuda.synchronize(); tot = time.perf_counter() - t
    print(json.dumps({"total": tot, **acc}))
print("segments", len(r["segments"]), "tokens/window", sum(len(s["tokens"]) for s in r["segments"]) / 120)
