"""Phase timeline of one decode-chain launch (K11): clock64 stamps of CTA 0 and CTA grid/2 for every phase.

    python tools/probe_chain_timeline.py [windows] [launch_index]   (launch 2 + 2 l = chain C of layer l, 1 + 2 l = chain B)
stamps per phase: 0 start, 1 loads issued, 2 first operands landed, 3 last MMA issued, 4 last accumulator complete,
5 epilogue done, 6 barrier arrive, 7 barrier released
"""
import ctypes as C, json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import build_model
from whisper_mlx_b200 import _lib as L
from whisper_mlx_b200.decoding import DecodingOptions, DecodingTask, DecodeSession

B = int(sys.argv[1]) if len(sys.argv) > 1 else 120
launch = int(sys.argv[2]) if len(sys.argv) > 2 else 12
model, _ = build_model("large-v3", 0, "cuda:0")
dm = model.dims
lib = L.load()
xa = torch.randn(B, dm.n_audio_ctx, dm.n_audio_state, device="cuda").bfloat16()
task = DecodingTask(model, DecodingOptions(language="en"))
sess = DecodeSession(model, xa, 1, max_tokens=3 + 224)
sess.set_tokens(torch.tensor(task.initial_tokens, dtype=torch.int32).repeat(B, 1))
sess.set_filter(task._filter_params(sess), task._get_suppress_tokens())
sess.prompt_step(len(task.initial_tokens), task.sot_index)
for _ in range(8):
    sess._step(1, -1, True)
torch.cuda.synchronize()
fn = lib.b200w_debug_chain_timeline
fn.restype, fn.argtypes = None, [C.c_void_p, C.c_int]
buf = torch.zeros(2 * 6 * 8, dtype=torch.int64, device="cuda")
acc = None
for rep in range(5):
    buf.zero_()
    fn(C.c_void_p(buf.data_ptr()), launch)
    sess._step(1, -1, True)
    torch.cuda.synchronize()
    t = buf.view(2, 6, 8).cpu().double()
    acc = t if acc is None else acc
fn(None, 0)
t = acc
out = {}
for c in range(2):
    t0 = float(t[c, 0, 0])
    rows = []
    for ph in range(6):
        if float(t[c, ph, 0]) == 0:
            continue
        rows.append([None if float(v) == 0 else round((float(v) - t0) / 1.9, 0) for v in t[c, ph]])  # ns at 1.9 GHz
    out[f"cta{c}"] = rows
print(json.dumps(out))
