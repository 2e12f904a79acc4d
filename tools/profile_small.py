"""Phase timeline of the one-launch small-batch decode step (K13): %globaltimer stamps of CTA 0 at every grid barrier.

    python tools/profile_small.py [model] [batch] [position]
"""
import ctypes as C, json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import build_model
from whisper_mlx_b200 import _lib as L
from whisper_mlx_b200.decoding import DecodingOptions, DecodingTask, DecodeSession

name = sys.argv[1] if len(sys.argv) > 1 else "large-v3"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
position = int(sys.argv[3]) if len(sys.argv) > 3 else 40
model, _ = build_model(name, 0, "cuda:0")
dm = model.dims
lib = L.load()
xa = torch.randn(B, dm.n_audio_ctx, dm.n_audio_state, device="cuda").bfloat16()
task = DecodingTask(model, DecodingOptions(language="en"))
sess = DecodeSession(model, xa, 1, max_tokens=3 + 224)
sess.set_tokens(torch.tensor(task.initial_tokens, dtype=torch.int32).repeat(B, 1))
sess.set_filter(task._filter_params(sess), task._get_suppress_tokens())
sess.prompt_step(len(task.initial_tokens), task.sot_index)
for _ in range(position):
    sess.sample_step()
torch.cuda.synchronize()
n = 2 * (8 * dm.n_text_layer + 1) + 2 + 8
buf = torch.zeros(n, dtype=torch.int64, device="cuda")
fn = lib.b200w_debug_small_timeline
fn.restype, fn.argtypes = None, [C.c_void_p]
fn(C.c_void_p(buf.data_ptr()))
acc = None
reps = 20
for _ in range(reps):
    sess._step(1, -1, True)  # eager: the graph captured above has no timeline pointer
    torch.cuda.synchronize()
    t = buf.cpu().numpy().astype("float64")
    e_ = 8 * dm.n_text_layer
    cyc = t[2 * e_ + 2: 2 * e_ + 10].copy()
    t = t - t[0]
    t[2 * e_ + 2: 2 * e_ + 10] = cyc
    acc = t if acc is None else acc + t
fn(C.c_void_p(0))
t = acc / reps / 1e3  # us
names = ["ln+qkv", "self_attn", "out_proj", "ln+cross_q", "cross_attn", "cross_out", "ln+mlp1", "mlp2"]
work = {k: 0.0 for k in names}
wait = {k: 0.0 for k in names}
L_ = dm.n_text_layer
for l in range(L_):
    for j, k in enumerate(names):
        e = 8 * l + j
        start = t[2 * e] if e > 0 else 0.0       # departure from the previous barrier
        work[k] += t[2 * e + 1] - start            # phase body on CTA 0
        wait[k] += t[2 * e + 2] - t[2 * e + 1]     # time CTA 0 sat in the barrier after it
e = 8 * L_
out = {"model": name, "batch": B, "position": position, "step_us_cta0": t[2 * e + 1],
       "per_layer_us": {k: {"body": work[k] / L_, "barrier_wait": wait[k] / L_} for k in names},
       "logits_us": t[2 * e + 1] - t[2 * e],
       "sum_body_us": sum(work.values()), "sum_barrier_wait_us": sum(wait.values())}
e = 8 * L_
cyc = acc[2 * e + 2: 2 * e + 10] / reps
out["cycles_per_layer_thread0_cta0"] = {k: cyc[i] / L_ for i, k in enumerate(["input_ln", "wait_weights", "dot_reduce", "epilogue", "self_attn", "cross_attn", "input_vec_or_merge", "-"])}
print(json.dumps(out, indent=1))
