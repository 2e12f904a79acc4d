"""Experiment: two half-batches decoded on two streams with their cross-attention launches chained by events
(anti-phase: one half streams K / V while the other runs its chains), eager launches.  Compares against one stream.

    python tools/time_dual_eager.py [--windows 120] [--steps 40]
"""
import argparse, ctypes as C, json, os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import build_model
from whisper_mlx_b200 import _lib as L
from whisper_mlx_b200.decoding import DecodingOptions, DecodingTask, DecodeSession

ap = argparse.ArgumentParser()
ap.add_argument("--windows", type=int, default=120)
ap.add_argument("--steps", type=int, default=40)
a = ap.parse_args()
model, _ = build_model("large-v3", 0, "cuda:0")
dm = model.dims
lib = L.load()
role = lib.b200w_debug_cross_role
role.restype, role.argtypes = None, [C.c_int]


def session(B):
    xa = torch.randn(B, dm.n_audio_ctx, dm.n_audio_state, device="cuda").bfloat16()
    task = DecodingTask(model, DecodingOptions(language="en"))
    sess = DecodeSession(model, xa, 1, max_tokens=3 + 224)
    sess.set_tokens(torch.tensor(task.initial_tokens, dtype=torch.int32).repeat(B, 1))
    sess.set_filter(task._filter_params(sess), task._get_suppress_tokens())
    sess.prompt_step(len(task.initial_tokens), task.sot_index)
    for _ in range(16):
        sess.sample_step()
    torch.cuda.synchronize()
    return sess


def eager_step(sess):
    sess._step(1, -1, True)


out = {}
one = session(a.windows)
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(a.steps):
        eager_step(one)
    torch.cuda.synchronize()
    out.setdefault("one_stream_eager_ms", []).append(round((time.perf_counter() - t0) / a.steps * 1e3, 3))
del one
halves = [session(a.windows // 2), session(a.windows - a.windows // 2)]
streams = [torch.cuda.Stream(), torch.cuda.Stream()]
for mode in ("free", "antiphase", "free", "antiphase"):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(a.steps):
        for i in (0, 1):
            role(i if mode == "antiphase" else -1)
            with torch.cuda.stream(streams[i]):
                eager_step(halves[i])
    role(-1)
    torch.cuda.synchronize()
    out.setdefault("two_streams_" + mode + "_ms", []).append(round((time.perf_counter() - t0) / a.steps * 1e3, 3))
# per-kernel times (CUDA events around every launch of the library) of one stream alone and of the anti-phase pair
from whisper_mlx_b200._lib import kernel_profile


def shares(prof):
    return {k: round(v["total_ms"] / v["launches"] * 1e3, 1) for k, v in prof.result.items() if k in ("dec_chain", "decoder_self_attention", "decoder_cross_attention")}


with kernel_profile() as prof:
    with torch.cuda.stream(streams[0]):
        eager_step(halves[0])
    torch.cuda.synchronize()
out["half_batch_alone_avg_us"] = shares(prof)
with kernel_profile() as prof:
    for _ in range(2):
        for i in (0, 1):
            role(i)
            with torch.cuda.stream(streams[i]):
                eager_step(halves[i])
    role(-1)
    torch.cuda.synchronize()
out["antiphase_avg_us"] = shares(prof)
print(json.dumps(out))
