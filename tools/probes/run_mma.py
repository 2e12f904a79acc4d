"""K13m (one-launch step on mma.sync) against the chain path: teacher-forced logits (whisper-small dims, random weights)
and graph-replayed step times at large-v3."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tools import synth
from whisper_mlx_b200.whisper import ModelDimensions, Whisper
from whisper_mlx_b200 import _lib as L

lib = L.load()
out = {}
if "--time-only" not in sys.argv:
    dims_d = synth.DIMS["small"]
    m = Whisper(ModelDimensions(**dims_d), dict(synth.random_weights(dims_d, 0, device="cuda")))
    g = torch.Generator().manual_seed(5)
    xa = (torch.randn(16, dims_d["n_audio_ctx"], dims_d["n_audio_state"], generator=g) * 0.7).bfloat16()
    seq = [50258, 50259, 50359, 50364 + 5, 300, 4000, 17, 50364 + 80, 50364 + 80, 900, 901, 902, 50364 + 200, 50364 + 200, 12, 13, 14, 15, 16, 50364 + 400, 50364 + 400, 21]
    for B in (1, 4, 5, 8, 9, 15, 16):
        toks = torch.tensor([seq] * B, dtype=torch.long)
        toks[:, 6] += torch.arange(B)
        os.environ["B200W_SMALL_MMA"] = "all"
        k0 = lib.b200w_launch_count()
        got = m.logits(toks, xa[:B].cuda()).cpu()
        n_mma = lib.b200w_launch_count() - k0
        os.environ["B200W_SMALL_MMA"] = "0"
        os.environ["B200W_SMALL"] = "0"
        k0 = lib.b200w_launch_count()
        big = m.logits(toks, xa[:B].cuda()).cpu()
        n_big = lib.b200w_launch_count() - k0
        os.environ.pop("B200W_SMALL")
        out[f"B{B}"] = {"max_abs_diff": (got - big).abs().max().item(), "argmax_equal": bool((got.argmax(-1) == big.argmax(-1)).all()),
                        "launches": (n_mma, n_big), "finite": bool(torch.isfinite(got).all())}
    print(json.dumps(out))
    del m

from bench import build_model
from whisper_mlx_b200.decoding import DecodingOptions, DecodingTask, DecodeSession
model, _ = build_model("large-v3", 0, "cuda:0")
dm = model.dims
t = {}
for mode in ("all", "0"):
    os.environ["B200W_SMALL_MMA"] = mode
    for B in (1, 2, 3):
        xa = torch.randn(B, dm.n_audio_ctx, dm.n_audio_state, device="cuda").bfloat16()
        task = DecodingTask(model, DecodingOptions(language="en"))
        sess = DecodeSession(model, xa, 1, max_tokens=3 + 224)
        sess.set_tokens(torch.tensor(task.initial_tokens, dtype=torch.int32).repeat(B, 1))
        sess.set_filter(task._filter_params(sess), task._get_suppress_tokens())
        sess.prompt_step(len(task.initial_tokens), task.sot_index)
        sess.sample_step(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(60):
            sess.sample_step()
        e1.record(); torch.cuda.synchronize()
        t.setdefault("mma" if mode == "all" else "chain", {})[B] = round(e0.elapsed_time(e1) / 60, 4)
        del sess
print(json.dumps(t))
