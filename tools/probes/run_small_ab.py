"""A/B of K13 builds in one process each: graph-replayed step time at batch 1 / 2 / 3 (B200W_LIB selects the build)."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from bench import build_model
from whisper_mlx_b200.decoding import DecodingOptions, DecodingTask, DecodeSession
model, _ = build_model("large-v3", 0, "cuda:0")
dm = model.dims
os.environ["B200W_SMALL"] = "1"
out = {}
for B in (1, 2, 3):
    xa = torch.randn(B, dm.n_audio_ctx, dm.n_audio_state, device="cuda").bfloat16()
    task = DecodingTask(model, DecodingOptions(language="en"))
    sess = DecodeSession(model, xa, 1, max_tokens=3 + 224)
    sess.set_tokens(torch.tensor(task.initial_tokens, dtype=torch.int32).repeat(B, 1))
    sess.set_filter(task._filter_params(sess), task._get_suppress_tokens())
    sess.prompt_step(len(task.initial_tokens), task.sot_index)
    sess.sample_step(); torch.cuda.synchronize()
    ts = []
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(60):
            sess.sample_step()
        e1.record(); torch.cuda.synchronize()
        ts.append(round(e0.elapsed_time(e1) / 60, 4))
    out[B] = ts
    del sess
print(os.path.basename(os.environ.get("B200W_LIB", "default")), json.dumps(out))
