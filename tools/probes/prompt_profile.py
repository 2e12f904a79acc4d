"""Per-kernel times of the prompt step (n_q = 3 tokens x 120 sequences, large-v3)."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from bench import build_model
from whisper_mlx_b200._lib import kernel_profile
from whisper_mlx_b200.decoding import DecodingOptions, DecodingTask, DecodeSession
model, _ = build_model("large-v3", 0, "cuda:0")
dm = model.dims
B = int(sys.argv[1]) if len(sys.argv) > 1 else 120
xa = torch.randn(B, dm.n_audio_ctx, dm.n_audio_state, device="cuda").bfloat16()
task = DecodingTask(model, DecodingOptions(language="en"))
sess = DecodeSession(model, xa, 1, max_tokens=3 + 224)
for rep in range(2):
    sess.set_tokens(torch.tensor(task.initial_tokens, dtype=torch.int32).repeat(B, 1))
    sess.set_filter(task._filter_params(sess), task._get_suppress_tokens())
    with kernel_profile() as prof:
        sess.prompt_step(len(task.initial_tokens), task.sot_index)
print(json.dumps({k: {"n": v["launches"], "ms": round(v["total_ms"], 3)} for k, v in sorted(prof.result.items(), key=lambda kv: -kv[1]["total_ms"])}))
