"""Tiny driver for `ncu` on the one-launch small-batch decode step (K13): a few eager steps at batch 1, large-v3."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import build_model
from whisper_mlx_b200.decoding import DecodingOptions, DecodingTask, DecodeSession

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
model, _ = build_model("large-v3", 0, "cuda:0")
dm = model.dims
xa = torch.randn(B, dm.n_audio_ctx, dm.n_audio_state, device="cuda").bfloat16()
task = DecodingTask(model, DecodingOptions(language="en"))
sess = DecodeSession(model, xa, 1, max_tokens=3 + 224)
sess.set_tokens(torch.tensor(task.initial_tokens, dtype=torch.int32).repeat(B, 1))
sess.set_filter(task._filter_params(sess), task._get_suppress_tokens())
sess.prompt_step(len(task.initial_tokens), task.sot_index)
for _ in range(6):
    sess._step(1, -1, True)
torch.cuda.synchronize()
print("ok", int(sess.n_tokens[0]))
