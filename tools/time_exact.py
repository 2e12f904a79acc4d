"""Exact (sequential-seek, batch 1) mode timing: the mode `transcribe()` / the CLI use by default (development aid)."""
import os, sys, time, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import build_model, make_audio
import whisper_mlx_b200.transcribe  # noqa: F401
T = sys.modules["whisper_mlx_b200.transcribe"]
from whisper_mlx_b200.decoding import DecodingOptions, DecodingTask, DecodeSession

model, _ = build_model(sys.argv[1] if len(sys.argv) > 1 else "large-v3", 0, "cuda:0")
minutes = float(sys.argv[2]) if len(sys.argv) > 2 else 3.0
audio = torch.from_numpy(make_audio(minutes / 60.0, 100)).cuda()
kw = dict(model=model, temperature=0.0, condition_on_previous_text=False, language="en")
for i in range(2):
    torch.cuda.synchronize(); t = time.perf_counter()
    r = T.transcribe(audio, **kw)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    print(json.dumps({"seconds": dt, "rtfx": minutes * 60 / dt, "segments": len(r["segments"])}))
# single decode steps from the CUDA graph at small batches, with the one-launch step (K13) and without it
dm = model.dims
for small in ("1", "0"):
    os.environ["B200W_SMALL"] = small
    for B in (1, 2, 3, 5):
        xa = torch.randn(B, dm.n_audio_ctx, dm.n_audio_state, device="cuda").bfloat16()
        task = DecodingTask(model, DecodingOptions(language="en"))
        sess = DecodeSession(model, xa, 1, max_tokens=3 + 224)
        sess.set_tokens(torch.tensor(task.initial_tokens, dtype=torch.int32).repeat(B, 1))
        sess.set_filter(task._filter_params(sess), task._get_suppress_tokens())
        sess.prompt_step(len(task.initial_tokens), task.sot_index)
        sess.sample_step(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(100):
            sess.sample_step()
        e1.record(); torch.cuda.synchronize()
        print(json.dumps({"one_launch_step": small == "1", "batch": B, "step_ms": e0.elapsed_time(e1) / 100, "kernels_per_step": sess._graph_kernels}))
        del sess
