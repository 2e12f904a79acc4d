"""A/B of the decode chain (K11): runs the same greedy decode with B200W_CHAIN=0 and =1 in two processes and checks
that every sampled token and the last logits are bit-identical (the chain runs the same arithmetic in the same order)."""
import hashlib, json, os, subprocess, sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child(model_name, windows, steps):
    import torch
    sys.path.insert(0, REPO)
    from bench import build_model
    from whisper_mlx_b200.decoding import DecodingOptions, DecodingTask, DecodeSession

    model, _ = build_model(model_name, 0, "cuda:0")
    dm = model.dims
    g = torch.Generator().manual_seed(1)
    xa = torch.randn(windows, dm.n_audio_ctx, dm.n_audio_state, generator=g).bfloat16().cuda()
    task = DecodingTask(model, DecodingOptions(language="en"))
    sess = DecodeSession(model, xa, 1, max_tokens=3 + steps + 1)
    sess.set_tokens(torch.tensor(task.initial_tokens, dtype=torch.int32).repeat(windows, 1))
    sess.set_filter(task._filter_params(sess), task._get_suppress_tokens())
    sess.prompt_step(len(task.initial_tokens), task.sot_index)
    for _ in range(steps):
        sess.sample_step()
    torch.cuda.synchronize()
    toks = sess.tokens[:, : 3 + steps + 1].cpu().numpy().tobytes()
    lg = sess.logits[:, : dm.n_vocab].cpu().numpy().tobytes()
    print(json.dumps({"tokens": hashlib.sha256(toks).hexdigest(), "logits": hashlib.sha256(lg).hexdigest(),
                      "sum_logprob": float(sess.sum_logprob.sum().item())}))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]))
        sys.exit(0)
    model = sys.argv[1] if len(sys.argv) > 1 else "large-v3"
    windows = int(sys.argv[2]) if len(sys.argv) > 2 else 120
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    res = {}
    for flag in ("0", "1", "1"):
        # the unfused path plans its K splits like the (multicast) chain does, so that the sums run in the same order
        env = dict(os.environ, B200W_CHAIN=flag, B200W_SPLIT_PLAN="mc")
        out = subprocess.run([sys.executable, __file__, "child", model, str(windows), str(steps)], env=env, capture_output=True,
                             text=True, timeout=600)
        if out.returncode != 0:
            print("child failed", flag, out.stderr[-2000:])
            sys.exit(1)
        res.setdefault(flag, []).append(json.loads(out.stdout.strip().splitlines()[-1]))
    print(json.dumps(res))
    same = all(r == res["0"][0] for r in res["1"])
    print("IDENTICAL" if same else "MISMATCH")
    sys.exit(0 if same else 1)
