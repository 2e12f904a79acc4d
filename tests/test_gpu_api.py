"""GPU tests of the reference-facing surface: CLI with the exact `./run` flags, weight formats, sampling,
option validation.  Run with `pytest -m gpu` on a B200."""
import json
import os
import wave

import numpy as np
import pytest
import torch

from tools import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def micro_dir(tmp_path_factory, built_lib):
    return synth.write_model(str(tmp_path_factory.mktemp("micro")), "micro", 0)


def _write_wav(path, x):
    pcm = np.clip(np.round(x * 32768.0), -32768, 32767).astype(np.int16)
    with wave.open(path, "wb") as w:
        w.setnchannels(1), w.setsampwidth(2), w.setframerate(16000)
        w.writeframes(pcm.tobytes())
    return pcm.astype(np.float32) / 32768.0


def test_cli_with_the_reference_flags(micro_dir, tmp_path, monkeypatch):
    """`./run in out` == cli input -f txt --output-name out --model M --condition-on-previous-text False
    --hallucination-silence-threshold 1 (/root/reference/run:3-6): writes out.txt, one line per segment."""
    from whisper_mlx_b200 import transcribe
    from whisper_mlx_b200.cli import main
    from whisper_mlx_b200.load_models import load_model

    wav = str(tmp_path / "in.wav")
    x = _write_wav(wav, synth.long_audio(40.0, 5))
    monkeypatch.chdir(tmp_path)
    main([wav, "-f", "txt", "--output-name", "out", "--model", micro_dir, "--condition-on-previous-text", "False",
          "--hallucination-silence-threshold", "1", "--verbose", "False", "--temperature-increment-on-fallback", "None",
          "--language", "en"])
    lines = open(tmp_path / "out.txt").read().splitlines()
    trace = []
    m = load_model(micro_dir)
    ref = transcribe(x, model=m, condition_on_previous_text=False, temperature=0.0, language="en", window_trace=trace)
    assert lines == [s["text"].strip() for s in ref["segments"]] and len(lines) >= 1
    # ... and that transcript is the reference control flow on tokens that follow the oracle model (not just self-equal)
    from oracle import model as OM
    from tests.gpu_cases import _replay_and_validate

    cfg, w = synth.load_weights_f32(micro_dir)
    kw = dict(temperature=0.0, condition_on_previous_text=False, language="en")
    n_seg, n_win, near = _replay_and_validate("cli", False, ref, trace, w, OM.ModelDimensions(**cfg), x, m, kw, 224)
    assert n_seg == len(lines)


def test_cli_renders_text_with_the_vocabulary_in_the_model_directory(micro_dir, tmp_path, monkeypatch):
    """`multilingual.tiktoken` next to the weights: the CLI's text goes through the tiktoken branch end to end (no
    surrogate allowed), and equals the BPE decoding of the segment tokens."""
    import shutil

    from tools import synth_vocab
    from whisper_mlx_b200 import transcribe
    from whisper_mlx_b200.cli import main
    from whisper_mlx_b200.load_models import load_model
    from whisper_mlx_b200.tokenizer import get_tokenizer

    mdir = str(tmp_path / "model")
    shutil.copytree(micro_dir, mdir)
    synth_vocab.write_vocab(mdir, "multilingual")
    monkeypatch.delenv("B200W_ALLOW_SURROGATE", raising=False)
    wav = str(tmp_path / "in.wav")
    x = _write_wav(wav, synth.long_audio(35.0, 9))
    monkeypatch.chdir(tmp_path)
    main([wav, "-f", "txt", "--output-name", "out", "--model", mdir, "--condition-on-previous-text", "False",
          "--verbose", "False", "--temperature-increment-on-fallback", "None", "--language", "en"])
    written = open(tmp_path / "out.txt", newline="").read()  # (random tokens decode to bytes that include line breaks)
    r = transcribe(x, model=load_model(mdir), condition_on_previous_text=False, temperature=0.0, language="en")
    tk = get_tokenizer(True, num_languages=99, language="en", task="transcribe", vocab_dir=mdir)
    assert tk.encoding.has_vocab and len(r["segments"]) >= 1
    assert written == "".join(tk.decode([t for t in seg["tokens"] if t < tk.eot]).strip() + "\n" for seg in r["segments"])
    # without the vocabulary (and without the test override) the CLI reports the file instead of writing made-up text
    os.remove(os.path.join(mdir, "multilingual.tiktoken"))
    os.remove(tmp_path / "out.txt")
    from whisper_mlx_b200.transcribe import ModelHolder

    ModelHolder.model = None
    main([wav, "-f", "txt", "--output-name", "out", "--model", mdir, "--verbose", "False", "--language", "en"])
    assert not os.path.exists(tmp_path / "out.txt")


def test_npz_and_fp16_weights_load_identically(micro_dir, tmp_path):
    """The reference's other on-disk formats: weights.npz, fp16 storage."""
    from safetensors.torch import load_file
    from whisper_mlx_b200.load_models import load_model

    w = load_file(os.path.join(micro_dir, "weights.safetensors"))
    d2 = tmp_path / "npz"
    d2.mkdir()
    np.savez(str(d2 / "weights.npz"), **{k: v.float().numpy().astype(np.float16) for k, v in w.items()})
    cfg = json.load(open(os.path.join(micro_dir, "config.json")))
    json.dump(cfg, open(d2 / "config.json", "w"))
    m1, m2 = load_model(micro_dir), load_model(str(d2))
    mel = torch.randn(1, 3000, 80)
    a, b = m1.embed_audio(mel).float(), m2.embed_audio(mel).float()
    # bf16 weights are exactly representable in... not in fp16 in general: allow the fp16->bf16 re-rounding
    assert (a - b).abs().max().item() <= 0.1 and (a - b).abs().mean().item() <= 5e-3


def test_sampling_and_best_of(micro_dir):
    """temperature > 0: Gumbel-max sampling on the device, best_of groups share the cross K/V slot, the ranker picks
    the highest mean log-probability; seeded runs are reproducible and different seeds differ."""
    from whisper_mlx_b200.decoding import DecodingOptions, DecodingTask
    from whisper_mlx_b200.load_models import load_model

    m = load_model(micro_dir)
    g = torch.Generator().manual_seed(0)
    xa = torch.randn(3, 1500, 128, generator=g).bfloat16().cuda()
    greedy = DecodingTask(m, DecodingOptions(language="en", sample_len=16)).run_features(xa)
    s1 = DecodingTask(m, DecodingOptions(language="en", sample_len=16, temperature=1.0, best_of=4, seed=1)).run_features(xa)
    s1b = DecodingTask(m, DecodingOptions(language="en", sample_len=16, temperature=1.0, best_of=4, seed=1)).run_features(xa)
    s2 = DecodingTask(m, DecodingOptions(language="en", sample_len=16, temperature=1.0, best_of=4, seed=2)).run_features(xa)
    assert [r.tokens for r in s1] == [r.tokens for r in s1b]
    assert [r.tokens for r in s1] != [r.tokens for r in s2]
    assert any(a.tokens != b.tokens for a, b in zip(s1, greedy))
    tb = 50364
    for r in s1 + s2:
        assert r.temperature == 1.0 and np.isfinite(r.avg_logprob) and r.avg_logprob < 0
        assert len(r.tokens) >= 1 and tb <= r.tokens[0] <= tb + 50  # the timestamp grammar holds under sampling too
    # near-zero temperature sampling reproduces greedy decoding
    cold = DecodingTask(m, DecodingOptions(language="en", sample_len=16, temperature=1e-4, best_of=1, seed=3)).run_features(xa)
    assert [r.tokens for r in cold] == [r.tokens for r in greedy]


def test_sampling_follows_the_oracle_distribution(micro_dir):
    """temperature 1 on the device against the oracle's categorical: 512 first-token samples of one window match
    softmax(filtered logits) within 5 sigma per token, every sampled token is one the rules allow, and the reported
    avg_logprob is the oracle's sum of log-probabilities of the sampled tokens."""
    from oracle import decoding as OD, model as OM
    from oracle.tokens import TokenIds
    from tests.gpu_cases import _check_greedy_trajectory, _product_logits_fn
    from whisper_mlx_b200.decoding import DecodeSession, DecodingOptions, DecodingTask
    from whisper_mlx_b200.load_models import load_model

    m = load_model(micro_dir)
    cfg, w = synth.load_weights_f32(micro_dir)
    dims = OM.ModelDimensions(**cfg)
    ids = TokenIds(dims.n_vocab)
    g = torch.Generator().manual_seed(5)
    xa = torch.randn(1, 1500, 128, generator=g).bfloat16()
    n = 512
    task = DecodingTask(m, DecodingOptions(language="en", sample_len=8, temperature=1.0, best_of=n, seed=11))
    sess = DecodeSession(m, xa.cuda(), n, max_tokens=16)
    n0 = len(task.initial_tokens)
    sess.set_tokens(torch.tensor(task.initial_tokens, dtype=torch.int32).repeat(n, 1))
    sess.set_filter(task._filter_params(sess), task._get_suppress_tokens())
    sess.prompt_step(n0, task.sot_index)
    torch.cuda.synchronize()
    first = sess.tokens[:, n0].cpu().numpy()
    toks = torch.tensor([list(task.initial_tokens)], dtype=torch.long)
    ref, _ = OM.decoder_forward(w, dims, toks, xa.float(), policy="bf16")
    row = ref[0, -1:].float().numpy().copy()
    OD.filter_logits(row, toks.numpy(), n0, ids, ids.suppress_set())
    p = torch.softmax(torch.from_numpy(row[0]), -1).numpy()
    assert np.all(p[first] > 0), "a forbidden token was sampled"
    counts = np.bincount(first, minlength=dims.n_vocab)
    checked = 0
    for t in np.argsort(-p)[:12]:
        if p[t] < 0.02:
            break
        sigma = np.sqrt(n * p[t] * (1 - p[t]))
        assert abs(counts[t] - n * p[t]) <= 5 * sigma + 1, (int(t), int(counts[t]), n * p[t])
        checked += 1
    assert checked >= 3 and len(set(first.tolist())) >= 5
    # whole sampled sequences: allowed tokens + log-probability bookkeeping (best_of picks by mean log-probability)
    res = DecodingTask(m, DecodingOptions(language="en", sample_len=10, temperature=0.7, best_of=3, seed=4)).run_features(xa.cuda())[0]
    chk = _check_greedy_trajectory(w, dims, xa[0].float(), _product_logits_fn(m, xa.cuda()), res.tokens, 10,
                                   avg_logprob=res.avg_logprob, greedy=False)
    assert chk["n"] >= 1


def test_temperature_fallback_in_transcribe(micro_dir):
    """Random weights give avg_logprob << -1, so every temperature of the ladder is tried; the result carries the
    last temperature, like the reference's decode_with_fallback."""
    from whisper_mlx_b200 import transcribe
    from whisper_mlx_b200.load_models import load_model

    m = load_model(micro_dir)
    r = transcribe(synth.white_noise(16000 * 20, 1), model=m, temperature=(0.0, 0.4, 0.8), best_of=2, language="en",
                   condition_on_previous_text=False, sample_len=8)
    assert len(r["segments"]) >= 1 and all(s["temperature"] == 0.8 for s in r["segments"])
    r0 = transcribe(synth.white_noise(16000 * 20, 1), model=m, temperature=(0.0, 0.4, 0.8), best_of=2, language="en",
                    condition_on_previous_text=False, sample_len=8, logprob_threshold=None, compression_ratio_threshold=None)
    assert all(s["temperature"] == 0.0 for s in r0["segments"])


def test_option_errors(micro_dir):
    from whisper_mlx_b200 import transcribe
    from whisper_mlx_b200.decoding import DecodingOptions, decode
    from whisper_mlx_b200.load_models import load_model

    m = load_model(micro_dir)
    with pytest.raises(NotImplementedError):
        decode(m, torch.zeros(3000, 80), DecodingOptions(language="en", beam_size=5))
    with pytest.raises(AssertionError, match="incorrect audio shape"):
        m.embed_audio(torch.zeros(1, 2999, 80))
    r = transcribe(synth.white_noise(16000, 0), model=m, word_timestamps=True, language="en", temperature=0.0, sample_len=8)
    assert all("words" in s for s in r["segments"])
    with pytest.raises(ValueError):
        transcribe(synth.white_noise(16000 * 40, 0), model=m, world_size=2, rank=0)  # sharding needs the fixed-window mode


def test_prompt_conditioning_runs(micro_dir):
    """condition_on_previous_text=True (the reference default): windows carry the previous tokens as a prompt."""
    from whisper_mlx_b200 import transcribe
    from whisper_mlx_b200.load_models import load_model

    m = load_model(micro_dir)
    r = transcribe(synth.long_audio(65.0, 2), model=m, temperature=0.0, language="en", sample_len=12,
                   logprob_threshold=None, compression_ratio_threshold=None)
    assert len(r["segments"]) >= 2 and r["segments"][-1]["seek"] > 0


def test_daemon_tool_transcribes(micro_dir, tmp_path):
    """The reference's plugin API (Tool.execute(**arguments) -> JSON string) over the transcription path."""
    from whisper_mlx_b200 import transcribe
    from whisper_mlx_b200.load_models import load_model
    from whisper_mlx_b200.tool import TOOL

    wav = str(tmp_path / "in.wav")
    x = _write_wav(wav, synth.long_audio(40.0, 5))
    r = json.loads(TOOL.execute(file_path=wav, language="en", model=micro_dir))
    assert r["status"] == "success", r
    ref = transcribe(x, model=load_model(micro_dir), condition_on_previous_text=False, language="en")
    assert r["text"] == ref["text"].strip() and r["segment_count"] == len(ref["segments"]) and r["language"] == "en"
    assert [s["text"] for s in r["segments"]] == [s["text"].strip() for s in ref["segments"]]


def test_mlx_quantised_checkpoint_loads(micro_dir, tmp_path):
    """A 4-bit MLX-quantised checkpoint (`quantization` in config.json, uint32 weight + scales + biases per Linear /
    Embedding) runs and equals the model built from the same weights expanded by hand."""
    from safetensors.torch import load_file, save_file
    from tests.test_host_logic import _mlx_quantize
    from whisper_mlx_b200.load_models import load_model

    w = load_file(os.path.join(micro_dir, "weights.safetensors"))
    cfg = json.load(open(os.path.join(micro_dir, "config.json")))
    qdir, ddir = tmp_path / "q4", tmp_path / "dense"
    qdir.mkdir(), ddir.mkdir()
    wq, wd = {}, {}
    for k, v in w.items():
        quantise = k.endswith(".weight") and v.ndim == 2 and (".attn." in k or ".cross_attn." in k or ".mlp1." in k or ".mlp2." in k or k == "decoder.token_embedding.weight")
        if quantise:
            words, scales, biases, dense = _mlx_quantize(v.float(), 64, 4)
            base = k[: -len(".weight")]
            wq[k], wq[base + ".scales"], wq[base + ".biases"] = words.view(torch.int32), scales.half(), biases.half()
            wd[k] = (_dequant_like(words, scales.half(), biases.half())).to(torch.bfloat16)
        else:
            wq[k], wd[k] = v.contiguous(), v.contiguous()
    save_file(wq, str(qdir / "weights.safetensors"))
    save_file(wd, str(ddir / "weights.safetensors"))
    json.dump({**cfg, "quantization": {"group_size": 64, "bits": 4}}, open(qdir / "config.json", "w"))
    json.dump(cfg, open(ddir / "config.json", "w"))
    mq, md = load_model(str(qdir)), load_model(str(ddir))
    mel = torch.randn(1, 3000, 80)
    assert torch.equal(mq.embed_audio(mel), md.embed_audio(mel))


@pytest.mark.parametrize("fmt,quant", [("npz", None), ("safetensors", dict(group_size=64, bits=4))])
def test_checkpoint_written_from_the_hf_side(tmp_path, built_lib, fmt, quant):
    """The loader + engine on an `mlx-community/whisper-*-mlx`-shaped directory produced from a transformers model by the
    published HF -> MLX conversion rules (tests/hf_bridge.py), i.e. NOT by this repo's own writer: logits must equal the
    transformers forward pass (fp32) within the bf16 tolerance (4-bit: within the quantisation error, measured on
    the oracle: see tests/test_assets.py)."""
    from oracle import audio as OA
    from tests.test_assets import _hf_logits, _hf_model, write_mlx_checkpoint
    from whisper_mlx_b200.load_models import load_model

    hf = _hf_model()
    m = load_model(write_mlx_checkpoint(str(tmp_path / "ckpt"), hf, fmt, quant))
    assert m.alignment_heads.tolist() == [[1, 0], [1, 1]]
    mel = torch.from_numpy(OA.log_mel_spectrogram(synth.make_audio("speech", 480000, 5), 80))[None]
    tokens = torch.tensor([[50258, 50259, 50359, 50364, 400, 500, 50400]])
    ref = _hf_logits(hf, mel, tokens)
    got = m.logits(tokens, m.embed_audio(mel)).cpu()
    err = (got - ref).abs().max().item()
    scale = max(1.0, ref.abs().max().item() / 10)
    assert err <= (8e-2 if quant is None else 0.6) * scale, (fmt, quant, err)
    if quant is None:
        assert bool((got.argmax(-1) == ref.argmax(-1)).float().mean() >= 0.7)


def _dequant_like(words, scales, biases):
    from whisper_mlx_b200.load_models import dequantize

    return dequantize(words, scales, biases, 64, 4)


def test_finished_sequences_are_skipped_without_touching_the_others(micro_dir):
    """Sequences that have emitted EOT stop streaming their K/V in the attention kernels (decoder_step passes the
    `finished` flags while sampling): their tokens stay EOT and every other sequence decodes exactly as before."""
    from whisper_mlx_b200.decoding import DecodeSession, DecodingOptions, DecodingTask
    from whisper_mlx_b200.load_models import load_model

    m = load_model(micro_dir)
    g = torch.Generator().manual_seed(3)
    xa = torch.randn(4, 1500, 128, generator=g).bfloat16().cuda()
    task = DecodingTask(m, DecodingOptions(language="en", sample_len=12))
    eot, n0 = task.tokenizer.eot, len(task.initial_tokens)

    def run(stop):
        sess = DecodeSession(m, xa, 1, max_tokens=n0 + 12)
        sess.set_tokens(torch.tensor(task.initial_tokens, dtype=torch.int32).repeat(4, 1))
        sess.set_filter(task._filter_params(sess), task._get_suppress_tokens())
        sess.prompt_step(n0, task.sot_index)
        for b in stop:  # pretend the first sampled token of these sequences was EOT
            sess.tokens[b, n0] = eot
            sess.finished[b] = 1
        for _ in range(10):
            sess.sample_step()
        torch.cuda.synchronize()
        return sess.tokens[:, : n0 + 11].cpu()

    full, part = run([]), run([1, 3])
    assert torch.equal(full[0], part[0]) and torch.equal(full[2], part[2])
    assert bool((part[1, n0:] == eot).all()) and bool((part[3, n0:] == eot).all())
    assert not bool((full[1, n0:] == eot).all())


def test_empty_and_very_short_audio(micro_dir):
    """Zero samples and a few milliseconds of audio transcribe without error (the 30 s zero padding makes the signal
    long enough for the reflect pad); log-mel of a too-short unpadded signal is rejected with a clear message."""
    from whisper_mlx_b200 import transcribe
    from whisper_mlx_b200.audio import log_mel_spectrogram
    from whisper_mlx_b200.load_models import load_model

    m = load_model(micro_dir)
    for x in (np.zeros(0, dtype=np.float32), synth.white_noise(100, 0), np.zeros(0, dtype=np.int16)):
        r = transcribe(x, model=m, language="en", temperature=0.0, sample_len=8)
        assert set(r) == {"text", "segments", "language"} and isinstance(r["segments"], list)
    mel = log_mel_spectrogram(np.zeros(0, dtype=np.float32), n_mels=80, padding=480000)
    assert mel.shape == (3000, 80) and bool(torch.isfinite(mel).all())
    with pytest.raises(ValueError, match="reflect pad"):
        log_mel_spectrogram(synth.white_noise(150, 0), n_mels=80)


def test_decode_chain_is_bit_identical_to_separate_launches():
    """K11 (the decode chain) runs the same arithmetic in the same order as the per-phase launches it replaces:
    tokens and logits of a 64-sequence whisper-small decode must match bit for bit (tools/ab_chain.py runs the
    B200W_CHAIN=0 / =1 halves in separate processes)."""
    import subprocess
    import sys

    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(repo, "tools", "ab_chain.py"), "small", "64", "12"], capture_output=True,
                         text=True, timeout=600, cwd=repo)
    assert out.returncode == 0 and "IDENTICAL" in out.stdout, out.stdout[-1500:] + out.stderr[-1500:]


def test_decode_chain_bit_identical_at_the_headline_shape():
    """The same A/B at BASELINE config 4's own shape: large-v3, 120 sequences, 24 graph-replayed steps."""
    import subprocess
    import sys

    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(repo, "tools", "ab_chain.py"), "large-v3", "120", "24"], capture_output=True,
                         text=True, timeout=900, cwd=repo)
    assert out.returncode == 0 and "IDENTICAL" in out.stdout, out.stdout[-1500:] + out.stderr[-1500:]


def test_cli_word_timestamps_json(micro_dir, tmp_path, monkeypatch):
    """`--word-timestamps True -f json`: every segment of the written JSON carries its words."""
    from whisper_mlx_b200.cli import main

    wav = str(tmp_path / "in.wav")
    _write_wav(wav, synth.long_audio(35.0, 6))
    monkeypatch.chdir(tmp_path)
    main([wav, "-f", "json", "--output-name", "out", "--model", micro_dir, "--condition-on-previous-text", "False",
          "--word-timestamps", "True", "--verbose", "False", "--temperature-increment-on-fallback", "None", "--language", "en"])
    r = json.load(open(tmp_path / "out.json"))
    assert len(r["segments"]) >= 1 and all("words" in s for s in r["segments"])
    words = [w for s in r["segments"] for w in s["words"]]
    assert len(words) >= 1 and all({"word", "start", "end", "probability"} <= set(w) for w in words)


def test_several_files_in_lockstep(micro_dir, tmp_path, monkeypatch):
    """`cli a.wav b.wav c.wav` in the exact mode decodes the three files' windows in shared batches (transcribe_many):
    on one kernel path the result of every file is exactly its file-by-file result, and with the default kernel choice
    (K13 for one window, K13m for three) every decoded window still follows the oracle model."""
    from whisper_mlx_b200 import transcribe
    from whisper_mlx_b200.cli import main
    from whisper_mlx_b200.load_models import load_model
    from whisper_mlx_b200.transcribe import transcribe_many

    m = load_model(micro_dir)
    xs = [synth.long_audio(t, seed) for t, seed in ((50.0, 21), (33.0, 22), (75.0, 23))]
    kw = dict(model=m, condition_on_previous_text=False, temperature=0.0, language="en")
    key = lambda r: [(s["seek"], s["start"], s["end"], tuple(s["tokens"])) for s in r["segments"]]  # noqa: E731
    for name in ("B200W_SMALL", "B200W_SMALL_MMA"):
        monkeypatch.setenv(name, "0")  # the chain path at every batch size: rows do not depend on their batch
    alone = [transcribe(x, **kw) for x in xs]
    traces = [[] for _ in xs]
    many = transcribe_many(xs, **kw)
    assert [key(r) for r in many] == [key(r) for r in alone] and all(len(r["segments"]) >= 1 for r in many)
    for name in ("B200W_SMALL", "B200W_SMALL_MMA"):
        monkeypatch.delenv(name)
    # default kernels: the CLI on three files, every file's text written
    wavs = []
    for i, x in enumerate(xs):
        wavs.append(str(tmp_path / f"f{i}.wav"))
        _write_wav(wavs[-1], x)
    monkeypatch.chdir(tmp_path)
    main([*wavs, "-f", "txt", "--model", micro_dir, "--condition-on-previous-text", "False", "--verbose", "False",
          "--temperature-increment-on-fallback", "None", "--language", "en"])
    for i in range(3):
        lines = open(tmp_path / f"f{i}.txt").read().splitlines()
        assert len(lines) >= 1
    del traces
