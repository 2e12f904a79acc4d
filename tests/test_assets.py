"""On-disk / wire formats either side of the path, exercised WITHOUT the real assets (no network in this image):

* the `*.tiktoken` vocabulary branch of the tokenizer, on a synthetic vocabulary of the real size and format
  (tools/synth_vocab.py); and the refusal to render text without a vocabulary;
* the ffmpeg branch of `load_audio`, with a stand-in `ffmpeg` executable on PATH;
* the MLX checkpoint layout (`config.json` + `weights.npz` / `weights.safetensors`, fp16, optionally group-quantised)
  written from an INDEPENDENT description -- a `transformers` Whisper state dict pushed through the published
  HF -> MLX conversion rules (tests/hf_bridge.py::hf_to_mlx) -- not from the loader's own inverse.
"""
import json
import os
import stat
import sys
import wave

import numpy as np
import pytest
import torch

from tools import synth, synth_vocab


@pytest.fixture(scope="module")
def vocab_dir(tmp_path_factory):
    d = str(tmp_path_factory.mktemp("vocab"))
    synth_vocab.write_vocab(d, "multilingual")
    synth_vocab.write_vocab(d, "gpt2")
    return d


# ------------------------------------------------------------------------------------------ tokenizer
def test_tokenizer_refuses_to_invent_text(monkeypatch, tmp_path):
    from whisper_mlx_b200 import tokenizer as T

    monkeypatch.delenv("B200W_ALLOW_SURROGATE", raising=False)
    monkeypatch.delenv("B200W_TIKTOKEN_DIR", raising=False)
    with pytest.raises(FileNotFoundError, match="multilingual.tiktoken"):
        T.get_tokenizer(True, num_languages=100, language="en", task="transcribe")
    with pytest.raises(FileNotFoundError, match="gpt2.tiktoken"):
        T.get_tokenizer(False, vocab_dir=str(tmp_path))
    monkeypatch.setenv("B200W_ALLOW_SURROGATE", "1")
    tk = T.get_tokenizer(True, num_languages=100, language="en", task="transcribe")
    assert not tk.encoding.has_vocab
    with pytest.raises(RuntimeError, match="BPE vocabulary"):
        tk.encode("hello")


def test_tiktoken_branch_roundtrip(vocab_dir, monkeypatch):
    from whisper_mlx_b200 import tokenizer as T

    monkeypatch.delenv("B200W_ALLOW_SURROGATE", raising=False)
    for n_lang, n_vocab in ((99, 51865), (100, 51866)):
        tk = T.get_tokenizer(True, num_languages=n_lang, language="de", task="translate", vocab_dir=vocab_dir)
        assert tk.encoding.has_vocab and tk.encoding.n_vocab == n_vocab
        # the special-token table sits right after the 50257 ranks, whatever the vocabulary is
        assert (tk.eot, tk.sot, tk.timestamp_begin) == (50257, 50258, n_vocab - 1501)
        assert tk.sot_sequence == (50258, 50259 + 2, tk.translate)
        assert tk.encode(" ") == [220]
        for text in (" the quick brown fox", "Hello, world! 123", " naïve café — ♪♪ 「日本語」", "it's (a) test--ok", ""):
            ids = tk.encode(text)
            assert all(0 <= t < 50257 for t in ids)
            assert tk.decode(ids) == text
        ids = tk.encode(" hello world")
        assert len(ids) < len(" hello world")  # merges are applied, not just bytes
        # decode() drops timestamps, decode_with_timestamps() renders them, specials render by name
        mixed = [tk.timestamp_begin + 54, *ids, tk.timestamp_begin + 100]
        assert tk.decode(mixed) == " hello world"
        assert tk.decode_with_timestamps(mixed) == "<|1.08|> hello world<|2.00|>"
        assert tk.decode_with_timestamps([tk.sot, tk.no_speech]) == "<|startoftranscript|><|nospeech|>"
        # non_speech_tokens is DERIVED from the vocabulary on this branch (not the hard-coded table)
        ns = tk.non_speech_tokens
        assert tk.encode(" -")[0] in ns and tk.encode(" '")[0] in ns and tk.encode("♪")[0] in ns
        assert tk.encode("(")[0] in ns and tk.encode(" hello")[0] not in ns and tuple(sorted(ns)) == ns
    en = T.get_tokenizer(False, vocab_dir=vocab_dir)
    assert en.encoding.n_vocab == 51864 and en.eot == 50256 and en.sot_sequence == (50257,)
    assert en.decode(en.encode(" plain english")) == " plain english" and len(en.non_speech_tokens) > 10


def test_tiktoken_branch_word_splitting(vocab_dir):
    from whisper_mlx_b200 import tokenizer as T

    tk = T.get_tokenizer(True, num_languages=100, language="en", task="transcribe", vocab_dir=vocab_dir)
    text = " Hello world, it's naïve."
    ids = tk.encode(text) + [tk.eot]
    words, word_tokens = tk.split_to_word_tokens(ids)
    assert "".join(words[:-1]) == text and [t for wt in word_tokens for t in wt] == ids
    assert words[:2] == [" Hello", " world"] and words[-1] == "<|endoftext|>"
    assert "," in words and words[-2] == "."  # punctuation splits off
    # multi-byte characters cut across tokens are only emitted once complete
    ja = T.get_tokenizer(True, num_languages=100, language="ja", task="transcribe", vocab_dir=vocab_dir)
    ids = ja.encode("日本語のテスト")
    words, word_tokens = ja.split_to_word_tokens(ids)
    assert "".join(words) == "日本語のテスト" and all("�" not in w for w in words)
    assert [t for wt in word_tokens for t in wt] == ids


def test_decoding_task_uses_the_model_directory_vocabulary(vocab_dir, monkeypatch):
    """one tokenizer end to end -- a vocabulary that lives only in the model directory must also be the
    one DecodingTask renders DecodingResult.text (and the compression ratio) with."""
    from whisper_mlx_b200.decoding import DecodingOptions, DecodingTask

    monkeypatch.delenv("B200W_ALLOW_SURROGATE", raising=False)

    class FakeModel:
        is_multilingual, num_languages, model_path = True, 100, vocab_dir

        class dims:
            n_text_ctx, n_vocab, n_audio_ctx = 448, 51866, 1500

    task = DecodingTask(FakeModel(), DecodingOptions(language="en", prompt=" previous text", prefix="So"))
    assert task.tokenizer.encoding.has_vocab
    toks, tk = list(task.initial_tokens), task.tokenizer
    i = toks.index(tk.sot)
    assert toks[0] == tk.sot_prev and tk.decode(toks[1:i]) == " previous text" and task.sot_index == i
    assert tuple(toks[i: i + 3]) == tk.sot_sequence and tk.decode(toks[i + 3:]) == " So" and task.sample_begin == len(toks)
    FakeModel.model_path = None
    with pytest.raises(FileNotFoundError):
        DecodingTask(FakeModel(), DecodingOptions(language="en"))


# ------------------------------------------------------------------------------------------ ffmpeg
_FFMPEG_STUB = r'''#!PYTHON
"""Stand-in for the ffmpeg CLI (tests): decodes a WAV input to what the reference asks ffmpeg for."""
import sys, wave
import numpy as np
a = sys.argv[1:]
assert a[0] == "-nostdin" and "-i" in a and a[-1] == "-", a
src = a[a.index("-i") + 1]
assert a[a.index("-f") + 1] == "s16le" and a[a.index("-ac") + 1] == "1" and a[a.index("-acodec") + 1] == "pcm_s16le"
sr = int(a[a.index("-ar") + 1])
try:
    w = wave.open(src, "rb")
except Exception as e:
    sys.stderr.write(f"{src}: Invalid data found when processing input\n")
    sys.exit(1)
x = np.frombuffer(w.readframes(w.getnframes()), dtype=np.int16).reshape(-1, w.getnchannels()).astype(np.float64).mean(1)
if w.getframerate() != sr:  # linear resampling is enough for a stand-in
    n = int(round(len(x) * sr / w.getframerate()))
    x = np.interp(np.arange(n) * w.getframerate() / sr, np.arange(len(x)), x)
sys.stdout.buffer.write(np.round(x).astype("<i2").tobytes())
'''


def test_load_audio_ffmpeg_branch(tmp_path, monkeypatch):
    from whisper_mlx_b200.audio import load_audio

    bindir = tmp_path / "bin"
    bindir.mkdir()
    exe = bindir / "ffmpeg"
    exe.write_text(_FFMPEG_STUB.replace("PYTHON", sys.executable))
    exe.chmod(exe.stat().st_mode | stat.S_IEXEC)
    monkeypatch.setenv("PATH", f"{bindir}{os.pathsep}{os.environ['PATH']}")
    # stereo 8 kHz input: only ffmpeg can turn this into mono 16 kHz (the WAV fallback refuses it)
    pcm = (np.stack([synth.white_noise(800, 0), synth.white_noise(800, 1)], 1) * 20000).astype(np.int16)
    p = str(tmp_path / "stereo8k.wav")
    with wave.open(p, "wb") as w:
        w.setnchannels(2), w.setsampwidth(2), w.setframerate(8000)
        w.writeframes(pcm.tobytes())
    x = load_audio(p)
    assert x.dtype == np.float32 and x.shape == (1600,) and np.abs(x).max() <= 1.0
    mono = pcm.astype(np.float64).mean(1)
    assert np.allclose(x[::2] * 32768.0, np.round(mono), atol=1.0)
    # ffmpeg's failure is reported the way the reference reports it
    bad = tmp_path / "not_audio.mp3"
    bad.write_bytes(b"\x00" * 64)
    with pytest.raises(RuntimeError, match="Failed to load audio: .*Invalid data found"):
        load_audio(str(bad))


# ------------------------------------------------------------------------------------------ checkpoints
_TINY2 = dict(n_mels=80, n_audio_ctx=1500, n_audio_state=128, n_audio_head=2, n_audio_layer=2, n_vocab=51865, n_text_ctx=448,
              n_text_state=128, n_text_head=2, n_text_layer=2)


def _hf_model(seed=0):
    from transformers import WhisperConfig, WhisperForConditionalGeneration

    d = _TINY2
    torch.manual_seed(seed)
    cfg = WhisperConfig(vocab_size=d["n_vocab"], num_mel_bins=d["n_mels"], d_model=d["n_audio_state"],
                        encoder_layers=d["n_audio_layer"], encoder_attention_heads=d["n_audio_head"],
                        decoder_layers=d["n_text_layer"], decoder_attention_heads=d["n_text_head"],
                        encoder_ffn_dim=4 * d["n_audio_state"], decoder_ffn_dim=4 * d["n_text_state"],
                        max_source_positions=1500, max_target_positions=448, activation_function="gelu", dropout=0.0,
                        attention_dropout=0.0, activation_dropout=0.0, tie_word_embeddings=True)
    cfg._attn_implementation = "eager"
    hf = WhisperForConditionalGeneration(cfg).eval().float()
    with torch.no_grad():  # HF initialises biases to zero and norms to one: make every tensor informative
        for n, p in hf.named_parameters():
            if p.ndim == 1:
                p.add_(0.1 * torch.randn_like(p))
    return hf


def write_mlx_checkpoint(path, hf, fmt="npz", quantization=None):
    """An `mlx-community/whisper-*-mlx`-shaped directory: config.json (+ model_type, + quantization) and fp16 weights."""
    from tests.hf_bridge import hf_to_mlx, mlx_quantize_dict

    os.makedirs(path, exist_ok=True)
    w = hf_to_mlx(hf.state_dict(), torch.float16)
    cfg = dict(_TINY2, model_type="whisper")
    if quantization:
        w = mlx_quantize_dict(w, **quantization)
        cfg["quantization"] = dict(quantization)
    w["alignment_heads"] = torch.tensor([[1, 0], [1, 1]], dtype=torch.int32)
    json.dump(cfg, open(os.path.join(path, "config.json"), "w"))
    if fmt == "npz":
        np.savez(os.path.join(path, "weights.npz"), **{k: v.numpy() for k, v in w.items()})
    else:
        from safetensors.torch import save_file

        save_file({k: v.contiguous() for k, v in w.items()}, os.path.join(path, "weights.safetensors"))
    return path


def _hf_logits(hf, mel, tokens):
    with torch.no_grad():
        return hf(input_features=mel.transpose(1, 2), decoder_input_ids=tokens).logits


@pytest.mark.parametrize("fmt,quant", [("npz", None), ("safetensors", None), ("safetensors", dict(group_size=64, bits=4)),
                                       ("npz", dict(group_size=32, bits=8))])
def test_mlx_checkpoint_layout_read_back(tmp_path, fmt, quant):
    """Loader (host half) on a checkpoint written from the HF side: names and shapes are exactly the MLX set, and the
    oracle run on what the loader read reproduces the HF model's logits (quantised: within the quantisation error)."""
    from oracle import audio as OA, model as OM
    from whisper_mlx_b200.load_models import read_model_files

    hf = _hf_model()
    d = write_mlx_checkpoint(str(tmp_path / "ckpt"), hf, fmt, quant)
    dims, w, heads = read_model_files(d)
    assert heads.tolist() == [[1, 0], [1, 1]]
    shapes = synth.weight_shapes(_TINY2)
    assert set(w) == set(shapes), (set(w) ^ set(shapes))
    for k, shp in shapes.items():
        assert tuple(w[k].shape) == tuple(shp), k
    assert w["encoder.conv1.weight"].shape == (128, 3, 80)  # MLX (out, k, in), not torch (out, in, k)
    mel = torch.from_numpy(OA.log_mel_spectrogram(synth.make_audio("speech", 480000, 5), 80))[None]
    tokens = torch.tensor([[50258, 50259, 50359, 50364, 400, 500, 50400]])
    ref = _hf_logits(hf, mel, tokens)
    odims = OM.ModelDimensions(**_TINY2)
    w32 = {k: v.float() for k, v in w.items()}
    xa = OM.encoder_forward(w32, odims, mel, policy="fp32")
    got, _ = OM.decoder_forward(w32, odims, tokens, xa, policy="fp32")
    err = (got - ref).abs().max().item()
    tol = 2e-2 if quant is None else (0.5 if quant["bits"] == 4 else 6e-2)  # fp16 storage / 4-bit / 8-bit weights
    assert err <= tol * max(1.0, ref.abs().max().item() / 10), (fmt, quant, err)
    if quant is None:
        assert torch.equal(w["decoder.blocks.1.cross_attn.key.weight"], hf.state_dict()["model.decoder.layers.1.encoder_attn.k_proj.weight"].half())
