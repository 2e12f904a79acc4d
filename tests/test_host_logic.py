"""Host-side mirror of the reference interface: tokenizer table, writers, segmentation, CLI flags, options.
CPU only; the segmentation is compared with the oracle's restatement on synthetic token streams."""
import io
import os

import numpy as np
import pytest

from tools import synth


def test_tokenizer_special_ids_match_oracle():
    from oracle.tokens import TokenIds
    from whisper_mlx_b200.tokenizer import get_tokenizer

    for n_lang, n_vocab in ((99, 51865), (100, 51866)):
        tk = get_tokenizer(True, num_languages=n_lang, language="fr", task="transcribe")
        ids = TokenIds(n_vocab)
        assert tk.encoding.n_vocab == n_vocab
        for name in ("eot", "sot", "translate", "transcribe", "sot_lm", "sot_prev", "no_speech", "no_timestamps", "timestamp_begin"):
            assert getattr(tk, name) == getattr(ids, name), name
        assert tk.sot_sequence == ids.sot_sequence("fr")
        assert tuple(sorted(set(tk.non_speech_tokens))) == tuple(t for t in ids.suppress_set() if t < 50257)
        assert len(tk.all_language_tokens) == n_lang and tk.language_code(tk.all_language_tokens[-1]) == ("yue" if n_lang == 100 else "su")
        assert tk.encode(" ") == [220]
        assert tk.decode_with_timestamps([tk.timestamp_begin + 54]) == "<|1.08|>"


def test_surrogate_vocabulary_matches_oracle():
    from oracle.tokens import decode_text
    from whisper_mlx_b200.tokenizer import get_tokenizer

    tk = get_tokenizer(True, num_languages=100)
    toks = [1, 17, 50256, 31337, tk.timestamp_begin + 3, 400]
    assert tk.decode(toks) == decode_text(toks, tk.timestamp_begin)


def test_suppress_set_matches_oracle():
    from oracle.tokens import TokenIds
    from whisper_mlx_b200.decoding import DecodingOptions, DecodingTask

    class FakeModel:
        is_multilingual = True
        num_languages = 100

        class dims:
            n_text_ctx = 448
            n_vocab = 51866
            n_audio_ctx = 1500

    task = DecodingTask(FakeModel(), DecodingOptions(language="en"))
    assert task._get_suppress_tokens() == TokenIds(51866).suppress_set()
    assert task.initial_tokens == TokenIds(51866).sot_sequence("en") and task.sample_begin == 3 and task.sample_len == 224
    t2 = DecodingTask(FakeModel(), DecodingOptions(language="en", prompt=list(range(1000, 1300)), without_timestamps=True))
    assert t2.initial_tokens[0] == 50362 and len(t2.initial_tokens) == 1 + 223 + 4 and t2.sot_index == 224
    with pytest.raises(NotImplementedError):
        DecodingTask(FakeModel(), DecodingOptions(language="en", beam_size=5))
    with pytest.raises(ValueError):
        DecodingTask(FakeModel(), DecodingOptions(language="en", best_of=5))


def _oracle_segments(tokens, seek, size, fixed):
    """Expected segmentation of one window, restated from oracle/transcribe.py's loop body: returns
    ([(start, end, tokens)], advance) in exact mode and ([...], tail or None) in the fixed-window contract."""
    import numpy as _np
    from oracle.tokens import TokenIds

    ids = TokenIds(51866)
    toks = _np.array(tokens, dtype=_np.int64)
    ts = toks >= ids.timestamp_begin
    single_ending = ts[-2:].tolist() == [False, True]
    consecutive = _np.where(_np.logical_and(ts[:-1], ts[1:]))[0] + 1
    segs, adv, tail = [], None, None
    t0 = seek * 160 / 16000
    if len(consecutive) > 0:
        slices = consecutive.tolist()
        if single_ending:
            slices.append(len(toks))
        last = 0
        for cur in slices:
            sl = toks[last:cur]
            segs.append((t0 + (int(sl[0]) - ids.timestamp_begin) * 0.02, t0 + (int(sl[-1]) - ids.timestamp_begin) * 0.02, sl.tolist()))
            last = cur
        adv = size if single_ending else (int(toks[last - 1]) - ids.timestamp_begin) * 2
        if fixed and not single_ending:
            trailing = toks[last:].tolist()
            if any(t < ids.eot for t in trailing):
                start = trailing[0] - ids.timestamp_begin if trailing[0] >= ids.timestamp_begin else adv // 2
                segs.append((t0 + start * 0.02, t0 + size * 0.01, trailing))
            elif 0 < adv < size:
                tail = (seek + adv, size - adv)
    else:
        dur = size * 160 / 16000
        stamps = toks[ts.nonzero()[0]]
        if len(stamps) > 0 and stamps[-1] != ids.timestamp_begin:
            dur = (int(stamps[-1]) - ids.timestamp_begin) * 0.02
        segs.append((t0, t0 + dur, toks.tolist()))
        adv = size
    return segs, (tail if fixed else adv)


@pytest.mark.parametrize("fixed", [False, True])
def test_segmentation_matches_oracle(fixed):
    from whisper_mlx_b200.decoding import DecodingResult
    from whisper_mlx_b200.tokenizer import get_tokenizer
    from whisper_mlx_b200.transcribe import _clear_empty_segments, _segments_fixed_window, _segments_for_window

    tk = get_tokenizer(True, num_languages=100, language="en", task="transcribe")
    tb = tk.timestamp_begin
    streams = [
        [tb, 10, 11, tb + 100, tb + 100, 12, tb + 250, tb + 250, 13, 14],   # unfinished tail text: seek to the last pair / keep it
        [tb, 10, 11, tb + 100, tb + 100, 12, tb + 250, tb + 250],           # closes early, nothing after: re-seek / tail window
        [tb, 10, tb + 1350, tb + 1350, tb + 1400],                          # ... with only an opening timestamp after the pair
        [tb + 5, 10, tb + 1500],                                           # single timestamp ending
        [tb, 10, 11, 12],                                                   # no closing timestamp
        [tb + 20, 7, tb + 60, tb + 60, 8, tb + 90],                         # pair then single ending
        [tb, tb],                                                           # empty pair at 0: no progress possible
        [tb, 10, tb + 1400, tb + 1400, 11],                                 # (1500-frame window) timestamps past the window end
        [tb, 10, tb + 1400, tb + 1400],                                     # ... and nothing after them: no tail to decode
        [10, 11, 12],
        [],
    ]
    for s, size in [(s, size) for s in streams for size in (3000, 1500)]:
        res = DecodingResult(audio_features=None, language="en", tokens=s, avg_logprob=-0.5, no_speech_prob=0.0, temperature=0.0,
                             compression_ratio=1.0)
        toks = np.array(s, dtype=np.int64)
        if fixed:
            segs, got2 = _segments_fixed_window(toks, 3000, size, res, tk, 2, 0.02)
        else:
            segs, got2, _, _ = _segments_for_window(toks, 3000, size, res, tk, 2, 0.02)
        _clear_empty_segments(segs, with_words=False)
        ref, ref2 = _oracle_segments(s, 3000, size, fixed)
        assert got2 == ref2, (s, size, got2, ref2)
        assert len(segs) == len(ref), (s, size)
        for g, (st, en, toks_ref) in zip(segs, ref):
            assert abs(g["start"] - st) < 1e-9 and abs(g["end"] - en) < 1e-9
            if g["start"] != g["end"] and g["text"].strip():
                assert g["tokens"] == toks_ref
            else:
                assert g["tokens"] == [] and g["text"] == ""
    if fixed:
        s = streams[1]
        _, tail = _segments_fixed_window(np.array(s, dtype=np.int64), 3000, 3000,
                                         DecodingResult(audio_features=None, language="en", tokens=s), tk, 2, 0.02)
        assert tail == (3000 + 500, 2500)


def test_writers(tmp_path):
    from whisper_mlx_b200.writers import get_writer

    result = {"text": " a b", "language": "en",
              "segments": [{"id": 0, "start": 0.0, "end": 1.5, "text": " hello there "}, {"id": 1, "start": 3661.25, "end": 3662.0, "text": "x --> y"}]}
    get_writer("txt", str(tmp_path))(result, "out")
    assert open(tmp_path / "out.txt").read() == "hello there\nx --> y\n"
    get_writer("all", str(tmp_path))(result, "all")
    srt = open(tmp_path / "all.srt").read()
    assert "00:00:00,000 --> 00:00:01,500" in srt and "01:01:01,250 --> 01:01:02,000" in srt and "x -> y" in srt
    assert open(tmp_path / "all.vtt").read().startswith("WEBVTT\n")
    assert open(tmp_path / "all.tsv").read().splitlines()[1] == "0\t1500\thello there"
    import json

    assert json.load(open(tmp_path / "all.json"))["segments"][1]["start"] == 3661.25


def test_cli_accepts_the_reference_flags():
    """The exact flag spelling of /root/reference/run:3-6."""
    from whisper_mlx_b200.cli import build_parser

    a = build_parser().parse_args(["in.mp3", "-f", "txt", "--output-name", "out", "--model", "mlx-community/whisper-large-v3-mlx",
                                   "--condition-on-previous-text", "False", "--hallucination-silence-threshold", "1"])
    assert a.audio == ["in.mp3"] and a.output_format == "txt" and a.output_name == "out"
    assert a.model == "mlx-community/whisper-large-v3-mlx" and a.condition_on_previous_text is False
    assert a.hallucination_silence_threshold == 1.0 and a.word_timestamps is False and a.best_of == 5
    assert a.temperature == 0 and a.temperature_increment_on_fallback == 0.2


def test_audio_helpers():
    import torch
    from oracle import audio as OA
    from whisper_mlx_b200 import audio as A

    assert (A.SAMPLE_RATE, A.N_FFT, A.HOP_LENGTH, A.N_SAMPLES, A.N_FRAMES) == (16000, 400, 160, 480000, 3000)
    for n in (80, 128):
        assert np.array_equal(A.mel_filters(n), OA.mel_filters(n))
    with pytest.raises(AssertionError):
        A.mel_filters(64)
    x = np.arange(10, dtype=np.float32)
    assert np.array_equal(A.pad_or_trim(x, 4), x[:4]) and A.pad_or_trim(x, 16).shape == (16,)
    t = torch.arange(12.0).view(3, 4)
    assert A.pad_or_trim(t, 6, axis=0).shape == (6, 4) and A.pad_or_trim(t, 2, axis=-1).shape == (3, 2)
    assert torch.equal(A.pad_or_trim(t, 6, axis=0)[3:], torch.zeros(3, 4))


def test_load_audio_wav_roundtrip(tmp_path):
    import wave
    from whisper_mlx_b200.audio import load_audio

    pcm = (synth.white_noise(1600, 0) * 32767).astype(np.int16)
    p = str(tmp_path / "a.wav")
    with wave.open(p, "wb") as w:
        w.setnchannels(1), w.setsampwidth(2), w.setframerate(16000)
        w.writeframes(pcm.tobytes())
    x = load_audio(p)
    assert x.dtype == np.float32 and np.array_equal(x, pcm.astype(np.float32) / 32768.0)
    with pytest.raises(RuntimeError, match="Failed to load audio"):
        load_audio(str(tmp_path / "missing.mp3"))


def test_load_model_errors(tmp_path):
    from whisper_mlx_b200.load_models import load_model

    with pytest.raises(FileNotFoundError):
        load_model(str(tmp_path / "nope"))


def test_daemon_tool_module(tmp_path):
    """The `transcribe_audio` tool follows the reference's plugin protocol (daemon/tools/base.py:23-105): a TOOL with
    spec + execute, JSON-string results, errors reported as {"error", "status": "error"} instead of raised."""
    import json

    from whisper_mlx_b200.tool import TOOL

    assert TOOL.name == "transcribe_audio" and TOOL.spec.name == TOOL.name
    schema = TOOL.to_schema()
    assert set(schema) == {"name", "description", "parameters"}
    assert schema["parameters"]["type"] == "object" and schema["parameters"]["required"] == ["file_path"]
    assert {"file_path", "language", "task", "model"} <= set(schema["parameters"]["properties"])
    r = json.loads(TOOL.execute(file_path=str(tmp_path / "missing.wav")))
    assert r["status"] == "error" and "File not found" in r["error"]
    p = tmp_path / "a.wav"
    p.write_bytes(b"")
    r = json.loads(TOOL.execute(file_path=str(p), task="summarise"))
    assert r["status"] == "error" and "Unsupported task" in r["error"]
    # a model that cannot be loaded is reported, not raised
    r = json.loads(TOOL.execute(file_path=str(p), model=str(tmp_path / "no_such_model")))
    assert r["status"] == "error"


def _mlx_quantize(w, group_size, bits):
    """Reference-side statement of mx.quantize: per group of `group_size` inputs, scale = (max - min) / (2^bits - 1),
    bias = min, q = round((w - bias) / scale), packed little-endian into uint32 words."""
    import torch

    o, i = w.shape
    g = w.view(o, i // group_size, group_size)
    lo, hi = g.min(-1).values, g.max(-1).values
    scale = ((hi - lo) / (2 ** bits - 1)).clamp_min(1e-8)
    q = torch.round((g - lo[..., None]) / scale[..., None]).clamp(0, 2 ** bits - 1).to(torch.int64).view(o, i)
    per = 32 // bits
    words = (q.view(o, i // per, per) << (torch.arange(per) * bits)).sum(-1)
    words = torch.where(words >= 2 ** 31, words - 2 ** 32, words).to(torch.int32)
    return words.view(torch.uint32) if hasattr(torch, "uint32") else words, scale, lo, (q.view(o, -1, group_size).float() * scale[..., None] + lo[..., None]).view(o, i)


@pytest.mark.parametrize("bits,group_size", [(4, 64), (8, 64), (2, 32), (4, 128)])
def test_mlx_dequantize(bits, group_size):
    import torch

    from whisper_mlx_b200.load_models import dequantize, dequantize_weights

    torch.manual_seed(bits * 100 + group_size)
    w = torch.randn(24, 256)
    words, scales, biases, expect = _mlx_quantize(w, group_size, bits)
    got = dequantize(words, scales, biases, group_size, bits)
    assert torch.equal(got, expect)
    assert (got - w).abs().max().item() <= (w.max() - w.min()).item() / (2 ** bits - 1)
    # fp16 scales / biases (what the checkpoints store) and the whole-dict form
    d = dequantize_weights({"a.weight": words, "a.scales": scales.half(), "a.biases": biases.half(), "b.weight": w}, group_size, bits)
    assert set(d) == {"a.weight", "b.weight"} and d["a.weight"].shape == w.shape
    with pytest.raises(NotImplementedError):
        dequantize(words, scales, biases, group_size, 3)


def test_word_distribution_matches_oracle():
    """The host half of timing.py (punctuation merging, duration heuristics, words -> segments, boundary reconciliation)
    against the oracle restatement, on synthetic alignments that hit every branch."""
    import copy
    import random

    from oracle import timing as OT
    from whisper_mlx_b200 import timing as T
    from whisper_mlx_b200.transcribe import _get_end, _is_segment_anomaly, _next_words_segment, _word_anomaly_score

    eot = 50257
    rnd = random.Random(7)
    vocab = [" hello", " world", ",", ".", " (", ")", " it", "'s", " a", " -", " test", "?", " “", "”", " long"]
    for trial in range(40):
        n = rnd.randint(1, 14)
        t = 0.0
        ali, tok = [], 1000
        for i in range(n):
            dur = rnd.choice([0.0, 0.08, 0.2, 0.3, 0.5, 2.6])
            word = rnd.choice(vocab)
            k = rnd.randint(1, 3)
            ali.append((word, list(range(tok, tok + k)), round(t, 2), round(t + dur, 2), rnd.random()))
            tok += k
            t += dur + rnd.choice([0.0, 0.1, 1.5])
        ali.append(("", [eot], round(t, 2), round(t, 2), 0.0))
        all_tokens = [x for a in ali[:-1] for x in a[1]]
        cut = rnd.randint(0, len(all_tokens))
        seek = rnd.choice([0, 3000, 4500])
        segments = [{"seek": seek, "start": seek / 100 + 0.0, "end": seek / 100 + t / 2, "tokens": [50365] + all_tokens[:cut] + [50400]},
                    {"seek": seek, "start": seek / 100 + t / 2, "end": seek / 100 + t + 0.7, "tokens": all_tokens[cut:]}]
        last = rnd.choice([0.0, seek / 100 - 3.0, seek / 100])
        s1, s2 = copy.deepcopy(segments), copy.deepcopy(segments)
        a1 = [T.WordTiming(*a) for a in copy.deepcopy(ali)]
        a2 = [OT.WordTiming(*a) for a in copy.deepcopy(ali)]
        kw = dict(eot=eot, prepend_punctuations="\"'“¿([{-", append_punctuations="\"'.。,，!！?？:：”)]}、", last_speech_timestamp=last)
        r1 = T.distribute_words(segments=s1, alignment=a1, **kw)
        r2 = OT.add_word_timestamps(segments=s2, alignment=a2, **kw)
        assert r1 == r2 and s1 == s2, trial
        assert [w.word for w in a1] == [w.word for w in a2]
        assert _get_end(s1) == next((w["end"] for s in reversed(s1) for w in reversed(s["words"])), s1[-1]["end"])
    # anomaly scoring (transcribe.py helpers of the hallucination filter)
    assert _word_anomaly_score({"word": "a", "start": 0.0, "end": 0.05, "probability": 0.1}) == pytest.approx(1.0 + (0.133 - 0.05) * 15)
    assert _word_anomaly_score({"word": "a", "start": 0.0, "end": 3.0, "probability": 0.9}) == pytest.approx(1.0)
    good = {"words": [{"word": " ok", "start": 0.0, "end": 0.3, "probability": 0.9}] * 4}
    bad = {"words": [{"word": " uh", "start": 0.0, "end": 0.02, "probability": 0.05}] * 4}
    assert not _is_segment_anomaly(good) and _is_segment_anomaly(bad) and not _is_segment_anomaly(None)
    assert _next_words_segment([{"words": []}, good]) is good and _next_words_segment([{"words": []}]) is None


def test_subtitle_word_options(tmp_path):
    """--highlight-words / --max-line-width / --max-line-count / --max-words-per-line act on the srt / vtt cues
    (they were parsed and ignored in r01); without word timings they are rejected, not dropped."""
    from whisper_mlx_b200.writers import get_writer

    def seg(i, t0, words):
        ws, t = [], t0
        for w in words:
            ws.append({"word": w, "start": round(t, 2), "end": round(t + 0.4, 2), "probability": 0.9})
            t += 0.5
        return {"id": i, "start": t0, "end": round(t, 2), "text": "".join(words), "words": ws}

    result = {"text": "", "language": "en", "segments": [seg(0, 0.0, [" The", " quick", " brown", " fox", " jumps"]),
                                                         seg(1, 10.0, [" over", " the", " lazy", " dog"])]}
    w = get_writer("srt", str(tmp_path))

    def cues(**kw):
        w(result, "o", **kw)
        blocks = open(tmp_path / "o.srt").read().strip().split("\n\n")
        return [b.split("\n", 2)[1:] for b in blocks]

    plain = cues()
    assert [c[1] for c in plain] == ["The quick brown fox jumps", "over the lazy dog"]
    assert plain[0][0] == "00:00:00,000 --> 00:00:02,400"  # first word start --> last word end
    wrapped = cues(max_line_width=12)
    assert [c[1] for c in wrapped] == ["The quick\nbrown fox\njumps", "over the\nlazy dog"]  # lines wrap, cues stay segments
    counted = cues(max_line_width=12, max_line_count=2)
    assert [c[1] for c in counted] == ["The quick\nbrown fox", "jumps", "over the\nlazy dog"]  # 3 s pause also closes a cue
    assert counted[1][0] == "00:00:02,000 --> 00:00:02,400"
    per = cues(max_words_per_line=2)
    assert [c[1] for c in per] == ["The quick", "brown fox", "jumps", "over the", "lazy dog"]
    hl = cues(highlight_words=True)
    assert hl[0] == ["00:00:00,000 --> 00:00:00,400", "<u>The</u> quick brown fox jumps"]
    assert hl[1] == ["00:00:00,400 --> 00:00:00,500", "The quick brown fox jumps"]  # the gap between two words
    assert hl[2][1] == "The <u>quick</u> brown fox jumps"
    # vtt goes through the same iterator
    get_writer("vtt", str(tmp_path))(result, "o", max_words_per_line=3)
    assert "The quick brown\n" in open(tmp_path / "o.vtt").read()
    no_words = {"segments": [{"id": 0, "start": 0.0, "end": 1.0, "text": " hi"}]}
    with pytest.raises(ValueError, match="word_timestamps=True"):
        w(no_words, "o", highlight_words=True)
    w(no_words, "o")
    # txt / tsv / json ignore the cue options (the CLI passes them to every writer)
    get_writer("all", str(tmp_path))(result, "all", max_line_width=12, max_line_count=None, highlight_words=False, max_words_per_line=None)
    assert open(tmp_path / "all.txt").read() == "The quick brown fox jumps\nover the lazy dog\n"
