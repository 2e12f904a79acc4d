"""Control flow of `transcribe()` on a CPU: the device side (log-mel, encoder, decoder) is replaced by a stand-in
"speaker" with a known word timeline, so that what the two modes RECOVER can be compared word by word.

The fixed-window batched mode must not lose audio: the reference re-seeks to the last closed
timestamp of a window, so speech running across a window end is decoded again; with fixed 30 s strides those words
have to come from the leftover-text segment or from the tail window (whisper-mlx_b200/transcribe.py docstring).
"""
import numpy as np
import pytest
import torch

from whisper_mlx_b200.decoding import DecodingResult
from whisper_mlx_b200.tokenizer import get_tokenizer
from whisper_mlx_b200.transcribe import transcribe


class _Dims:
    n_mels, n_audio_ctx, n_vocab, n_text_ctx = 128, 1500, 51866, 448


class _Model:
    dims = _Dims()
    is_multilingual, num_languages, device, model_path = True, 100, "cpu", None


class SpeakerBackend:
    """A deterministic stand-in for Whisper on continuous speech: sentences [(start_s, end_s, [word ids])] back to back.

    decode() of the window [t0, t1) behaves like the model does on real speech: sentences that end inside the window
    come out as <|start|> words <|end|>; the sentence running across the window end is either only opened
    ("<|27.00|><|27.00|>" then EOT, even sentences: the reference re-seeks to 27.00) or left unfinished (odd ones);
    a window that starts inside a sentence yields the remaining words from <|0.00|>."""

    def __init__(self, sentences, total_s):
        self.model = _Model()
        self.sentences = sentences
        self.total_s = total_s
        self.calls = []  # (seek, size) of every decoded window

    def mel_frames(self, audio):
        return int(self.total_s * 100) + 3000

    def features(self, seeks, sizes):
        return torch.tensor([list(p) for p in zip(seeks, sizes)], dtype=torch.int64)

    def detect_language(self, xa):
        return [50259], [{"en": 1.0}]

    def decode(self, features, options, tokenizer):
        tb = tokenizer.timestamp_begin
        out = []
        for seek, size in features.tolist():
            self.calls.append((seek, size))
            t0, t1 = seek / 100.0, (seek + size) / 100.0
            toks, closed = [], 0
            for si, (s, e, words) in enumerate(self.sentences):
                if e <= t0 or s >= t1:
                    continue
                # words are spread evenly over the sentence; a word belongs to the window its time falls in
                times = [s + (e - s) * (k + 0.5) / len(words) for k in range(len(words))]
                inside = [w for w, t in zip(words, times) if t0 <= t < t1]
                start = max(s, t0)
                if e <= t1:
                    if inside:
                        toks += [tb + round((start - t0) / 0.02), *inside, tb + round((e - t0) / 0.02)]
                        closed += 1
                else:  # runs across the end of the window
                    if si % 2 == 0 and closed > 0:
                        toks.append(tb + round((start - t0) / 0.02))  # "...<|27.00|><|27.00|>" EOT: opened, nothing fits
                        break
                    if inside:
                        toks += [tb + round((start - t0) / 0.02), *inside]
            out.append(DecodingResult(audio_features=None, language="en", tokens=toks, text=tokenizer.decode(toks),
                                      avg_logprob=-0.1, no_speech_prob=0.0, temperature=0.0, compression_ratio=1.0))
        return out


def _speech(total_s=200.0, seed=0):
    rng = np.random.default_rng(seed)
    sentences, t, wid = [], 0.0, 1000
    while t < total_s - 0.5:
        dur = float(rng.choice([1.5, 2.5, 3.7, 5.1, 8.3]))
        e = min(round((t + dur) / 0.02) * 0.02, total_s)
        n = int(rng.integers(3, 9))
        sentences.append((t, e, list(range(wid, wid + n))))
        wid += n
        t = e
    return sentences


def _words(result, tb=50365):
    return [t for s in result["segments"] for t in s["tokens"] if t < 50257]


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_fixed_windows_recover_every_word(seed):
    total = 200.0
    sentences = _speech(total, seed)
    truth = [w for _, _, ws in sentences for w in ws]
    audio = np.zeros(int(total * 16000), dtype=np.float32)
    kw = dict(temperature=0.0, condition_on_previous_text=False, language="en")

    exact_b = SpeakerBackend(sentences, total)
    exact = transcribe(audio, _backend=exact_b, window_batch=0, **kw)
    assert _words(exact) == truth  # the reference's seek loop: every word exactly once, in order

    batched_b = SpeakerBackend(sentences, total)
    batched = transcribe(audio, _backend=batched_b, window_batch=4, **kw)
    assert _words(batched) == truth, "the fixed-window mode lost or duplicated words"
    planned = [c for c in batched_b.calls if c[0] % 3000 == 0 and c[1] in (3000, int(total * 100) % 3000)]
    tails = [c for c in batched_b.calls if c not in planned]
    assert len(planned) == 7 and len(tails) >= 1  # some windows closed early: their tails were decoded
    for seek, size in tails:
        assert 0 < size < 3000 and (seek + size) % 3000 == 0  # a tail ends where its planned window ends
    # segments are ordered and inside the file; leftover-text segments end at a window end
    last = 0.0
    for s in batched["segments"]:
        assert s["start"] >= last - 1e-6 and s["end"] <= total + 1e-6 and s["end"] >= s["start"]
        last = s["start"]
    assert [s["id"] for s in batched["segments"]] == list(range(len(batched["segments"])))

    # what ADVICE r01 described: without the tail pass the words between a window's last timestamp and its end are gone
    lossy = transcribe(audio, _backend=SpeakerBackend(sentences, total), window_batch=4, max_tail_rounds=0, **kw)
    assert len(_words(lossy)) < len(truth)


def test_fixed_windows_shard_like_one_rank():
    """rank / world_size without torch.distributed initialised: a rank decodes only its own windows and their tails."""
    total = 200.0
    sentences = _speech(total, 3)
    audio = np.zeros(int(total * 16000), dtype=np.float32)
    kw = dict(temperature=0.0, condition_on_previous_text=False, language="en", window_batch=2)
    b = SpeakerBackend(sentences, total)
    with pytest.raises(RuntimeError, match="no rank produced windows"):
        transcribe(audio, _backend=b, rank=1, world_size=2, **kw)  # (the gather needs the other rank)
    assert all(seek >= 9000 for seek, _ in b.calls)  # rank 1 of 2 owns windows 3..6 of 7
    with pytest.raises(ValueError, match="fixed-window mode"):
        transcribe(audio, _backend=SpeakerBackend(sentences, total), rank=0, world_size=2, temperature=0.0, language="en")


def test_exact_mode_keeps_reference_clear_order():
    """Instantaneous segments keep their tokens until after the word-timestamp pass (the reference clears once, at the
    end of the window loop body): without word timestamps the observable result is cleared text / tokens."""

    class B(SpeakerBackend):
        def decode(self, features, options, tokenizer):
            tb = tokenizer.timestamp_begin
            toks = [tb + 10, 500, 501, tb + 10, tb + 10, 502, tb + 200, tb + 200]  # first segment has start == end
            return [DecodingResult(audio_features=None, language="en", tokens=toks, text="x", avg_logprob=-0.1,
                                   no_speech_prob=0.0, temperature=0.0, compression_ratio=1.0) for _ in features.tolist()]

    r = transcribe(np.zeros(16000 * 3, dtype=np.float32), _backend=B([], 3.0), temperature=0.0, language="en",
                   condition_on_previous_text=False)
    assert r["segments"][0]["tokens"] == [] and r["segments"][0]["text"] == ""
    assert r["segments"][1]["tokens"][1] == 502


class _TaggedBackend(SpeakerBackend):
    """A SpeakerBackend whose feature rows carry the file they belong to: (file, seek, size)."""

    def __init__(self, fid, sentences, total_s):
        super().__init__(sentences, total_s)
        self.fid = fid

    def features(self, seeks, sizes):
        return torch.tensor([[self.fid, s, z] for s, z in zip(seeks, sizes)], dtype=torch.int64)

    def decode(self, features, options, tokenizer):
        return super().decode(features[:, 1:], options, tokenizer)


def test_transcribe_many_is_the_exact_loop_per_file():
    """Several files in lockstep (transcribe_many): every file's result equals its own exact-mode run, the decoder calls
    of the files share batches, and a file that fails is reported without stopping the others."""
    from whisper_mlx_b200.transcribe import transcribe_many

    totals = [200.0, 95.0, 310.0]
    speeches = [_speech(t, seed) for t, seed in zip(totals, (11, 12, 13))]
    kw = dict(language="en", temperature=0.0, condition_on_previous_text=False, no_speech_threshold=None)
    alone = [transcribe(np.zeros(1), _backend=_TaggedBackend(i, sp, t), **kw) for i, (sp, t) in enumerate(zip(speeches, totals))]
    backends = [_TaggedBackend(i, sp, t) for i, (sp, t) in enumerate(zip(speeches, totals))]
    batches = []

    def decode(features, options, tokenizer):
        batches.append(features[:, 0].tolist())
        out = []
        for row in features:
            out.extend(backends[int(row[0])].decode(row[None], options, tokenizer))
        return out

    got = transcribe_many([np.zeros(1)] * 3, _backends=backends, _decode=decode, **kw)
    assert len(got) == 3
    for g, a in zip(got, alone):
        assert isinstance(g, dict) and g["text"] == a["text"]
        assert [(s["seek"], s["start"], s["end"], s["tokens"]) for s in g["segments"]] == \
               [(s["seek"], s["start"], s["end"], s["tokens"]) for s in a["segments"]]
    # lockstep: while all three files run, their windows are decoded together; the longest file finishes alone
    assert batches[0] == [0, 1, 2] and batches[-1] == [2] and max(len(b) for b in batches) == 3
    assert sum(len(b) for b in batches) == sum(len(b.calls) for b in backends)

    # a failing file does not take the others down
    class _Broken(_TaggedBackend):
        def mel_frames(self, audio):
            raise RuntimeError("cannot read this file")

    backends = [_TaggedBackend(0, speeches[0], totals[0]), _Broken(1, speeches[1], totals[1]), _TaggedBackend(2, speeches[2], totals[2])]
    got = transcribe_many([np.zeros(1)] * 3, _backends=backends, _decode=decode, **kw)
    assert isinstance(got[1], RuntimeError) and got[0]["text"] == alone[0]["text"] and got[2]["text"] == alone[2]["text"]
