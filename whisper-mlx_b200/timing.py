"""Word-level timestamps: host mirror of `mlx_whisper/timing.py` (UPSTREAM; the `--word-timestamps` neighbour of the
`./run` path -- it is what makes the `--hallucination-silence-threshold 1` of /root/reference/run:6 take effect;
SURVEY.md section 8f-3).

`find_alignment` runs on the GPU: one teacher-forced decoder forward that stores the cross-attention probabilities of
the alignment heads (K8 with probability output, `b200w_decoder_forward_full`), the normalise / median-filter / head-mean
stage (`b200w_alignment_matrix`) and dynamic time warping (`b200w_dtw`); the word bookkeeping (`merge_punctuations`,
the duration heuristics of `add_word_timestamps`) is host Python like the reference's.  Unlike the reference, which
re-encodes the mel segment, the encoder states of the window (already computed for decoding) are passed in.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List

import numpy as np
import torch

from . import _lib
from .audio import HOP_LENGTH, SAMPLE_RATE, TOKENS_PER_SECOND
from .tokenizer import Tokenizer


@dataclass
class WordTiming:
    word: str
    tokens: List[int]
    start: float
    end: float
    probability: float


def alignment_matrix(model, probs: torch.Tensor, first_layer: int, num_frames: int, seq: int = 0) -> torch.Tensor:
    """probs (layers >= first_layer, n_seq, n_tokens, n_head, 1500) f32 -> (n_tokens, num_frames // 2) f32 (K12a/b)."""
    lib = model._lib
    n_layers, n_seq, n_tok, n_head, n_ctx = probs.shape
    heads = np.asarray(model.alignment_heads, dtype=np.int32).reshape(-1, 2).copy()
    heads[:, 0] -= first_layer
    assert heads[:, 0].min() >= 0 and heads[:, 0].max() < n_layers
    heads_dev = torch.from_numpy(heads).to(probs.device)
    n_sel, n_frames = heads.shape[0], num_frames // 2
    stats = torch.empty(2 * n_sel * n_frames + n_sel * n_tok, dtype=torch.float32, device=probs.device)
    matrix = torch.empty((n_tok, n_frames), dtype=torch.float32, device=probs.device)
    with torch.cuda.device(probs.device):
        _lib.check(lib.b200w_alignment_matrix(_lib.ptr(probs), n_layers, n_seq, seq, n_tok, n_head, n_ctx, _lib.ptr(heads_dev),
                                              n_sel, n_frames, _lib.ptr(stats), _lib.ptr(matrix), _lib.stream()))
    return matrix


def dtw(model, matrix: torch.Tensor):
    """Monotonic alignment over -matrix (N, M) f32 CUDA (K12c).  Returns (text_indices, time_indices) like the reference."""
    lib = model._lib
    _lib.require_cuda(matrix, "matrix")
    N, M = matrix.shape
    dev = matrix.device
    cost = torch.empty((N + 1) * (M + 1), dtype=torch.float32, device=dev)
    trace = torch.empty((N + 1) * (M + 1), dtype=torch.int8, device=dev)
    idx = torch.empty((2, N + M), dtype=torch.int32, device=dev)
    n = torch.zeros(1, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.b200w_dtw(_lib.ptr(matrix), matrix.stride(0), N, M, _lib.ptr(cost), _lib.ptr(trace), _lib.ptr(idx[0]),
                                 _lib.ptr(idx[1]), _lib.ptr(n), _lib.stream()))
    n = int(n.item())
    path = idx[:, :n].cpu().numpy()[:, ::-1]
    return path[0].astype(np.int64), path[1].astype(np.int64)


def find_alignment(model, tokenizer: Tokenizer, text_tokens: List[int], features: torch.Tensor, num_frames: int, *,
                   medfilt_width: int = 7, qk_scale: float = 1.0) -> List[WordTiming]:
    if len(text_tokens) == 0:
        return []
    if medfilt_width != 7 or qk_scale != 1.0:
        raise NotImplementedError("the alignment kernels implement the reference defaults medfilt_width=7, qk_scale=1.0")
    dm = model.dims
    n_sot = len(tokenizer.sot_sequence)
    tokens = [*tokenizer.sot_sequence, tokenizer.no_timestamps, *text_tokens, tokenizer.eot]
    n = len(tokens)
    if n > dm.n_text_ctx:
        raise ValueError(f"{n} tokens exceed the text context ({dm.n_text_ctx})")
    sess = model.decode_session(1, 1, dm.n_text_ctx, slot=7)
    sess.load(features)
    sess.set_tokens(torch.tensor([tokens], dtype=torch.int32))
    first_layer = int(np.asarray(model.alignment_heads).reshape(-1, 2)[:, 0].min())
    logits, probs = sess.forward_full(n, first_layer)

    sampled = logits[n_sot: n_sot + len(text_tokens), : tokenizer.eot]
    logp = torch.log_softmax(sampled, dim=-1)
    tt = torch.tensor(text_tokens, dtype=torch.long, device=logp.device)
    text_token_probs = logp.gather(1, tt[:, None])[:, 0].exp().cpu().numpy()

    matrix = alignment_matrix(model, probs, first_layer, num_frames)
    text_indices, time_indices = dtw(model, matrix[n_sot: n - 1])

    words, word_tokens = tokenizer.split_to_word_tokens(text_tokens + [tokenizer.eot])
    if len(word_tokens) <= 1:
        # return on eot only: 'word_tokens[:-1]' below would be empty and the padded cumsum a 1-element boundary list
        return []
    word_boundaries = np.pad(np.cumsum([len(t) for t in word_tokens[:-1]]), (1, 0))
    jumps = np.pad(np.diff(text_indices), (1, 0), constant_values=1).astype(bool)
    jump_times = time_indices[jumps] / TOKENS_PER_SECOND
    start_times = jump_times[word_boundaries[:-1]]
    end_times = jump_times[word_boundaries[1:]]
    word_probabilities = [float(np.mean(text_token_probs[i:j])) for i, j in zip(word_boundaries[:-1], word_boundaries[1:])]
    return [WordTiming(word, toks, float(start), float(end), prob)
            for word, toks, start, end, prob in zip(words, word_tokens, start_times, end_times, word_probabilities)]


def merge_punctuations(alignment: List[WordTiming], prepended: str, appended: str) -> None:
    # merge prepended punctuations
    i = len(alignment) - 2
    j = len(alignment) - 1
    while i >= 0:
        previous, following = alignment[i], alignment[j]
        if previous.word.startswith(" ") and previous.word.strip() in prepended:
            # prepend it to the following word
            following.word = previous.word + following.word
            following.tokens = previous.tokens + following.tokens
            previous.word = ""
            previous.tokens = []
        else:
            j = i
        i -= 1
    # merge appended punctuations
    i = 0
    j = 1
    while j < len(alignment):
        previous, following = alignment[i], alignment[j]
        if not previous.word.endswith(" ") and following.word in appended:
            # append it to the previous word
            previous.word = previous.word + following.word
            previous.tokens = previous.tokens + following.tokens
            following.word = ""
            following.tokens = []
        else:
            i = j
        j += 1


def distribute_words(*, segments: List[dict], alignment: List[WordTiming], eot: int, prepend_punctuations: str,
                     append_punctuations: str, last_speech_timestamp: float) -> float:
    """Everything of `add_word_timestamps` after the alignment: long-word truncation at sentence ends, punctuation
    merging, handing the words to their segments and reconciling word and segment boundaries."""
    text_tokens_per_segment = [[token for token in segment["tokens"] if token < eot] for segment in segments]
    word_durations = np.array([t.end - t.start for t in alignment])
    word_durations = word_durations[word_durations.nonzero()]
    median_duration = float(np.median(word_durations)) if len(word_durations) > 0 else 0.0
    median_duration = min(0.7, median_duration)
    max_duration = median_duration * 2

    # hack: truncate long words at sentence boundaries (a bug of the alignment, not of the audio)
    if len(word_durations) > 0:
        sentence_end_marks = ".。!！?？"
        for i in range(1, len(alignment)):
            if alignment[i].end - alignment[i].start > max_duration:
                if alignment[i].word in sentence_end_marks:
                    alignment[i].end = alignment[i].start + max_duration
                elif alignment[i - 1].word in sentence_end_marks:
                    alignment[i].start = alignment[i].end - max_duration

    merge_punctuations(alignment, prepend_punctuations, append_punctuations)

    time_offset = segments[0]["seek"] * HOP_LENGTH / SAMPLE_RATE
    word_index = 0
    for segment, text_tokens in zip(segments, text_tokens_per_segment):
        saved_tokens = 0
        words = []
        while word_index < len(alignment) and saved_tokens < len(text_tokens):
            timing = alignment[word_index]
            if timing.word:
                words.append(dict(word=timing.word, start=round(time_offset + timing.start, 2),
                                  end=round(time_offset + timing.end, 2), probability=timing.probability))
            saved_tokens += len(timing.tokens)
            word_index += 1

        if len(words) > 0:
            # hack: ensure the first word does not start long before the previous speech ended
            if words[0]["end"] - last_speech_timestamp > median_duration * 4 and (
                    words[0]["end"] - words[0]["start"] > max_duration
                    or (len(words) > 1 and words[1]["end"] - words[0]["start"] > max_duration * 2)):
                if len(words) > 1 and words[1]["end"] - words[1]["start"] > max_duration:
                    boundary = max(words[1]["end"] / 2, words[1]["end"] - max_duration)
                    words[0]["end"] = words[1]["start"] = boundary
                words[0]["start"] = max(0, words[0]["end"] - max_duration)
            # prefer the segment-level start timestamp if the first word is too long
            if segment["start"] < words[0]["end"] and segment["start"] - 0.5 > words[0]["start"]:
                words[0]["start"] = max(0, min(words[0]["end"] - median_duration, segment["start"]))
            else:
                segment["start"] = words[0]["start"]
            # prefer the segment-level end timestamp if the last word is too long
            if segment["end"] > words[-1]["start"] and segment["end"] + 0.5 < words[-1]["end"]:
                words[-1]["end"] = max(words[-1]["start"] + median_duration, segment["end"])
            else:
                segment["end"] = words[-1]["end"]
            last_speech_timestamp = segment["end"]
        segment["words"] = words
    return last_speech_timestamp


def add_word_timestamps(*, segments: List[dict], model, tokenizer: Tokenizer, features: torch.Tensor, num_frames: int,
                        prepend_punctuations: str = "\"'“¿([{-", append_punctuations: str = "\"'.。,，!！?？:：”)]}、",
                        last_speech_timestamp: float, **kwargs) -> float:
    """Attach `words` ({word, start, end, probability}) to every segment of one window.  Returns the updated
    last_speech_timestamp (the reference mutates a caller-side variable through its loop instead)."""
    if len(segments) == 0:
        return last_speech_timestamp
    if num_frames // 2 < 2:
        # a window of under 40 ms has a single attention frame: the per-frame standard deviation over tokens is zero and
        # the reference's normalisation divides by it (NaN alignment); no word boundaries can be placed there
        for segment in segments:
            segment["words"] = []
        return last_speech_timestamp
    text_tokens = [t for segment in segments for t in segment["tokens"] if t < tokenizer.eot]
    alignment = find_alignment(model, tokenizer, text_tokens, features, num_frames, **kwargs)
    return distribute_words(segments=segments, alignment=alignment, eot=tokenizer.eot,
                            prepend_punctuations=prepend_punctuations, append_punctuations=append_punctuations,
                            last_speech_timestamp=last_speech_timestamp)
