"""Oracle: word-level timestamps (test infrastructure only; see oracle/__init__.py -- parity unpinned).

Restates `mlx_whisper/timing.py` (UPSTREAM, not under /root/reference; the `--word-timestamps` neighbour of the
`./run` path), itself the MLX port of the published openai-whisper algorithm:
`median_filter`, `dtw` (the numba `dtw_cpu` recurrence and `backtrace`), `find_alignment`, `merge_punctuations`,
`add_word_timestamps`.  NumPy / pure Python, sized for the short sequences of the tests.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Sequence

import numpy as np
import torch

from . import audio as A
from . import model as M

TOKENS_PER_SECOND = A.SAMPLE_RATE // (A.HOP_LENGTH * 2)  # 50: one encoder position per 20 ms


def median_filter(x: np.ndarray, filter_width: int) -> np.ndarray:
    """Median filter of odd width along the last axis, reflect padding (timing.py::median_filter)."""
    pad = filter_width // 2
    if x.shape[-1] <= pad:
        return x
    assert filter_width > 0 and filter_width % 2 == 1, "`filter_width` should be an odd number"
    xp = np.pad(x, [(0, 0)] * (x.ndim - 1) + [(pad, pad)], mode="reflect")
    win = np.lib.stride_tricks.sliding_window_view(xp, filter_width, axis=-1)
    return np.sort(win, axis=-1)[..., pad]


def dtw(x: np.ndarray):
    """timing.py::dtw_cpu + backtrace on cost matrix x (N, M), float32 accumulation like the numba kernel."""
    N, Mm = x.shape
    cost = np.full((N + 1, Mm + 1), np.inf, dtype=np.float32)
    trace = -np.ones((N + 1, Mm + 1), dtype=np.float32)
    cost[0, 0] = 0
    x = x.astype(np.float32)
    for j in range(1, Mm + 1):
        for i in range(1, N + 1):
            c0, c1, c2 = cost[i - 1, j - 1], cost[i - 1, j], cost[i, j - 1]
            if c0 < c1 and c0 < c2:
                c, t = c0, 0
            elif c1 < c0 and c1 < c2:
                c, t = c1, 1
            else:
                c, t = c2, 2
            cost[i, j] = x[i - 1, j - 1] + c
            trace[i, j] = t
    i, j = N, Mm
    trace[0, :] = 2
    trace[:, 0] = 1
    result = []
    while i > 0 or j > 0:
        result.append((i - 1, j - 1))
        if trace[i, j] == 0:
            i -= 1
            j -= 1
        elif trace[i, j] == 1:
            i -= 1
        elif trace[i, j] == 2:
            j -= 1
        else:
            raise ValueError("Unexpected trace[i, j]")
    result = np.array(result)
    return result[::-1, :].T


@dataclass
class WordTiming:
    word: str
    tokens: List[int]
    start: float
    end: float
    probability: float


def alignment_matrix(cross_qk: Sequence[torch.Tensor], alignment_heads: np.ndarray, num_frames: int,
                     medfilt_width: int = 7, qk_scale: float = 1.0) -> np.ndarray:
    """(n_tokens, num_frames // 2): softmax over the kept frames, normalise over tokens, median filter, head mean."""
    weights = np.stack([cross_qk[l][0, h].float().numpy() for l, h in alignment_heads])  # (sel, tokens, 1500)
    weights = weights[:, :, : num_frames // 2].astype(np.float64) * qk_scale
    weights = np.exp(weights - weights.max(-1, keepdims=True))
    weights = weights / weights.sum(-1, keepdims=True)
    mean = weights.mean(axis=-2, keepdims=True)
    std = weights.std(axis=-2, keepdims=True)
    weights = (weights - mean) / std
    weights = median_filter(weights, medfilt_width)
    return weights.mean(axis=0)


@torch.no_grad()
def find_alignment(w, dims: M.ModelDimensions, alignment_heads: np.ndarray, sot_sequence: Sequence[int], no_timestamps: int,
                   eot: int, split_to_word_tokens: Callable, text_tokens: List[int], xa: torch.Tensor, num_frames: int,
                   *, medfilt_width: int = 7, qk_scale: float = 1.0, policy: str = "fp32") -> List[WordTiming]:
    if len(text_tokens) == 0:
        return []
    tokens = torch.tensor([*sot_sequence, no_timestamps, *text_tokens, eot], dtype=torch.long)
    logits, _, cross_qk = M.decoder_forward(w, dims, tokens[None], xa, policy=policy, return_cross_qk=True)
    logits = logits[0]
    sampled_logits = logits[len(sot_sequence):, :eot].double()
    token_probs = torch.softmax(sampled_logits, dim=-1).numpy()
    text_token_probs = token_probs[np.arange(len(text_tokens)), text_tokens]

    matrix = alignment_matrix(cross_qk, alignment_heads, num_frames, medfilt_width, qk_scale)
    matrix = matrix[len(sot_sequence): -1]
    text_indices, time_indices = dtw(-matrix)

    words, word_tokens = split_to_word_tokens(text_tokens + [eot])
    if len(word_tokens) <= 1:
        return []
    word_boundaries = np.pad(np.cumsum([len(t) for t in word_tokens[:-1]]), (1, 0))
    jumps = np.pad(np.diff(text_indices), (1, 0), constant_values=1).astype(bool)
    jump_times = time_indices[jumps] / TOKENS_PER_SECOND
    start_times = jump_times[word_boundaries[:-1]]
    end_times = jump_times[word_boundaries[1:]]
    word_probabilities = [np.mean(text_token_probs[i:j]) for i, j in zip(word_boundaries[:-1], word_boundaries[1:])]
    return [WordTiming(word, tokens, float(start), float(end), float(prob))
            for word, tokens, start, end, prob in zip(words, word_tokens, start_times, end_times, word_probabilities)]


def merge_punctuations(alignment: List[WordTiming], prepended: str, appended: str) -> None:
    # merge prepended punctuations
    i = len(alignment) - 2
    j = len(alignment) - 1
    while i >= 0:
        previous, following = alignment[i], alignment[j]
        if previous.word.startswith(" ") and previous.word.strip() in prepended:
            following.word = previous.word + following.word
            following.tokens = previous.tokens + following.tokens
            previous.word = ""
            previous.tokens = []
        else:
            j = i
        i -= 1
    # merge appended punctuations
    i, j = 0, 1
    while j < len(alignment):
        previous, following = alignment[i], alignment[j]
        if not previous.word.endswith(" ") and following.word in appended:
            previous.word = previous.word + following.word
            previous.tokens = previous.tokens + following.tokens
            following.word = ""
            following.tokens = []
        else:
            i = j
        j += 1


def add_word_timestamps(*, segments: List[dict], alignment: List[WordTiming], eot: int, prepend_punctuations: str,
                        append_punctuations: str, last_speech_timestamp: float) -> float:
    """The post-alignment half of timing.py::add_word_timestamps: duration heuristics, punctuation merging and the
    distribution of words over `segments` (which gain "words" and possibly adjusted start / end).  Returns the new
    last_speech_timestamp."""
    if len(segments) == 0:
        return last_speech_timestamp
    text_tokens_per_segment = [[t for t in s["tokens"] if t < eot] for s in segments]
    word_durations = np.array([t.end - t.start for t in alignment])
    word_durations = word_durations[word_durations.nonzero()]
    median_duration = float(np.median(word_durations)) if len(word_durations) > 0 else 0.0
    median_duration = min(0.7, median_duration)
    max_duration = median_duration * 2
    # hack: truncate long words at sentence boundaries
    if len(word_durations) > 0:
        sentence_end_marks = ".。!！?？"
        for i in range(1, len(alignment)):
            if alignment[i].end - alignment[i].start > max_duration:
                if alignment[i].word in sentence_end_marks:
                    alignment[i].end = alignment[i].start + max_duration
                elif alignment[i - 1].word in sentence_end_marks:
                    alignment[i].start = alignment[i].end - max_duration
    merge_punctuations(alignment, prepend_punctuations, append_punctuations)

    time_offset = segments[0]["seek"] * A.HOP_LENGTH / A.SAMPLE_RATE
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
