"""Oracle: long-form driver (test infrastructure only; see oracle/__init__.py).

Restates `mlx_whisper/transcribe.py::transcribe` (UPSTREAM, not under /root/reference; call site
/root/reference/run:3-6) per SURVEY.md section 3.1 / Appendix A.5, for the path `./run` reaches
(word_timestamps=False).  Sequential seek, one 30 s window at a time, batch 1.
"""
from __future__ import annotations

from typing import Optional, Sequence, Union

import numpy as np
import torch

from . import audio as A
from . import decoding as D
from . import model as M
from .tokens import TokenIds, decode_text, LANGUAGE_CODES

HOP, SR, N_FRAMES = A.HOP_LENGTH, A.SAMPLE_RATE, A.N_FRAMES


def _pad_or_trim_frames(mel: np.ndarray, length: int = N_FRAMES) -> np.ndarray:
    return A.pad_or_trim(mel, length, axis=-2)


@torch.no_grad()
def transcribe(w, dims: M.ModelDimensions, audio: np.ndarray, *,
               temperature: Union[float, Sequence[float]] = (0.0, 0.2, 0.4, 0.6, 0.8, 1.0),
               compression_ratio_threshold: Optional[float] = 2.4, logprob_threshold: Optional[float] = -1.0,
               no_speech_threshold: Optional[float] = 0.6, condition_on_previous_text: bool = True,
               language: Optional[str] = None, task: str = "transcribe", policy: str = "fp32",
               sample_len: Optional[int] = None, fixed_windows: bool = False,
               max_tail_rounds: int = 8, decode_fn=None):
    """Returns {"text", "segments", "language"}.

    `fixed_windows=True` is the batched-mode contract of the product (not a mode of the
    reference): windows start at fixed 30 s strides instead of at the last decoded timestamp.  Where the reference
    would re-seek (tokens left after the last closed timestamp pair) the leftover text becomes a segment ending at
    the window end; if there is no leftover text the uncovered tail [last timestamp, window end) is decoded as a
    window of its own before the next planned window (at most `max_tail_rounds` times per planned window).

    `decode_fn(seek, segment_size, segment_mel, prompt_tokens, temperature) -> DecodingResult` replaces the model for
    one window (tests replay the tokens another implementation decoded, to check the control flow around them).
    """
    ids = TokenIds(dims.n_vocab)
    mel = A.log_mel_spectrogram(audio, dims.n_mels, padding=A.N_SAMPLES)
    content_frames = mel.shape[-2] - N_FRAMES
    if language is None:
        seg = torch.from_numpy(_pad_or_trim_frames(mel))[None]
        xa = M.encoder_forward(w, dims, seg, policy=policy)
        lang_tok, _ = D.detect_language(w, dims, xa, policy=policy)
        language = LANGUAGE_CODES[lang_tok[0] - ids.language_begin]
    input_stride = N_FRAMES // dims.n_audio_ctx
    time_precision = input_stride * HOP / SR
    temps = [temperature] if isinstance(temperature, (int, float)) else list(temperature)

    all_tokens, all_segments = [], []
    prompt_reset_since = 0
    seek = 0

    def decode_with_fallback(segment, seek, segment_size):
        result = None
        for t in temps:
            if decode_fn is not None:
                result = decode_fn(seek, segment_size, segment, all_tokens[prompt_reset_since:], float(t))
            else:
                result = D.decode(w, dims, segment, language=language, task=task, temperature=float(t),
                                  prompt=all_tokens[prompt_reset_since:], policy=policy, sample_len=sample_len)[0]
            needs_fallback = False
            if compression_ratio_threshold is not None and result.compression_ratio > compression_ratio_threshold:
                needs_fallback = True
            if logprob_threshold is not None and result.avg_logprob < logprob_threshold:
                needs_fallback = True
            if no_speech_threshold is not None and result.no_speech_prob > no_speech_threshold:
                needs_fallback = False
            if not needs_fallback:
                break
        return result

    def new_segment(start, end, toks, result):
        toks = [int(t) for t in toks]
        text_tokens = [t for t in toks if t < ids.eot]
        return {"seek": seek, "start": start, "end": end, "text": decode_text(text_tokens, ids.timestamp_begin),
                "tokens": toks, "temperature": result.temperature, "avg_logprob": result.avg_logprob,
                "compression_ratio": result.compression_ratio, "no_speech_prob": result.no_speech_prob}

    window_end = 0     # fixed-window mode: end of the planned window being decoded, and how many of its tails were taken
    tail_rounds = 0
    while seek < content_frames:
        time_offset = seek * HOP / SR
        segment_size = min(N_FRAMES, content_frames - seek)
        if fixed_windows:
            if seek >= window_end:  # a new planned window
                window_end = seek + segment_size
                tail_rounds = 0
            segment_size = window_end - seek
        segment_duration = segment_size * HOP / SR
        segment = torch.from_numpy(_pad_or_trim_frames(mel[seek: seek + segment_size]))[None]
        result = decode_with_fallback(segment, seek, segment_size)
        tokens = np.array(result.tokens, dtype=np.int64)

        if no_speech_threshold is not None:
            should_skip = result.no_speech_prob > no_speech_threshold
            if logprob_threshold is not None and result.avg_logprob > logprob_threshold:
                should_skip = False
            if should_skip:
                seek += segment_size
                continue
        window_start = seek

        current = []
        ts = tokens >= ids.timestamp_begin
        single_ending = ts[-2:].tolist() == [False, True]
        consecutive = np.where(np.logical_and(ts[:-1], ts[1:]))[0] + 1
        if len(consecutive) > 0:
            slices = consecutive.tolist()
            if single_ending:
                slices.append(len(tokens))
            last = 0
            for cur in slices:
                sl = tokens[last:cur]
                start_pos = int(sl[0]) - ids.timestamp_begin
                end_pos = int(sl[-1]) - ids.timestamp_begin
                current.append(new_segment(time_offset + start_pos * time_precision,
                                           time_offset + end_pos * time_precision, sl, result))
                last = cur
            if single_ending:
                seek += segment_size
            elif not fixed_windows:
                seek += (int(tokens[last - 1]) - ids.timestamp_begin) * input_stride
            else:
                advance = (int(tokens[last - 1]) - ids.timestamp_begin) * input_stride
                trailing = tokens[last:]
                if any(int(t) < ids.eot for t in trailing):
                    first = int(trailing[0])
                    start_pos = first - ids.timestamp_begin if first >= ids.timestamp_begin else advance // input_stride
                    current.append(new_segment(time_offset + start_pos * time_precision, time_offset + segment_duration,
                                               trailing, result))
                    seek += segment_size
                elif 0 < advance < segment_size and tail_rounds < max_tail_rounds:
                    tail_rounds += 1
                    seek += advance  # the tail [seek, window_end) is decoded next
                else:
                    seek += segment_size
        else:
            duration = segment_duration
            stamps = tokens[ts.nonzero()[0]]
            if len(stamps) > 0 and stamps[-1] != ids.timestamp_begin:
                duration = (int(stamps[-1]) - ids.timestamp_begin) * time_precision
            current.append(new_segment(time_offset, time_offset + duration, tokens, result))
            seek += segment_size

        for s in current:
            if s["start"] == s["end"] or s["text"].strip() == "":
                s["text"] = ""
                s["tokens"] = []
        for s in current:
            s["seek"] = window_start
        all_segments.extend({"id": i, **s} for i, s in enumerate(current, start=len(all_segments)))
        all_tokens.extend(t for s in current for t in s["tokens"])
        if not condition_on_previous_text or result.temperature > 0.5:
            prompt_reset_since = len(all_tokens)

    return {"text": decode_text(all_tokens, ids.timestamp_begin), "segments": all_segments, "language": language}
