"""Long-form driver: host mirror of `mlx_whisper/transcribe.py::transcribe` (UPSTREAM; the function
`./run` reaches through the `mlx_whisper` console script, /root/reference/run:3-6; restated in SURVEY.md
section 3.1 and A.5).

Same signature, option names, thresholds and result dictionary.  Two execution modes:

* exact (default): the reference's sequential seek loop -- one 30 s window at a time, the next window
  starts at the last decoded timestamp.  Bit-for-bit the reference control flow.
* fixed-window batches (`window_batch=N`): the file is cut into back-to-back 30 s windows which are
  independent when `condition_on_previous_text=False` (the `./run` setting); N windows are encoded and
  decoded together.  This is the throughput mode the benchmarks measure.  The
  reference re-seeks to the last closed timestamp of a window, so that speech running across the window
  end is decoded again by the next window; with fixed strides that audio would be lost.  Here no audio is
  dropped: text left unfinished at the end of a window becomes a segment that ends at the window end, and
  a window that closes its last segment early WITHOUT further text gets its uncovered tail
  [last timestamp, window end) decoded as a window of its own in a follow-up batch (tails of one rank's
  windows stay on that rank, so sharding needs no exchange).  Segment boundaries can still differ from
  exact mode because window starts do not follow the decoded timestamps.
"""
from __future__ import annotations

import os
import sys
import warnings
from typing import List, Optional, Tuple, Union

import numpy as np
import torch

from .audio import FRAMES_PER_SECOND, HOP_LENGTH, N_FRAMES, N_SAMPLES, SAMPLE_RATE, _to_device_audio, log_mel_unclamped
from .decoding import DecodingOptions, DecodingResult, DecodingTask, detect_language
from .load_models import load_model
from .sharding import gather_by_index, plan_windows, shard_indices
from .timing import add_word_timestamps
from .tokenizer import LANGUAGES, get_tokenizer


def _format_timestamp(seconds: float):
    assert seconds >= 0, "non-negative timestamp expected"
    milliseconds = round(seconds * 1000.0)
    hours = milliseconds // 3_600_000
    milliseconds -= hours * 3_600_000
    minutes = milliseconds // 60_000
    milliseconds -= minutes * 60_000
    seconds = milliseconds // 1_000
    milliseconds -= seconds * 1_000
    hours_marker = f"{hours:02d}:" if hours > 0 else ""
    return f"{hours_marker}{minutes:02d}:{seconds:02d}.{milliseconds:03d}"


class ModelHolder:
    """One cached model per process, like the reference (not re-entrant)."""

    model = None
    model_path = None

    @classmethod
    def get_model(cls, model_path: str, dtype=torch.bfloat16):
        if cls.model is None or model_path != cls.model_path:
            cls.model = load_model(model_path, dtype=dtype)
            cls.model_path = model_path
        return cls.model


_PUNCTUATION = "\"'“¿([{-\"'.。,，!！?？:：”)]}、"


def _word_anomaly_score(word: dict) -> float:
    probability = word.get("probability", 0.0)
    duration = word["end"] - word["start"]
    score = 0.0
    if probability < 0.15:
        score += 1.0
    if duration < 0.133:
        score += (0.133 - duration) * 15
    if duration > 2.0:
        score += duration - 2.0
    return score


def _is_segment_anomaly(segment: Optional[dict]) -> bool:
    if segment is None or not segment["words"]:
        return False
    words = [w for w in segment["words"] if w["word"] not in _PUNCTUATION]
    words = words[:8]
    score = sum(_word_anomaly_score(w) for w in words)
    return score >= 3 or score + 0.01 >= len(words)


def _next_words_segment(segments: List[dict]) -> Optional[dict]:
    return next((s for s in segments if s["words"]), None)


def _get_end(segments: List[dict]) -> Optional[float]:
    return next((w["end"] for s in reversed(segments) for w in reversed(s["words"])),
                segments[-1]["end"] if segments else None)


def _clear_empty_segments(segments: List[dict], with_words: bool) -> None:
    """If a segment is instantaneous or does not contain text, clear it (the reference does this once per window,
    AFTER the word-timestamp pass)."""
    for segment in segments:
        if segment["start"] == segment["end"] or segment["text"].strip() == "":
            segment["text"] = ""
            segment["tokens"] = []
            if with_words:
                segment["words"] = []


def _segments_for_window(tokens: np.ndarray, seek: int, segment_size: int, result: DecodingResult, tokenizer,
                         input_stride: int, time_precision: float):
    """Split one window's tokens at consecutive timestamp pairs, exactly like the reference's loop body.

    Returns (segments, seek advance, single_timestamp_ending, trailing): `trailing` are the tokens after the last closed
    pair which the reference ignores because it re-seeks to that timestamp (empty when nothing is left over)."""
    time_offset = float(seek * HOP_LENGTH / SAMPLE_RATE)
    segment_duration = segment_size * HOP_LENGTH / SAMPLE_RATE

    def new_segment(*, start: float, end: float, toks, res: DecodingResult):
        toks = [int(t) for t in toks]
        text_tokens = [t for t in toks if t < tokenizer.eot]
        return {"seek": seek, "start": start, "end": end, "text": tokenizer.decode(text_tokens), "tokens": toks,
                "temperature": res.temperature, "avg_logprob": res.avg_logprob,
                "compression_ratio": res.compression_ratio, "no_speech_prob": res.no_speech_prob}

    current_segments = []
    trailing = tokens[:0]
    timestamp_tokens = tokens >= tokenizer.timestamp_begin
    single_timestamp_ending = timestamp_tokens[-2:].tolist() == [False, True]
    consecutive = np.where(np.logical_and(timestamp_tokens[:-1], timestamp_tokens[1:]))[0]
    consecutive += 1
    if len(consecutive) > 0:
        slices = consecutive.tolist()
        if single_timestamp_ending:
            slices.append(len(tokens))
        last_slice = 0
        for current_slice in slices:
            sliced_tokens = tokens[last_slice:current_slice]
            start_timestamp_pos = int(sliced_tokens[0]) - tokenizer.timestamp_begin
            end_timestamp_pos = int(sliced_tokens[-1]) - tokenizer.timestamp_begin
            current_segments.append(new_segment(start=time_offset + start_timestamp_pos * time_precision,
                                                end=time_offset + end_timestamp_pos * time_precision,
                                                toks=sliced_tokens, res=result))
            last_slice = current_slice
        if single_timestamp_ending:
            # single timestamp at the end means no speech after the last timestamp
            advance = segment_size
        else:
            # otherwise, ignore the unfinished segment and seek to the last timestamp
            last_timestamp_pos = int(tokens[last_slice - 1]) - tokenizer.timestamp_begin
            advance = last_timestamp_pos * input_stride
            trailing = tokens[last_slice:]
    else:
        duration = segment_duration
        timestamps = tokens[timestamp_tokens.nonzero()[0]]
        if len(timestamps) > 0 and timestamps[-1] != tokenizer.timestamp_begin:
            # no consecutive timestamps but it has a timestamp; use the last one
            last_timestamp_pos = int(timestamps[-1]) - tokenizer.timestamp_begin
            duration = last_timestamp_pos * time_precision
        current_segments.append(new_segment(start=time_offset, end=time_offset + duration, toks=tokens, res=result))
        advance = segment_size
    return current_segments, advance, single_timestamp_ending, trailing


def _segments_fixed_window(tokens: np.ndarray, seek: int, segment_size: int, result: DecodingResult, tokenizer,
                           input_stride: int, time_precision: float):
    """Fixed-window mode: the reference's segmentation, then what it would have left to the next (re-seeked) window.

    Returns (segments, tail): `tail` is None or the (seek, size) of the uncovered end of this window, to be decoded as
    a window of its own (module docstring)."""
    segs, advance, _, trailing = _segments_for_window(tokens, seek, segment_size, result, tokenizer, input_stride,
                                                      time_precision)
    tail = None
    if len(trailing) > 0:  # the reference would re-seek to the last closed timestamp (trailing starts with the opening one)
        if any(int(t) < tokenizer.eot for t in trailing):
            # unfinished text: keep it, as a segment from its opening timestamp to the end of the window
            time_offset = float(seek * HOP_LENGTH / SAMPLE_RATE)
            first = int(trailing[0])
            start_pos = (first - tokenizer.timestamp_begin) if first >= tokenizer.timestamp_begin else advance // input_stride
            toks = [int(t) for t in trailing]
            segs.append({"seek": seek, "start": time_offset + start_pos * time_precision,
                         "end": time_offset + segment_size * HOP_LENGTH / SAMPLE_RATE,
                         "text": tokenizer.decode([t for t in toks if t < tokenizer.eot]), "tokens": toks,
                         "temperature": result.temperature, "avg_logprob": result.avg_logprob,
                         "compression_ratio": result.compression_ratio, "no_speech_prob": result.no_speech_prob})
        elif 0 < advance < segment_size:
            tail = (seek + advance, segment_size - advance)
    return segs, tail


def transcribe(
    audio: Union[str, np.ndarray, torch.Tensor],
    *,
    path_or_hf_repo: str = "mlx-community/whisper-tiny",
    verbose: Optional[bool] = None,
    temperature: Union[float, Tuple[float, ...]] = (0.0, 0.2, 0.4, 0.6, 0.8, 1.0),
    compression_ratio_threshold: Optional[float] = 2.4,
    logprob_threshold: Optional[float] = -1.0,
    no_speech_threshold: Optional[float] = 0.6,
    condition_on_previous_text: bool = True,
    initial_prompt: Optional[str] = None,
    word_timestamps: bool = False,
    prepend_punctuations: str = "\"'“¿([{-",
    append_punctuations: str = "\"'.。,，!！?？:：”)]}、",
    clip_timestamps: Union[str, List[float]] = "0",
    hallucination_silence_threshold: Optional[float] = None,
    window_batch: Optional[int] = None,
    encoder_batch: int = 32,
    model=None,
    rank: int = 0,
    world_size: int = 1,
    max_tail_rounds: int = 8,
    window_trace: Optional[list] = None,
    _backend=None,
    _run_decode=None,
    **decode_options,
):
    """Transcribe an audio file (path, NumPy array or torch tensor of 16 kHz mono samples).

    Returns {"text": str, "segments": [...], "language": str} exactly like the reference.  Extra keyword
    arguments (not in the reference): `window_batch` selects the fixed-window batched mode (module
    docstring; default from $B200W_WINDOW_BATCH, else exact sequential mode), `encoder_batch` bounds how
    many windows go through one encoder call, `model` passes an already loaded `Whisper`; `rank` /
    `world_size` (batched mode, one process per GPU with torch.distributed initialised) make this process
    decode only its block of windows and gather the per-window segments on the host, so every rank returns
    the full result; `max_tail_rounds` bounds the follow-up batches that decode uncovered window tails
    (0: never re-decode a tail); `window_trace`, a list, receives one dict per decoded window (seek, size, the
    DecodingResult fields) in decoding order -- diagnostics / parity tests.  `_backend` (tests only) replaces the device side -- log-mel, encoder and
    decoder -- with a stand-in so that the control flow can be exercised on a CPU.
    """
    if word_timestamps and decode_options.get("task", "transcribe") == "translate" and verbose:
        warnings.warn("Word-level timestamps on translations may not be reliable.")
    if hallucination_silence_threshold is not None and not word_timestamps and verbose:
        warnings.warn("--hallucination_silence_threshold requires --word_timestamps True; it has no effect")

    dtype = torch.bfloat16 if decode_options.get("fp16", True) else torch.float32
    if _backend is not None:
        model = _backend.model
    elif model is None:
        model = ModelHolder.get_model(path_or_hf_repo, dtype)
    if window_batch is None:
        window_batch = int(os.environ.get("B200W_WINDOW_BATCH", "0"))

    if _backend is not None:
        n_mel_frames = _backend.mel_frames(audio)
        features_for = _backend.features
        run_decode = _backend.decode
    else:
        # whole-file log-mel with 30 s of zero padding; the clamp uses the file-wide maximum (K1 + K1b)
        pcm = _to_device_audio(audio, model.device)
        if pcm.ndim != 1:
            raise ValueError("transcribe() takes one mono signal")
        mel, gmax = log_mel_unclamped(pcm, n_mels=model.dims.n_mels, padding=N_SAMPLES)
        mel2d = mel[0]
        n_mel_frames = mel2d.shape[-2]

        def features_for(seeks: List[int], sizes: List[int]) -> torch.Tensor:
            """Encoder states (n, 1500, d) of the windows starting at mel rows `seeks` with `sizes` valid frames."""
            feats = []
            for e0 in range(0, len(seeks), encoder_batch):
                sk, sz = seeks[e0: e0 + encoder_batch], sizes[e0: e0 + encoder_batch]
                feats.append(model.encode_slabs(model.mel_windows(mel2d, gmax, sk, sz, [0] * len(sk))))
            return torch.cat(feats, 0) if len(feats) > 1 else feats[0]

        def run_decode(features, options: DecodingOptions, tokenizer) -> List[DecodingResult]:
            return DecodingTask(model, options, tokenizer=tokenizer).run_features(features)

    if _run_decode is not None:  # transcribe_many(): the decoder calls of several files meet in one batch
        run_decode = _run_decode

    content_frames = n_mel_frames - N_FRAMES
    content_duration = float(content_frames * HOP_LENGTH / SAMPLE_RATE)

    if decode_options.get("language", None) is None:
        if not model.is_multilingual:
            decode_options["language"] = "en"
        else:
            if verbose:
                print("Detecting language using up to the first 30 seconds. Use the `language` decoding option to specify the language")
            xa0 = features_for([0], [min(N_FRAMES, n_mel_frames)])
            _, probs = _backend.detect_language(xa0) if _backend is not None else detect_language(model, xa0)
            decode_options["language"] = max(probs[0], key=probs[0].get)
            if verbose is not None:
                print(f"Detected language: {LANGUAGES[decode_options['language']].title()}")

    language: str = decode_options["language"]
    task: str = decode_options.get("task", "transcribe")
    tokenizer = get_tokenizer(model.is_multilingual, num_languages=model.num_languages, language=language, task=task,
                              vocab_dir=getattr(model, "model_path", None))

    if isinstance(clip_timestamps, str):
        clip_timestamps = [float(ts) for ts in (clip_timestamps.split(",") if clip_timestamps else [])]
    seek_points: List[int] = [round(ts * FRAMES_PER_SECOND) for ts in clip_timestamps]
    if len(seek_points) == 0:
        seek_points.append(0)
    if len(seek_points) % 2 == 1:
        seek_points.append(content_frames)
    else:
        seek_points[-1] = min(content_frames, seek_points[-1])
    seek_clips: List[Tuple[int, int]] = list(zip(seek_points[::2], seek_points[1::2]))

    temperatures = [temperature] if isinstance(temperature, (int, float)) else list(temperature)

    def needs_fallback(res: DecodingResult) -> bool:
        fallback = False
        if compression_ratio_threshold is not None and res.compression_ratio > compression_ratio_threshold:
            fallback = True  # too repetitive
        if logprob_threshold is not None and res.avg_logprob < logprob_threshold:
            fallback = True  # average log probability is too low
        if no_speech_threshold is not None and res.no_speech_prob > no_speech_threshold:
            fallback = False  # silence
        return fallback

    def options_for(t: float, prompt) -> DecodingOptions:
        kwargs = {**decode_options}
        if t > 0:
            kwargs.pop("beam_size", None)  # disable beam_size and patience when t > 0
            kwargs.pop("patience", None)
        else:
            kwargs.pop("best_of", None)  # disable best_of when t == 0
        kwargs.pop("fp16", None)
        return DecodingOptions(**kwargs, prompt=prompt, temperature=t)

    def decode_with_fallback(features: torch.Tensor, prompt) -> List[DecodingResult]:
        """features (n, 1500, d): every window is retried at the next temperature until it passes."""
        n = features.shape[0]
        results: List[Optional[DecodingResult]] = [None] * n
        pending = list(range(n))
        for t in temperatures:
            out = run_decode(features[pending], options_for(float(t), prompt), tokenizer)
            still = []
            for i, r in zip(pending, out):
                results[i] = r
                if needs_fallback(r):
                    still.append(i)
            pending = still
            if not pending:
                break
        return results

    def trace(seek: int, size: int, res: DecodingResult) -> None:
        if window_trace is not None:
            window_trace.append(dict(seek=seek, size=size, tokens=list(res.tokens), avg_logprob=res.avg_logprob,
                                     no_speech_prob=res.no_speech_prob, temperature=res.temperature,
                                     compression_ratio=res.compression_ratio, text=res.text))

    input_stride = N_FRAMES // model.dims.n_audio_ctx  # mel frames per output token: 2
    time_precision = input_stride * HOP_LENGTH / SAMPLE_RATE  # time per output token: 0.02 (seconds)
    all_tokens: List[int] = []
    all_segments: List[dict] = []
    prompt_reset_since = 0
    last_speech_timestamp = 0.0
    if initial_prompt is not None:
        initial_prompt_tokens = tokenizer.encode(" " + initial_prompt.strip())
        all_tokens.extend(initial_prompt_tokens)
    else:
        initial_prompt_tokens = []

    def should_skip(res: DecodingResult) -> bool:
        if no_speech_threshold is None:
            return False
        skip = res.no_speech_prob > no_speech_threshold
        if logprob_threshold is not None and res.avg_logprob > logprob_threshold:
            skip = False  # don't skip if the logprob is high enough, despite the no_speech_prob
        return skip

    def emit(current_segments: List[dict]):
        if verbose:
            for segment in current_segments:
                print(f"[{_format_timestamp(segment['start'])} --> {_format_timestamp(segment['end'])}] {segment['text']}")
        # the reference clears instantaneous / empty segments once per window, after the word pass and the printing
        _clear_empty_segments(current_segments, with_words=word_timestamps)
        all_segments.extend({"id": i, **segment} for i, segment in enumerate(current_segments, start=len(all_segments)))
        all_tokens.extend(token for segment in current_segments for token in segment["tokens"])

    batched = window_batch > 0 and not condition_on_previous_text and initial_prompt is None and not word_timestamps
    if world_size > 1 and not batched:
        raise ValueError("sharding over ranks needs the fixed-window mode (window_batch > 0, "
                         "condition_on_previous_text=False, no initial_prompt): exact mode is sequential per file")
    if window_batch > 0 and not batched and verbose:
        warnings.warn("window_batch needs condition_on_previous_text=False and no initial_prompt; using exact mode")

    if batched:
        # ---------------- fixed 30 s windows, window_batch at a time ----------------
        windows = plan_windows(content_frames, seek_clips, N_FRAMES)
        mine = shard_indices(len(windows), rank, world_size)  # all windows when world_size == 1
        local = {i: [] for i in mine}
        pending = [(i, windows[i][0], windows[i][1]) for i in mine]  # (planned window, seek, size)
        rounds = 0
        while pending:
            tails = []
            for i0 in range(0, len(pending), window_batch):
                chunk = pending[i0: i0 + window_batch]
                features = features_for([c[1] for c in chunk], [c[2] for c in chunk])
                results = decode_with_fallback(features, [])
                for (i, seek, size), res in zip(chunk, results):
                    trace(seek, size, res)
                    if should_skip(res):
                        continue
                    tokens = np.array(res.tokens, dtype=np.int64)
                    segs, tail = _segments_fixed_window(tokens, seek, size, res, tokenizer, input_stride, time_precision)
                    local[i].extend(segs)
                    if tail is not None and rounds < max_tail_rounds:
                        tails.append((i, tail[0], tail[1]))
            pending = tails  # uncovered window ends, decoded as windows of their own (a tail is shorter than its parent)
            rounds += 1
        # host-side gather of the per-window segments (no device collective on the data path)
        for segs in gather_by_index(local, len(windows)) if world_size > 1 else [local[i] for i in range(len(windows))]:
            if segs:
                emit(segs)
    else:
        # ---------------- exact mode: the reference's sequential seek loop ----------------
        clip_idx = 0
        seek = seek_clips[0][0]
        while clip_idx < len(seek_clips):
            seek_clip_start, seek_clip_end = seek_clips[clip_idx]
            if seek < seek_clip_start:
                seek = seek_clip_start
            if seek >= seek_clip_end:
                clip_idx += 1
                if clip_idx < len(seek_clips):
                    seek = seek_clips[clip_idx][0]
                continue
            segment_size = min(N_FRAMES, content_frames - seek, seek_clip_end - seek)
            features = features_for([seek], [segment_size])
            result = decode_with_fallback(features, all_tokens[prompt_reset_since:])[0]
            tokens = np.array(result.tokens, dtype=np.int64)
            trace(seek, segment_size, result)
            if should_skip(result):
                seek += segment_size  # fast-forward to the next segment boundary
                continue
            previous_seek = seek
            time_offset = float(seek * HOP_LENGTH / SAMPLE_RATE)
            window_end_time = float((seek + N_FRAMES) * HOP_LENGTH / SAMPLE_RATE)
            segment_duration = segment_size * HOP_LENGTH / SAMPLE_RATE
            segs, advance, single_timestamp_ending, _ = _segments_for_window(tokens, seek, segment_size, result, tokenizer,
                                                                             input_stride, time_precision)
            seek += advance
            if word_timestamps:
                last_speech_timestamp = add_word_timestamps(
                    segments=segs, model=model, tokenizer=tokenizer, features=features, num_frames=segment_size,
                    prepend_punctuations=prepend_punctuations, append_punctuations=append_punctuations,
                    last_speech_timestamp=last_speech_timestamp)
                if not single_timestamp_ending:
                    last_word_end = _get_end(segs)
                    if last_word_end is not None and last_word_end > time_offset:
                        seek = round(last_word_end * FRAMES_PER_SECOND)
                # skip silence before possible hallucinations
                if hallucination_silence_threshold is not None:
                    threshold = hallucination_silence_threshold
                    if not single_timestamp_ending:
                        last_word_end = _get_end(segs)
                        if last_word_end is not None and last_word_end > time_offset:
                            remaining_duration = window_end_time - last_word_end
                            if remaining_duration > threshold:
                                seek = round(last_word_end * FRAMES_PER_SECOND)
                            else:
                                seek = previous_seek + segment_size
                    # if the first segment might be a hallucination, skip the leading silence
                    first_segment = _next_words_segment(segs)
                    if first_segment is not None and _is_segment_anomaly(first_segment):
                        gap = first_segment["start"] - time_offset
                        if gap > threshold:
                            seek = previous_seek + round(gap * FRAMES_PER_SECOND)
                            continue
                    # skip silence before any possible hallucination that is surrounded by silence or more hallucinations
                    hal_last_end = last_speech_timestamp
                    for si in range(len(segs)):
                        segment = segs[si]
                        if not segment["words"]:
                            continue
                        if _is_segment_anomaly(segment):
                            next_segment = _next_words_segment(segs[si + 1:])
                            if next_segment is not None:
                                hal_next_start = next_segment["words"][0]["start"]
                            else:
                                hal_next_start = time_offset + segment_duration
                            silence_before = (segment["start"] - hal_last_end > threshold or segment["start"] < threshold
                                              or segment["start"] - time_offset < 2.0)
                            silence_after = (hal_next_start - segment["end"] > threshold or _is_segment_anomaly(next_segment)
                                             or window_end_time - segment["end"] < 2.0)
                            if silence_before and silence_after:
                                seek = round(max(time_offset + 1, segment["start"]) * FRAMES_PER_SECOND)
                                if content_duration - segment["end"] < threshold:
                                    seek = content_frames
                                segs[si:] = []
                                break
                        hal_last_end = segment["end"]
                last_word_end = _get_end(segs)
                if last_word_end is not None:
                    last_speech_timestamp = last_word_end
            emit(segs)
            if not condition_on_previous_text or result.temperature > 0.5:
                # do not feed the prompt tokens if a high temperature was used
                prompt_reset_since = len(all_tokens)

    return dict(text=tokenizer.decode(all_tokens[len(initial_prompt_tokens):]), segments=all_segments, language=language)


def transcribe_many(audios, *, path_or_hf_repo: str = "mlx-community/whisper-tiny", model=None, _backends=None, _decode=None,
                    **kwargs) -> list:
    """Exact (sequential-seek) transcription of SEVERAL files in lockstep -- not in the reference, whose CLI walks its
    `audio+` arguments one by one at batch 1.

    Every file runs the reference's seek loop unchanged (`transcribe()` in a thread of its own); only the decoder calls
    meet: when every file that is still running waits for a window, the windows are decoded as ONE batch (requests with
    equal DecodingOptions share a batch, so the temperature fallback and per-file prompts keep their meaning), and each
    file goes on with its own result.  A single-token step costs 1.31 ms for one window and 1.9-2.0 ms for three to
    seven (K13m), so n files finish in well under n times the time of one.  One lock serialises everything that
    touches the device, so the library sees one caller at a time.  Results can differ from file-by-file runs where two
    tokens tie within rounding (the batch size selects the kernel: different summation order), as with `window_batch`.

    Returns one entry per file: the result dictionary of `transcribe()` or the exception that file raised.
    `_backends` / `_decode` (tests only): stand-ins for the device side per file and for the batched decoder call."""
    import threading

    n = len(audios)
    if n == 0:
        return []
    if _backends is None:
        if model is None:
            model = ModelHolder.get_model(path_or_hf_repo, torch.bfloat16 if kwargs.get("fp16", True) else torch.float32)

        def batched_decode(features, options, tokenizer):
            return DecodingTask(model, options, tokenizer=tokenizer).run_features(features)
    else:
        batched_decode = _decode
    device_lock = threading.Lock()
    cond = threading.Condition()
    state = {"active": n, "pending": []}
    results: list = [None] * n

    def hook(features, options, tokenizer):
        req = {"features": features, "options": options, "tokenizer": tokenizer, "result": None, "error": None, "done": False}
        device_lock.release()  # (held by this worker whenever it is not waiting here)
        try:
            with cond:
                state["pending"].append(req)
                cond.notify_all()
                while not req["done"]:
                    cond.wait()
        finally:
            device_lock.acquire()
        if req["error"] is not None:
            raise req["error"]
        return req["result"]

    def worker(i: int) -> None:
        device_lock.acquire()
        try:
            kw = dict(kwargs)
            if _backends is not None:
                kw["_backend"] = _backends[i]
            else:
                kw["model"] = model
            results[i] = transcribe(audios[i], _run_decode=hook, **kw)
        except BaseException as e:  # noqa: BLE001 - reported per file, like the reference CLI
            results[i] = e
        finally:
            device_lock.release()
            with cond:
                state["active"] -= 1
                cond.notify_all()

    threads = [threading.Thread(target=worker, args=(i,), daemon=True) for i in range(n)]
    for t in threads:
        t.start()
    while True:
        with cond:
            while state["active"] > 0 and len(state["pending"]) < state["active"]:
                cond.wait()
            if not state["pending"]:
                break  # every file has finished
            batch, state["pending"] = state["pending"], []
        groups: dict = {}
        for r in batch:
            tk = r["tokenizer"]
            groups.setdefault((repr(r["options"]), getattr(tk, "language", None), getattr(tk, "task", None)), []).append(r)
        with device_lock:
            for reqs in groups.values():
                try:
                    feats = torch.cat([r["features"] for r in reqs], 0) if len(reqs) > 1 else reqs[0]["features"]
                    out = batched_decode(feats, reqs[0]["options"], reqs[0]["tokenizer"])
                    k = 0
                    for r in reqs:
                        m = r["features"].shape[0]
                        r["result"] = list(out[k: k + m])
                        k += m
                except Exception as e:  # noqa: BLE001 - handed to the files of this batch
                    for r in reqs:
                        r["error"] = e
        with cond:
            for r in batch:
                r["done"] = True
            cond.notify_all()
    for t in threads:
        t.join()
    return results
