"""Audio front-end: host mirror of `mlx_whisper/audio.py` (UPSTREAM; reached from /root/reference/run:3;
restated in SURVEY.md A.1) over the fused sm_100a log-mel kernel (csrc/logmel.cu).

Same names, argument meaning and error behaviour as the reference module: `load_audio`, `pad_or_trim`,
`mel_filters`, `hanning`, `log_mel_spectrogram` and the constants.  Arrays are torch CUDA tensors
instead of `mx.array`s.
"""
from __future__ import annotations

import os
import subprocess
from functools import lru_cache
from typing import Optional, Tuple, Union

import numpy as np
import torch

from . import _lib

# hard-coded audio hyperparameters
SAMPLE_RATE = 16000
N_FFT = 400
HOP_LENGTH = 160
CHUNK_LENGTH = 30
N_SAMPLES = CHUNK_LENGTH * SAMPLE_RATE  # 480000 samples in a 30-second chunk
N_FRAMES = N_SAMPLES // HOP_LENGTH  # 3000 frames in a mel spectrogram input
N_SAMPLES_PER_TOKEN = HOP_LENGTH * 2  # the initial convolutions has stride 2
FRAMES_PER_SECOND = SAMPLE_RATE // HOP_LENGTH  # 10ms per audio frame
TOKENS_PER_SECOND = SAMPLE_RATE // N_SAMPLES_PER_TOKEN  # 20ms per audio token


def _read_wav(file: str, sr: int) -> Optional[np.ndarray]:
    """PCM WAV reader used when ffmpeg is unavailable (mono / 16 kHz s16le, the format ffmpeg is asked for)."""
    import wave

    try:
        with wave.open(file, "rb") as w:
            if w.getframerate() != sr or w.getsampwidth() != 2:
                return None
            data = np.frombuffer(w.readframes(w.getnframes()), dtype=np.int16)
            if w.getnchannels() > 1:
                data = data.reshape(-1, w.getnchannels()).astype(np.int32).mean(axis=1).astype(np.int16)
            return data
    except (wave.Error, EOFError, OSError):
        return None


def load_audio(file: str, sr: int = SAMPLE_RATE) -> np.ndarray:
    """Decode `file` to mono float32 at `sr` Hz with the ffmpeg CLI (s16le), scaled by 1/32768.

    Raises RuntimeError("Failed to load audio: ...") like the reference when decoding fails.
    """
    cmd = ["ffmpeg", "-nostdin", "-threads", "0", "-i", file, "-f", "s16le", "-ac", "1", "-acodec", "pcm_s16le",
           "-ar", str(sr), "-"]
    try:
        out = subprocess.run(cmd, capture_output=True, check=True).stdout
        pcm = np.frombuffer(out, np.int16)
    except FileNotFoundError:
        pcm = _read_wav(file, sr)
        if pcm is None:
            raise RuntimeError(f"Failed to load audio: ffmpeg is not installed and {file!r} is not a "
                               f"{sr} Hz 16-bit PCM WAV file") from None
    except subprocess.CalledProcessError as e:
        raise RuntimeError(f"Failed to load audio: {e.stderr.decode()}") from e
    return pcm.flatten().astype(np.float32) / 32768.0


def pad_or_trim(array, length: int = N_SAMPLES, *, axis: int = -1):
    """Pad with zeros at the end or trim `array` to `length` along `axis` (numpy arrays or torch tensors)."""
    if isinstance(array, torch.Tensor):
        if array.shape[axis] > length:
            array = array.narrow(axis, 0, length)
        if array.shape[axis] < length:
            pad = [0, 0] * array.ndim
            ax = axis % array.ndim
            pad[2 * (array.ndim - 1 - ax) + 1] = length - array.shape[axis]
            array = torch.nn.functional.pad(array, pad)
        return array
    if array.shape[axis] > length:
        sl = [slice(None)] * array.ndim
        sl[axis] = slice(0, length)
        array = array[tuple(sl)]
    if array.shape[axis] < length:
        pad_widths = [(0, 0)] * array.ndim
        pad_widths[axis] = (0, length - array.shape[axis])
        array = np.pad(array, pad_widths)
    return array


def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    lin = f / (200.0 / 3)
    log = 15.0 + np.log(np.maximum(f, 1e-30) / 1000.0) / (np.log(6.4) / 27.0)
    return np.where(f >= 1000.0, log, lin)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    return np.where(m >= 15.0, 1000.0 * np.exp((np.log(6.4) / 27.0) * (m - 15.0)), (200.0 / 3) * m)


@lru_cache(maxsize=None)
def _mel_filters_np(n_mels: int) -> np.ndarray:
    n_freqs = N_FFT // 2 + 1
    freqs = np.linspace(0.0, SAMPLE_RATE / 2, n_freqs)
    edges = _mel_to_hz(np.linspace(_hz_to_mel(0.0), _hz_to_mel(SAMPLE_RATE / 2), n_mels + 2))
    width = np.diff(edges)
    slopes = edges[:, None] - freqs[None, :]
    fb = np.maximum(0.0, np.minimum(-slopes[:-2] / width[:-1, None], slopes[2:] / width[1:, None]))
    fb *= (2.0 / (edges[2:] - edges[:-2]))[:, None]  # slaney area normalisation
    return fb.astype(np.float32)


def mel_filters(n_mels: int) -> np.ndarray:
    """(n_mels, 201) f32 mel filterbank (slaney scale and norm, = librosa.filters.mel(16000, 400, n_mels)).

    The reference loads the same matrix from assets/mel_filters.npz; it is regenerated here.
    """
    assert n_mels in {80, 128}, f"Unsupported n_mels: {n_mels}"
    return _mel_filters_np(n_mels)


@lru_cache(maxsize=None)
def hanning(size: int) -> np.ndarray:
    return np.hanning(size + 1)[:-1]


class _Tables:
    """Device-resident constant tables of the log-mel kernel (window, FFT twiddles) for one device.

    The mel filterbank is compiled into the kernel (csrc/mel_tables.h, generated from `mel_filters`)."""

    def __init__(self, device: torch.device, n_mels: int):
        n2, k1 = np.arange(25)[:, None], np.arange(16)[None, :]
        ang = -2.0 * np.pi * (n2 * k1) / 400.0
        tw = np.stack([np.cos(ang), np.sin(ang)], axis=-1).astype(np.float32)
        t = lambda a, dt: torch.tensor(np.asarray(a), dtype=dt, device=device)  # noqa: E731
        self.hann = t(hanning(N_FFT).astype(np.float32), torch.float32)
        self.tw = t(tw, torch.float32)
        self.struct = _lib.LogmelTables(_lib.ptr(self.hann), _lib.ptr(self.tw))


_tables = {}


def _get_tables(device: torch.device, n_mels: int) -> _Tables:
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device(), n_mels)
    if key not in _tables:
        _tables[key] = _Tables(device, n_mels)
    return _tables[key]


def load_audio_pcm16(file: str, sr: int = SAMPLE_RATE) -> np.ndarray:
    """`load_audio` without its last line: the mono s16le samples as int16.  The log-mel kernel applies the
    `/ 32768.0` itself (b200w_logmel_pcm16), so a file goes to the GPU at two bytes per sample."""
    cmd = ["ffmpeg", "-nostdin", "-threads", "0", "-i", file, "-f", "s16le", "-ac", "1", "-acodec", "pcm_s16le",
           "-ar", str(sr), "-"]
    try:
        return np.frombuffer(subprocess.run(cmd, capture_output=True, check=True).stdout, np.int16).flatten()
    except FileNotFoundError:
        pcm = _read_wav(file, sr)
        if pcm is None:
            raise RuntimeError(f"Failed to load audio: ffmpeg is not installed and {file!r} is not a "
                               f"{sr} Hz 16-bit PCM WAV file") from None
        return pcm.flatten()
    except subprocess.CalledProcessError as e:
        raise RuntimeError(f"Failed to load audio: {e.stderr.decode()}") from e


def _to_device_audio(audio, device=None) -> torch.Tensor:
    """Samples on the GPU: f32 in [-1, 1], or int16 PCM (files, int16 arrays) which K1 scales itself."""
    if isinstance(audio, str):
        audio = load_audio_pcm16(audio)
    if isinstance(audio, np.ndarray):
        audio = torch.from_numpy(np.ascontiguousarray(audio, dtype=np.int16 if audio.dtype == np.int16 else np.float32))
    if not isinstance(audio, torch.Tensor):
        raise TypeError(f"Unsupported audio type: {type(audio)}")
    if not audio.is_cuda:
        if not torch.cuda.is_available():
            raise RuntimeError("log_mel_spectrogram needs a CUDA device (B200); there is no CPU fallback")
        audio = audio.to(device or "cuda", non_blocking=True)
    return audio.to(torch.int16 if audio.dtype == torch.int16 else torch.float32).contiguous()


def _run_logmel(audio: torch.Tensor, n_mels: int, padding: int, normalized: bool,
                out: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    lib = _lib.load()
    x = audio if audio.ndim == 2 else audio[None]
    _lib.require_cuda(x, "audio")
    n_audio, n_valid = x.shape
    n_total = n_valid + padding
    if n_total <= N_FFT // 2:
        raise ValueError(f"audio of {n_valid} samples (+{padding} padding) is shorter than the {N_FFT // 2}-sample reflect pad")
    if n_valid == 0:  # empty input with padding: all-zero signal; the kernel still wants a valid pointer
        x = torch.zeros((n_audio, 1), dtype=x.dtype, device=x.device)
    n_frames = n_total // HOP_LENGTH
    tb = _get_tables(x.device, n_mels)
    if out is None:
        out = torch.empty((n_audio, n_frames, n_mels), dtype=torch.float32, device=x.device)
    elif out.shape != (n_audio, n_frames, n_mels) or out.dtype != torch.float32 or not out.is_contiguous() or out.device != x.device:
        raise ValueError(f"out must be a contiguous f32 tensor of shape {(n_audio, n_frames, n_mels)} on {x.device}")
    gmax = torch.empty((n_audio,), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        if normalized:
            done = torch.empty((n_audio,), dtype=torch.int32, device=x.device)
            _lib.check(lib.b200w_logmel_normalized(_lib.ptr(x), int(x.dtype == torch.int16), n_audio, x.stride(0), n_valid, n_total,
                                                   n_mels, tb.struct, _lib.ptr(out), _lib.ptr(gmax), _lib.ptr(done), _lib.stream()))
        else:
            fn = lib.b200w_logmel_pcm16 if x.dtype == torch.int16 else lib.b200w_logmel
            _lib.check(fn(_lib.ptr(x), n_audio, x.stride(0), n_valid, n_total, n_mels, tb.struct, _lib.ptr(out), _lib.ptr(gmax),
                          _lib.stream()))
    return out, gmax


def log_mel_unclamped(audio: torch.Tensor, n_mels: int = 80, padding: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
    """Run K1 on `audio` ((n,) or (n_audio, n) f32 or int16-PCM CUDA).  Returns (log10 mel before the clamp, per-audio max).

    The clamp / scale is fused into the bf16 window gather that feeds the encoder (`Whisper.mel_windows`);
    `log_mel_spectrogram` runs the kernel with the normalisation fused instead.
    """
    return _run_logmel(audio, n_mels, padding, normalized=False)


def log_mel_spectrogram(audio: Union[str, np.ndarray, torch.Tensor], n_mels: int = 80, padding: int = 0,
                        device=None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Log-mel spectrogram, (frames, n_mels) f32 on the GPU, time-major like the reference.

    `audio`: path (decoded with ffmpeg), NumPy array or torch tensor of 16 kHz mono samples in [-1, 1] (int16
    arrays are taken as s16le PCM and scaled by 1/32768 on the device);
    `padding` zero samples are appended first.  A 2-D input (n_audio, n) gives (n_audio, frames, n_mels)
    with one clamp maximum per row (the batched contract of BASELINE config 2).  `out` (not in the reference): a
    preallocated (n_audio, frames, n_mels) f32 result buffer to write into instead of allocating one per call.
    """
    x = _to_device_audio(audio, device)
    batched = x.ndim == 2
    if os.environ.get("B200W_LOGMEL_TWO_PASS") == "1":  # A/B: the r01 form, clamp as a second pass over the result
        out, gmax = _run_logmel(x, n_mels, padding, normalized=False, out=out)
        lib = _lib.load()
        with torch.cuda.device(out.device):
            _lib.check(lib.b200w_logmel_finalize(_lib.ptr(out), _lib.ptr(gmax), out.shape[0], out.shape[1] * out.shape[2],
                                                 _lib.stream()))
    else:
        out, _ = _run_logmel(x, n_mels, padding, normalized=True, out=out)
    return out if batched else out[0]
