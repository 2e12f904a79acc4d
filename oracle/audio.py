"""Oracle: log-mel front-end (test infrastructure only; see oracle/__init__.py).

Restates `mlx_whisper/audio.py` (UPSTREAM, not under /root/reference; call site
/root/reference/run:3) as described in SURVEY.md Appendix A.1, cross-pinned against
transformers/models/whisper/feature_extraction_whisper.py:105-133.
"""
from __future__ import annotations

import numpy as np

SAMPLE_RATE = 16000
N_FFT = 400
HOP_LENGTH = 160
CHUNK_LENGTH = 30
N_SAMPLES = CHUNK_LENGTH * SAMPLE_RATE  # 480000
N_FRAMES = N_SAMPLES // HOP_LENGTH  # 3000


def _hz_to_mel_slaney(f):
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, mels)


def _mel_to_hz_slaney(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filters(n_mels: int) -> np.ndarray:
    """(n_mels, 201) f32 slaney-scale, slaney-normalised triangular filterbank.

    Equals librosa.filters.mel(sr=16000, n_fft=400, n_mels=n) which upstream ships as
    assets/mel_filters.npz; cf. HF feature_extraction_whisper.py:95-103.
    """
    assert n_mels in (80, 128), n_mels
    n_freqs = N_FFT // 2 + 1
    fft_freqs = np.linspace(0.0, SAMPLE_RATE / 2, n_freqs)
    mel_pts = np.linspace(_hz_to_mel_slaney(0.0), _hz_to_mel_slaney(SAMPLE_RATE / 2), n_mels + 2)
    hz_pts = _mel_to_hz_slaney(mel_pts)
    fdiff = np.diff(hz_pts)
    ramps = hz_pts[:, None] - fft_freqs[None, :]
    w = np.zeros((n_mels, n_freqs), dtype=np.float64)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        w[i] = np.maximum(0.0, np.minimum(lower, upper))
    enorm = 2.0 / (hz_pts[2 : n_mels + 2] - hz_pts[:n_mels])
    w *= enorm[:, None]
    return w.astype(np.float32)


def hanning(size: int) -> np.ndarray:
    """Periodic Hann: np.hanning(size + 1)[:-1]."""
    return np.hanning(size + 1)[:-1]


def pad_or_trim(array: np.ndarray, length: int = N_SAMPLES, axis: int = -1) -> np.ndarray:
    if array.shape[axis] > length:
        sl = [slice(None)] * array.ndim
        sl[axis] = slice(0, length)
        array = array[tuple(sl)]
    if array.shape[axis] < length:
        pad = [(0, 0)] * array.ndim
        pad[axis] = (0, length - array.shape[axis])
        array = np.pad(array, pad)
    return array


def stft_power(x: np.ndarray, dtype=np.float64) -> np.ndarray:
    """|rfft(hann * frames)|^2 with reflect padding 200, hop 160; returns (t, 201).

    `t = (len + 400 - 400 + 160) // 160` frames (the "noverlap" argument is
    the hop).  The caller drops the last frame.
    """
    x = np.asarray(x, dtype=dtype)
    pad = N_FFT // 2
    assert x.shape[0] > pad, "audio shorter than the reflect pad"
    prefix = x[1 : pad + 1][::-1]
    suffix = x[-(pad + 1) : -1][::-1]
    xp = np.concatenate([prefix, x, suffix])
    t = (xp.shape[0] - N_FFT + HOP_LENGTH) // HOP_LENGTH
    idx = np.arange(N_FFT)[None, :] + HOP_LENGTH * np.arange(t)[:, None]
    frames = xp[idx] * hanning(N_FFT).astype(dtype)[None, :]
    spec = np.fft.rfft(frames, axis=-1)
    return (spec.real**2 + spec.imag**2).astype(dtype)


def log_mel_unclamped(audio: np.ndarray, n_mels: int = 80, padding: int = 0, dtype=np.float64) -> np.ndarray:
    """log10(max(mel, 1e-10)) before the max-8 clamp; (frames, n_mels), time-major."""
    audio = np.asarray(audio, dtype=dtype)
    if padding > 0:
        audio = np.concatenate([audio, np.zeros(padding, dtype=dtype)])
    power = stft_power(audio, dtype)[:-1]  # drop the last frame
    filt = mel_filters(n_mels).astype(dtype)
    mel = power @ filt.T
    return np.log10(np.maximum(mel, 1e-10))


def log_mel_spectrogram(audio: np.ndarray, n_mels: int = 80, padding: int = 0, dtype=np.float64) -> np.ndarray:
    """Whisper log-mel: global max-8 clamp, (x+4)/4; returns f32 (frames, n_mels)."""
    log_spec = log_mel_unclamped(audio, n_mels, padding, dtype)
    log_spec = np.maximum(log_spec, log_spec.max() - 8.0)
    log_spec = (log_spec + 4.0) / 4.0
    return log_spec.astype(np.float32)
