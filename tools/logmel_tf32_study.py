"""CPU study (NumPy emulation): could the 400-point transform of the log-mel front-end run on tcgen05 in tf32?

The transform is emulated exactly as a tensor-core version would compute it -- 16 x 25 two-stage DFT as real GEMMs,
operands rounded to tf32 (10 explicit mantissa bits), products exact, fp32 accumulation -- in three variants:
  tf32x1 : data and DFT matrices rounded to tf32 once                       (1 MMA per product)
  tf32x3 : data = hi + lo, matrix = hi + lo, terms hi*hi + lo*hi + hi*lo    (3 MMAs per product)
  fp32   : the shipped arithmetic (FP32 FFT)
and the resulting normalised log-mel is compared with a float64 reference on the signal kinds of the parity tests.
The tolerance of the GPU parity tests is 1e-4 (1e-3 for tones / clicks with > 60 dB of inter-frame dynamic range).

    python tools/logmel_tf32_study.py  -> JSON
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools import synth  # noqa: E402
from whisper_mlx_b200.audio import mel_filters  # noqa: E402


def tf32(x):
    """Round-to-nearest-even to 10 explicit mantissa bits."""
    b = np.asarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    b = (b + 0xFFF + ((b >> 13) & 1)) & ~np.uint64(0x1FFF)
    return b.astype(np.uint32).view(np.float32)


def gemm(a, w, mode):
    """a (..., K) x w (K, N) with the given operand precision; fp32 accumulation (emulated by float32 matmul of
    exactly representable products: each tf32 x tf32 product is exact in fp32)."""
    a = a.astype(np.float32)
    w = w.astype(np.float32)
    if mode == "fp32":
        return a @ w
    ah, wh = tf32(a), tf32(w)
    if mode == "tf32x1":
        return ah @ wh
    al, wl = tf32(a - ah), tf32(w - wh)
    return (ah @ wh) + (al @ wh) + (ah @ wl)


def frames_of(x, n_frames):
    pad = np.pad(x, (200, 200), mode="reflect")
    idx = np.arange(400)[None, :] + 160 * np.arange(n_frames)[:, None]
    return pad[idx]


def dft_mats():
    n1, k1 = np.meshgrid(np.arange(16), np.arange(16), indexing="ij")
    w16 = np.exp(-2j * np.pi * n1 * k1 / 16)
    n2, k2 = np.meshgrid(np.arange(25), np.arange(25), indexing="ij")
    w25 = np.exp(-2j * np.pi * n2 * k2 / 25)
    tw = np.exp(-2j * np.pi * np.arange(25)[:, None] * np.arange(16)[None, :] / 400)  # [n2][k1]

    def real_form(w):  # complex (K, N) -> real (2K, 2N): [re | im] rows and columns
        k, n = w.shape
        m = np.zeros((2 * k, 2 * n))
        m[:k, :n], m[:k, n:], m[k:, :n], m[k:, n:] = w.real, w.imag, -w.imag, w.real
        return m
    return real_form(w16), real_form(w25), tw


def power_spectrum(fr, mode):
    """fr (F, 400) windowed real frames -> (F, 201) power via the 16 x 25 decomposition n = 25 n1 + n2, k = k1 + 16 k2."""
    m16, m25, tw = dft_mats()
    F = fr.shape[0]
    z = fr.reshape(F, 16, 25).transpose(0, 2, 1)                      # [F][n2][n1] (real input)
    a = np.concatenate([z, np.zeros_like(z)], axis=-1)                # [re | im] over n1
    y = gemm(a, m16, mode)                                            # [F][n2][k1 re | k1 im]
    yc = (y[..., :16] + 1j * y[..., 16:]).astype(np.complex64)
    yc = (yc * tw[None].astype(np.complex64)).astype(np.complex64)     # twiddle in fp32
    b = yc.transpose(0, 2, 1)                                          # [F][k1][n2]
    b = np.concatenate([b.real, b.imag], axis=-1).astype(np.float32)
    x = gemm(b, m25, mode)                                            # [F][k1][k2 re | k2 im]
    xc = x[..., :25] + 1j * x[..., 25:]
    spec = xc.transpose(0, 2, 1).reshape(F, 400)                       # k = k1 + 16 k2
    return (spec.real.astype(np.float32) ** 2 + spec.imag.astype(np.float32) ** 2)[:, :201]


def logmel_from_power(p, n_mels):
    mel = p.astype(np.float32) @ np.asarray(mel_filters(n_mels), dtype=np.float32).T
    lm = np.log10(np.maximum(mel, 1e-10))
    lm = np.maximum(lm, lm.max() - 8.0)
    return (lm + 4.0) / 4.0


def main():
    n = 16000 * 10
    hann = (0.5 - 0.5 * np.cos(2 * np.pi * np.arange(400) / 400)).astype(np.float32)
    out = {}
    for kind in ("noise", "tones", "click", "clip", "speech"):
        x = synth.make_audio(kind, n, 3).astype(np.float32)
        fr = frames_of(x, n // 160) * hann[None]
        ref_p = np.abs(np.fft.rfft(fr.astype(np.float64), axis=-1)) ** 2
        ref = logmel_from_power(ref_p, 128)
        row = {}
        for mode in ("fp32", "tf32x3", "tf32x1"):
            got = logmel_from_power(power_spectrum(fr, mode), 128)
            e = np.abs(got - ref)
            row[mode] = {"max": float(e.max()), "frac_above_1e-4": float((e > 1e-4).mean()), "frac_above_1e-3": float((e > 1e-3).mean())}
        out[kind] = row
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
