"""The exact radix-16 x radix-25 FFT arithmetic of csrc/logmel.cu (shared header logmel_core.h), run on
the CPU through csrc/logmel_host_check.cpp and pinned against numpy's FFT."""
import ctypes
import os

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host(built_lib):
    return ctypes.CDLL(os.path.join(REPO, "whisper-mlx_b200", "csrc", "liblogmel_host_check.so"))


def _tables():
    hann = np.hanning(401)[:-1].astype(np.float32)
    n2, k1 = np.arange(25)[:, None], np.arange(16)[None, :]
    ang = -2 * np.pi * (n2 * k1) / 400
    return hann, np.stack([np.cos(ang), np.sin(ang)], -1).astype(np.float32).copy()


def _pair_power(host, fa, fb):
    hann, tw = _tables()
    pa, pb = np.zeros(201, np.float32), np.zeros(201, np.float32)
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    host.lm_host_pair_power(P(fa), P(fb), P(hann), P(tw), P(pa), P(pb))
    return pa, pb


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_pair_fft_matches_numpy(host, seed):
    rng = np.random.default_rng(seed)
    fa, fb = rng.standard_normal(400).astype(np.float32), rng.standard_normal(400).astype(np.float32)
    pa, pb = _pair_power(host, fa, fb)
    w = np.hanning(401)[:-1]
    ra = np.abs(np.fft.rfft(fa.astype(np.float64) * w)) ** 2
    rb = np.abs(np.fft.rfft(fb.astype(np.float64) * w)) ** 2
    assert np.abs(pa - ra).max() <= 2e-6 * ra.max()
    assert np.abs(pb - rb).max() <= 2e-6 * rb.max()


def test_impulse_and_dc(host):
    imp = np.zeros(400, np.float32)
    imp[200] = 1.0  # hann[200] == 1 -> flat spectrum of power 1
    dc = np.ones(400, np.float32)
    pa, pb = _pair_power(host, imp, dc)
    assert np.allclose(pa, 1.0, atol=1e-5)
    w = np.hanning(401)[:-1]
    assert np.allclose(pb, np.abs(np.fft.rfft(w)) ** 2, atol=1e-2 * 200 ** 2 * 1e-4)


def test_reflect_index(host):
    host.lm_host_reflect_index.restype = ctypes.c_longlong
    host.lm_host_reflect_index.argtypes = [ctypes.c_longlong] * 3
    n_valid, n_total = 1000, 1600
    x = np.arange(n_valid, dtype=np.float64) + 1
    ext = np.concatenate([x, np.zeros(n_total - n_valid)])
    padded = np.pad(ext, 200, mode="reflect")
    for j in range(len(padded)):
        i = host.lm_host_reflect_index(j - 200, n_valid, n_total)
        got = 0.0 if i < 0 else x[i]
        assert got == padded[j], (j, i)
