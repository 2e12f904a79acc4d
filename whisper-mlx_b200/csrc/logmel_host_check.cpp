// CPU harness around logmel_core.h: runs the exact radix-16 x radix-25 arithmetic of logmel.cu's
// phases 1-4 for one frame pair so tests can pin the FFT math without a GPU (tests/test_logmel_core.py).
#include <cmath>
#include <vector>

#include "logmel_core.h"

using b200w::lm::cpx;

extern "C" void lm_host_pair_power(const float* fa, const float* fb, const float* hann, const float* tw400,
                                   float* pa, float* pb) {
  std::vector<cpx> work(400);
  for (int n2 = 0; n2 < 25; ++n2) {
    cpx a[16];
    for (int n1 = 0; n1 < 16; ++n1) {
      int n = 25 * n1 + n2;
      a[n1].re = hann[n] * fa[n];
      a[n1].im = hann[n] * fb[n];
    }
    b200w::lm::dft16(a);
    for (int k1 = 0; k1 < 16; ++k1) {
      cpx t{tw400[2 * (n2 * 16 + k1)], tw400[2 * (n2 * 16 + k1) + 1]};
      work[k1 * 25 + n2] = b200w::lm::cmul(a[k1], t);
    }
  }
  std::vector<cpx> spec(400);
  for (int k1 = 0; k1 < 16; ++k1) {
    cpx a[25];
    for (int n2 = 0; n2 < 25; ++n2) a[n2] = work[k1 * 25 + n2];
    b200w::lm::dft25(a);
    for (int k2 = 0; k2 < 25; ++k2) spec[k1 + 16 * k2] = a[k2];
  }
  for (int k = 0; k <= 200; ++k) {
    cpx zk = spec[k], zn = spec[(400 - k) % 400];
    float ar = zk.re + zn.re, ai = zk.im - zn.im, br = zk.re - zn.re, bi = zk.im + zn.im;
    pa[k] = 0.25f * (ar * ar + ai * ai);
    pb[k] = 0.25f * (br * br + bi * bi);
  }
}

extern "C" long long lm_host_reflect_index(long long i, long long n_valid, long long n_total) {
  return b200w::lm::reflect_index(i, n_valid, n_total);
}
