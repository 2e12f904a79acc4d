// Mixed-radix (16 x 25) 400-point complex FFT building blocks for the fused log-mel kernel.
//
// Everything here is __host__ __device__ so the exact arithmetic the kernel performs can be run on
// the CPU by tests (tests/test_logmel_core.py builds csrc/logmel_host_check.cpp around this header).
//
// Decomposition (forward DFT, W_N = exp(-2*pi*i/N)), n = 25*n1 + n2, k = k1 + 16*k2:
//   A[k1][n2]  = sum_{n1<16} z[25*n1 + n2] * W16^(n1*k1)          (25 radix-16 column DFTs)
//   A[k1][n2] *= W400^(n2*k1)                                     (twiddle)
//   X[k1+16*k2]= sum_{n2<25} A[k1][n2] * W25^(n2*k2)              (16 radix-25 row DFTs)
// Two real frames a, b ride one complex transform as z = a + i*b; their power spectra are
//   |Xa[k]|^2 = |Z[k] + conj(Z[400-k])|^2 / 4,   |Xb[k]|^2 = |Z[k] - conj(Z[400-k])|^2 / 4.
#pragma once

#if defined(__CUDACC__)
#define LM_HD __host__ __device__ __forceinline__
#else
#define LM_HD inline
#endif

namespace b200w {
namespace lm {

struct cpx {
  float re, im;
};

LM_HD cpx cadd(cpx a, cpx b) { return {a.re + b.re, a.im + b.im}; }
LM_HD cpx csub(cpx a, cpx b) { return {a.re - b.re, a.im - b.im}; }
LM_HD cpx cmul(cpx a, cpx w) { return {a.re * w.re - a.im * w.im, a.re * w.im + a.im * w.re}; }
// multiply by -i
LM_HD cpx mul_mi(cpx a) { return {a.im, -a.re}; }

// In-place 4-point forward DFT.
LM_HD void dft4(cpx& x0, cpx& x1, cpx& x2, cpx& x3) {
  cpx s02 = cadd(x0, x2), d02 = csub(x0, x2);
  cpx s13 = cadd(x1, x3), d13 = mul_mi(csub(x1, x3));
  x0 = cadd(s02, s13);
  x1 = cadd(d02, d13);
  x2 = csub(s02, s13);
  x3 = csub(d02, d13);
}

// 16-point forward DFT, in place, natural order in and out.
LM_HD void dft16(cpx (&a)[16]) {
  const float C1 = 0.92387953251128674f, S1 = 0.38268343236508977f, R2 = 0.70710678118654752f;
  // W16^m for m = 0..9 (only m = n2*k1 with n2,k1 < 4 are used)
  const cpx W16[10] = {{1.f, 0.f}, {C1, -S1}, {R2, -R2}, {S1, -C1}, {0.f, -1.f},
                       {-S1, -C1}, {-R2, -R2}, {-C1, -S1}, {-1.f, 0.f}, {-C1, S1}};
  cpx b[16];
#pragma unroll
  for (int n2 = 0; n2 < 4; ++n2) {
    cpx t0 = a[n2], t1 = a[4 + n2], t2 = a[8 + n2], t3 = a[12 + n2];
    dft4(t0, t1, t2, t3);
    b[0 * 4 + n2] = t0;
    b[1 * 4 + n2] = (n2 == 0) ? t1 : cmul(t1, W16[n2 * 1]);
    b[2 * 4 + n2] = (n2 == 0) ? t2 : cmul(t2, W16[n2 * 2]);
    b[3 * 4 + n2] = (n2 == 0) ? t3 : cmul(t3, W16[n2 * 3]);
  }
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1) {
    cpx t0 = b[k1 * 4 + 0], t1 = b[k1 * 4 + 1], t2 = b[k1 * 4 + 2], t3 = b[k1 * 4 + 3];
    dft4(t0, t1, t2, t3);
    a[k1 + 0] = t0;
    a[k1 + 4] = t1;
    a[k1 + 8] = t2;
    a[k1 + 12] = t3;
  }
}

// In-place 5-point forward DFT.
LM_HD void dft5(cpx& x0, cpx& x1, cpx& x2, cpx& x3, cpx& x4) {
  const float c1 = 0.30901699437494745f, c2 = -0.80901699437494734f;
  const float s1 = 0.95105651629515353f, s2 = 0.58778525229247325f;
  cpx t1 = cadd(x1, x4), t2 = cadd(x2, x3), t3 = csub(x1, x4), t4 = csub(x2, x3);
  cpx m1 = {x0.re + c1 * t1.re + c2 * t2.re, x0.im + c1 * t1.im + c2 * t2.im};
  cpx m2 = {x0.re + c2 * t1.re + c1 * t2.re, x0.im + c2 * t1.im + c1 * t2.im};
  cpx u1 = {s1 * t3.re + s2 * t4.re, s1 * t3.im + s2 * t4.im};
  cpx u2 = {s2 * t3.re - s1 * t4.re, s2 * t3.im - s1 * t4.im};
  x0 = cadd(x0, cadd(t1, t2));
  cpx mu1 = mul_mi(u1), mu2 = mul_mi(u2);
  x1 = cadd(m1, mu1);
  x4 = csub(m1, mu1);
  x2 = cadd(m2, mu2);
  x3 = csub(m2, mu2);
}

// 25-point forward DFT, in place, natural order in and out.
LM_HD void dft25(cpx (&a)[25]) {
  // W25^m = (cos(2*pi*m/25), -sin(2*pi*m/25)), m = 0..16 (only m = n2*k1 with n2,k1 < 5 are used)
  const cpx W25[17] = {
      {1.f, -0.f},
      {0.96858316112863108f, -0.24868988716485479f},
      {0.87630668004386358f, -0.48175367410171532f},
      {0.72896862742141155f, -0.68454710592868862f},
      {0.53582679497899655f, -0.84432792550201508f},
      {0.30901699437494745f, -0.95105651629515353f},
      {0.062790519529313527f, -0.99802672842827156f},
      {-0.1873813145857246f, -0.98228725072868872f},
      {-0.42577929156507272f, -0.90482705246601947f},
      {-0.63742398974868975f, -0.77051324277578925f},
      {-0.80901699437494734f, -0.58778525229247325f},
      {-0.92977648588825135f, -0.36812455268467814f},
      {-0.99211470131447776f, -0.12533323356430454f},
      {-0.99211470131447788f, 0.12533323356430429f},
      {-0.92977648588825146f, 0.36812455268467792f},
      {-0.80901699437494778f, 0.58778525229247269f},
      {-0.63742398974868952f, 0.77051324277578936f}};
  cpx b[25];
#pragma unroll
  for (int n2 = 0; n2 < 5; ++n2) {
    cpx t0 = a[n2], t1 = a[5 + n2], t2 = a[10 + n2], t3 = a[15 + n2], t4 = a[20 + n2];
    dft5(t0, t1, t2, t3, t4);
    b[0 * 5 + n2] = t0;
    b[1 * 5 + n2] = (n2 == 0) ? t1 : cmul(t1, W25[n2 * 1]);
    b[2 * 5 + n2] = (n2 == 0) ? t2 : cmul(t2, W25[n2 * 2]);
    b[3 * 5 + n2] = (n2 == 0) ? t3 : cmul(t3, W25[n2 * 3]);
    b[4 * 5 + n2] = (n2 == 0) ? t4 : cmul(t4, W25[n2 * 4]);
  }
#pragma unroll
  for (int k1 = 0; k1 < 5; ++k1) {
    cpx t0 = b[k1 * 5 + 0], t1 = b[k1 * 5 + 1], t2 = b[k1 * 5 + 2], t3 = b[k1 * 5 + 3], t4 = b[k1 * 5 + 4];
    dft5(t0, t1, t2, t3, t4);
    a[k1 + 0] = t0;
    a[k1 + 5] = t1;
    a[k1 + 10] = t2;
    a[k1 + 15] = t3;
    a[k1 + 20] = t4;
  }
}

// Index of sample i of the zero-extended, reflect-padded signal:
//   x has n_valid real samples followed by zeros up to n_total; positions i < 0 and i >= n_total
//   reflect without repeating the edge.  Returns -1 when the sample is a (padding) zero.
LM_HD long long reflect_index(long long i, long long n_valid, long long n_total) {
  if (i < 0) i = -i;
  if (i >= n_total) i = 2 * (n_total - 1) - i;
  if (i < 0 || i >= n_valid) return -1;
  return i;
}

}  // namespace lm
}  // namespace b200w
