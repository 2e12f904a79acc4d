// K1: fused log-mel front-end (reference: mlx_whisper/audio.py::log_mel_spectrogram, reached from
// /root/reference/run:3; restated in SURVEY.md A.1).
//
// One CTA turns kFrames consecutive STFT frames of one audio into log10(mel) rows:
//   samples (coalesced float4 loads, reflect / zero-extension only in edge tiles) -> shared memory tile
//   with halo (each sample is reused by its 2.5 overlapping frames) -> Hann window -> 400-point FP32 FFT
//   as 16 x 25 mixed radix in shared memory, two real frames per complex transform -> |X|^2 -> sparse
//   mel filterbank (<= 2 non-zeros per frequency row) -> log10(max(., 1e-10)) -> unclamped store
//   + per-audio running max (one atomic per CTA).
// K1b kernels apply max(x, gmax-8), (x+4)/4: in place (f32 API result) or fused into the bf16
// window gather that feeds the conv stem.
#include "common.cuh"
#include "kernels.h"
#include "logmel_core.h"

namespace b200w {

constexpr int kNfft = 400;
constexpr int kHop = 160;
constexpr int kBins = 201;
constexpr int kFrames = 16;             // frames per CTA
constexpr int kPairs = kFrames / 2;     // complex transforms per CTA
constexpr int kTile = (kFrames - 1) * kHop + kNfft;  // 2800 samples incl. halo
constexpr int kLmThreads = 128;
constexpr int kPowStride = 208;         // floats per power spectrum row

struct LogmelTables {
  const float* hann;      // [400] periodic Hann
  const float2* tw400;    // [25][16] W400^(n2*k1)
  const int* mel_lo;      // [n_mels] first bin
  const int* mel_cnt;     // [n_mels] bins in the contiguous support
  const int* mel_off;     // [n_mels] offset into mel_w
  const float* mel_w;     // flattened non-zero weights
};

__global__ void __launch_bounds__(kLmThreads)
logmel_kernel(const float* __restrict__ pcm, long long audio_stride, long long n_valid, long long n_total,
              int n_frames, int n_mels, LogmelTables tb, float* __restrict__ out, float* __restrict__ gmax) {
  extern __shared__ __align__(16) unsigned char lm_smem[];
  // region 0: sample tile (phase 0-1), later the power spectra (phase 4-5)
  float* s_samples = reinterpret_cast<float*>(lm_smem);
  float* s_power = reinterpret_cast<float*>(lm_smem);
  constexpr int kRegion0 = (kPairs * 2 * kPowStride > kTile ? kPairs * 2 * kPowStride : kTile);
  lm::cpx* s_work = reinterpret_cast<lm::cpx*>(lm_smem + sizeof(float) * kRegion0);  // [kPairs][400]
  __shared__ float s_red[kLmThreads / 32];

  const int tid = threadIdx.x;
  const int audio = blockIdx.y;
  const int f0 = blockIdx.x * kFrames;
  const float* x = pcm + (long long)audio * audio_stride;
  const long long base = (long long)f0 * kHop - kNfft / 2;  // signal index of tile sample 0

  // ---- phase 0: stage samples ------------------------------------------------------------------
  const bool interior = (base >= 0) && (base + kTile <= n_valid) && ((audio_stride & 3) == 0) &&
                        ((reinterpret_cast<uintptr_t>(pcm) & 15) == 0);
  if (interior) {
    const float4* src = reinterpret_cast<const float4*>(x + base);  // base % 8 == 0
    float4* dst = reinterpret_cast<float4*>(s_samples);
#pragma unroll 2
    for (int i = tid; i < kTile / 4; i += kLmThreads) dst[i] = __ldg(src + i);
  } else {
    for (int i = tid; i < kTile; i += kLmThreads) {
      long long j = lm::reflect_index(base + i, n_valid, n_total);
      s_samples[i] = (j >= 0) ? __ldg(x + j) : 0.0f;
    }
  }
  __syncthreads();

  // ---- phase 1: 25 column DFT-16 per pair + W400 twiddle ------------------------------------------
  for (int item = tid; item < kPairs * 25; item += kLmThreads) {
    const int p = item / 25, n2 = item - p * 25;
    const float* fa = s_samples + (2 * p) * kHop;
    const float* fb = fa + kHop;
    lm::cpx a[16];
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) {
      const int n = 25 * n1 + n2;
      const float w = __ldg(tb.hann + n);
      a[n1].re = w * fa[n];
      a[n1].im = w * fb[n];
    }
    lm::dft16(a);
    const float4* tw = reinterpret_cast<const float4*>(tb.tw400 + n2 * 16);
    lm::cpx* dst = s_work + p * kNfft + n2;
#pragma unroll
    for (int k1 = 0; k1 < 16; k1 += 2) {
      const float4 t = __ldg(tw + (k1 >> 1));
      dst[(k1 + 0) * 25] = lm::cmul(a[k1 + 0], lm::cpx{t.x, t.y});
      dst[(k1 + 1) * 25] = lm::cmul(a[k1 + 1], lm::cpx{t.z, t.w});
    }
  }
  __syncthreads();

  // ---- phase 2/3: 16 row DFT-25 per pair (read rows, sync, scatter to natural order) ---------------
  {
    static_assert(kPairs * 16 == kLmThreads, "one DFT-25 per thread");
    const int p = tid >> 4, k1 = tid & 15;
    lm::cpx* row = s_work + p * kNfft;
    lm::cpx a[25];
#pragma unroll
    for (int n2 = 0; n2 < 25; ++n2) a[n2] = row[k1 * 25 + n2];
    lm::dft25(a);
    __syncthreads();
#pragma unroll
    for (int k2 = 0; k2 < 25; ++k2) row[k1 + 16 * k2] = a[k2];
  }
  __syncthreads();

  // ---- phase 4: split the two real spectra, power ----------------------------------------------
  for (int item = tid; item < kPairs * kBins; item += kLmThreads) {
    const int p = item / kBins, k = item - p * kBins;
    const lm::cpx zk = s_work[p * kNfft + k];
    const lm::cpx zn = s_work[p * kNfft + ((kNfft - k) % kNfft)];
    const float ar = zk.re + zn.re, ai = zk.im - zn.im;
    const float br = zk.re - zn.re, bi = zk.im + zn.im;
    s_power[(2 * p) * kPowStride + k] = 0.25f * (ar * ar + ai * ai);
    s_power[(2 * p + 1) * kPowStride + k] = 0.25f * (br * br + bi * bi);
  }
  __syncthreads();

  // ---- phase 5: sparse mel, log10, store, running max ----------------------------------------------
  float vmax = -INFINITY;
  const int valid_frames = min(kFrames, n_frames - f0);
  float* o = out + ((long long)audio * n_frames + f0) * n_mels;
  for (int item = tid; item < valid_frames * n_mels; item += kLmThreads) {
    const int f = item / n_mels, m = item - f * n_mels;
    const int lo = __ldg(tb.mel_lo + m), cnt = __ldg(tb.mel_cnt + m);
    const float* w = tb.mel_w + __ldg(tb.mel_off + m);
    const float* pw = s_power + f * kPowStride + lo;
    float acc = 0.0f;
    for (int j = 0; j < cnt; ++j) acc = fmaf(__ldg(w + j), pw[j], acc);
    const float v = 0.30102999566398120f * __log2f(fmaxf(acc, 1e-10f));
    o[item] = v;
    vmax = fmaxf(vmax, v);
  }
  vmax = warp_max(vmax);
  if ((tid & 31) == 0) s_red[tid >> 5] = vmax;
  __syncthreads();
  if (tid == 0) {
    float m = s_red[0];
#pragma unroll
    for (int i = 1; i < kLmThreads / 32; ++i) m = fmaxf(m, s_red[i]);
    if (m > -INFINITY) atomic_max_float(gmax + audio, m);
  }
}

__global__ void fill_f32_kernel(float* p, float v, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// K1b (f32 API result): x <- (max(x, gmax[audio] - 8) + 4) / 4, in place.
__global__ void logmel_finalize_kernel(float* __restrict__ x, const float* __restrict__ gmax,
                                       long long per_audio, long long total) {
  long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= total) return;
  if ((per_audio & 3) == 0 && i + 3 < total) {
    const float fl = __ldg(gmax + i / per_audio) - 8.0f;
    float4 v = *reinterpret_cast<float4*>(x + i);
    v.x = (fmaxf(v.x, fl) + 4.0f) * 0.25f;
    v.y = (fmaxf(v.y, fl) + 4.0f) * 0.25f;
    v.z = (fmaxf(v.z, fl) + 4.0f) * 0.25f;
    v.w = (fmaxf(v.w, fl) + 4.0f) * 0.25f;
    *reinterpret_cast<float4*>(x + i) = v;
  } else {
    for (int j = 0; j < 4 && i + j < total; ++j) {
      const float fl = __ldg(gmax + (i + j) / per_audio) - 8.0f;
      x[i + j] = (fmaxf(x[i + j], fl) + 4.0f) * 0.25f;
    }
  }
}

// K1b fused with the window gather feeding the conv stem: for window w take `size[w]` frames starting
// at row `row0[w]` of the unclamped log-mel, clamp with gmax[gidx[w]], scale, cast to bf16 and write
// rows 1..3000 of a (3002, n_mels) slab whose rows 0 and 3001 are the conv's zero padding.  Frames past
// size[w] are 0.0 (pad_or_trim on the *normalised* mel).
__global__ void mel_window_kernel(const float* __restrict__ mel, const float* __restrict__ gmax,
                                  const long long* __restrict__ row0, const int* __restrict__ size,
                                  const int* __restrict__ gidx, int n_mels, __nv_bfloat16* __restrict__ dst) {
  const int w = blockIdx.y;
  const int per_win = 3002 * n_mels;
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) * 2;
  if (i >= per_win) return;
  const int r = i / n_mels;  // n_mels is even: both elements are in the same row
  float v0 = 0.0f, v1 = 0.0f;
  const int t = r - 1;
  if (t >= 0 && t < size[w]) {
    const float2 s = *reinterpret_cast<const float2*>(mel + (row0[w] + t) * n_mels + (i - r * n_mels));
    if (gmax != nullptr) {
      const float fl = __ldg(gmax + gidx[w]) - 8.0f;
      v0 = (fmaxf(s.x, fl) + 4.0f) * 0.25f;
      v1 = (fmaxf(s.y, fl) + 4.0f) * 0.25f;
    } else {  // rows are already normalised log-mel: only the bf16 cast and the padding remain
      v0 = s.x;
      v1 = s.y;
    }
  }
  *reinterpret_cast<uint32_t*>(dst + (long long)w * per_win + i) = pack_bf16x2(v0, v1);
}

// ---------------------------------------------------------------------------------------------- host
static size_t logmel_smem_bytes() {
  constexpr int kRegion0 = (kPairs * 2 * kPowStride > kTile ? kPairs * 2 * kPowStride : kTile);
  return sizeof(float) * kRegion0 + sizeof(lm::cpx) * kPairs * kNfft;
}

int init_logmel() {
  static bool done = false;
  if (done) return kOk;
  B200W_CUDA_OK(cudaFuncSetAttribute(logmel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)logmel_smem_bytes()));
  done = true;
  return kOk;
}

int launch_logmel(const float* pcm, int n_audio, long long audio_stride, long long n_valid, long long n_total,
                  int n_mels, const float* hann, const float* tw400, const int* mel_lo, const int* mel_cnt,
                  const int* mel_off, const float* mel_w, float* out_unclamped, float* gmax, cudaStream_t stream) {
  B200W_CHECK_ARG(n_audio > 0 && n_valid > kNfft / 2 && n_total >= n_valid, "logmel: bad sizes");
  B200W_CHECK_ARG(n_mels == 80 || n_mels == 128, "logmel: n_mels must be 80 or 128, got %d", n_mels);
  const long long n_frames_ll = n_total / kHop;
  B200W_CHECK_ARG(n_frames_ll > 0 && n_frames_ll < (1ll << 31) / 128, "logmel: frame count out of range");
  const int n_frames = (int)n_frames_ll;
  const size_t smem = logmel_smem_bytes();
  B200W_TRY(init_logmel());
  fill_f32_kernel<<<ceil_div(n_audio, 256), 256, 0, stream>>>(gmax, -INFINITY, n_audio);
  B200W_LAUNCH_OK();
  LogmelTables tb{hann, reinterpret_cast<const float2*>(tw400), mel_lo, mel_cnt, mel_off, mel_w};
  dim3 grid(ceil_div(n_frames, kFrames), n_audio);
  B200W_CHECK_ARG(n_audio <= 65535, "logmel: at most 65535 audios per call");
  logmel_kernel<<<grid, kLmThreads, smem, stream>>>(pcm, audio_stride, n_valid, n_total, n_frames, n_mels, tb,
                                                    out_unclamped, gmax);
  B200W_LAUNCH_OK();
  count_launch(2);
  return kOk;
}

int launch_logmel_finalize(float* x, const float* gmax, int n_audio, long long per_audio, cudaStream_t stream) {
  const long long total = per_audio * n_audio;
  const long long nthreads = ceil_div_ll(total, 4);
  logmel_finalize_kernel<<<(unsigned)ceil_div_ll(nthreads, 256), 256, 0, stream>>>(x, gmax, per_audio, total);
  B200W_LAUNCH_OK();
  count_launch();
  return kOk;
}

int launch_mel_windows(const float* mel, const float* gmax, const long long* row0, const int* size, const int* gidx,
                       int n_windows, int n_mels, __nv_bfloat16* dst, cudaStream_t stream) {
  B200W_CHECK_ARG(n_mels % 2 == 0 && n_windows > 0 && n_windows <= 65535, "mel_windows: bad sizes");
  dim3 grid(ceil_div(3002 * n_mels / 2, 256), n_windows);
  mel_window_kernel<<<grid, 256, 0, stream>>>(mel, gmax, row0, size, gidx, n_mels, dst);
  B200W_LAUNCH_OK();
  count_launch();
  return kOk;
}

}  // namespace b200w
