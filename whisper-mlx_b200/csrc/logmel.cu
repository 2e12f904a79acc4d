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
#include "mel_tables.h"

namespace b200w {

constexpr int kNfft = 400;
constexpr int kHop = 160;
constexpr int kBins = 201;
constexpr int kFrames = 32;             // frames per CTA
constexpr int kPairs = kFrames / 2;     // complex transforms per CTA (two real frames each)
constexpr int kTile = (kFrames - 1) * kHop + kNfft;  // 5360 samples incl. halo
constexpr int kLmThreads = 256;
constexpr int kPowStride = 203;         // floats per power-spectrum row; odd => conflict-free lane-per-frame reads
constexpr int kOutStride = 129;         // floats per staged output row (n_mels <= 128), odd for the same reason
constexpr int kPairStride = 409;        // complex elements between the FFT work areas of two pairs (400 + bank skew)
constexpr int kRegion0Floats = (kFrames * kPowStride > kTile ? kFrames * kPowStride : kTile);

struct LogmelTables {
  const float* hann;      // [400] periodic Hann
  const float2* tw400;    // [25][16] W400^(n2*k1)
};

// Sparse mel stage for the mels m = W, W + 8, ... of one warp; lane == frame.  The filterbank is a compile-time
// table (mel_tables.h): after unrolling every weight is an FFMA immediate and every power-spectrum read an LDS
// with an immediate offset -- no table loads, no loop overhead, no divergence.
template <int NM, int W>
__device__ __forceinline__ void mel_for_warp(const float* __restrict__ pw, float* __restrict__ so, float& vmax) {
  using Bank = lm::MelBank<NM>;
#pragma unroll
  for (int i = 0; i < (NM - W + 7) / 8; ++i) {
    const int m = W + 8 * i;
    float acc = 0.0f;
#pragma unroll
    for (int j = 0; j < Bank::kMaxCnt; ++j)
      if (j < Bank::cnt(m)) acc = fmaf(0.25f * Bank::w(Bank::off(m) + j), pw[Bank::lo(m) + j], acc);  // (the /4 of the split)
    const float v = 0.30102999566398120f * __log2f(fmaxf(acc, 1e-10f));
    so[m] = v;
    vmax = fmaxf(vmax, v);
  }
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Persistent: 2 CTAs per SM walk the (audio, 32-frame tile) list; the next tile's samples stream into the second
// sample buffer with cp.async while the current tile is transformed.
// Shared memory plan (102.6 KB, two CTAs per SM):
//   buffers 0/1 : sample tile with halo (phases 0-1) -> power spectra [32][203] of the same tile (phases 3-5);
//                 the other buffer receives the next tile's samples meanwhile
//   work        : FFT work [16 pairs][409] complex   -> staged log-mel rows [32][129] (phase 5)
// PcmT = float (samples in [-1, 1]) or int16_t (s16le PCM as ffmpeg / a WAV file delivers it: the /32768 scaling of
// load_audio is applied while the tile is staged, so the int16 -> f32 pass over the file and half of the
// host-to-device bytes disappear; the shared-memory tile and everything after it are identical).
template <int NM, typename PcmT>
__global__ void __launch_bounds__(kLmThreads, 2)
logmel_kernel(const PcmT* __restrict__ pcm, long long audio_stride, long long n_valid, long long n_total,
              int n_frames, int tiles_per_audio, int n_tiles, LogmelTables tb, float* __restrict__ out,
              float* __restrict__ gmax, int* __restrict__ done_tiles) {
  constexpr int n_mels = NM;
  extern __shared__ __align__(16) unsigned char lm_smem[];
  float* buf_cur = reinterpret_cast<float*>(lm_smem);
  float* buf_nxt = buf_cur + kRegion0Floats;
  lm::cpx* s_work = reinterpret_cast<lm::cpx*>(lm_smem + 2 * sizeof(float) * kRegion0Floats);
  float* s_out = reinterpret_cast<float*>(s_work);
  __shared__ float s_red[kLmThreads / 32];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr bool kPcm16 = sizeof(PcmT) == 2;
  const bool can_vec = ((audio_stride & (kPcm16 ? 7 : 3)) == 0) && ((reinterpret_cast<uintptr_t>(pcm) & 15) == 0);

  // phase 0 of a tile: interior tiles stream in asynchronously, edge tiles (reflect / zero extension) synchronously
  auto stage_tile = [&](float* dst, int t) {
    const int a = t / tiles_per_audio;
    const long long base = (long long)(t - a * tiles_per_audio) * kFrames * kHop - kNfft / 2;
    const PcmT* x = pcm + (long long)a * audio_stride;
    if (can_vec && base >= 0 && base + kTile <= n_valid) {
      if constexpr (kPcm16) {
        static_assert(kTile % 8 == 0, "eight samples per 16-byte load");
        for (int i = tid; i < kTile / 8; i += kLmThreads) {  // base % 8 == 0
          const uint4 v = __ldg(reinterpret_cast<const uint4*>(x + base) + i);
          const uint32_t w[4] = {v.x, v.y, v.z, v.w};
          float f[8];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            f[2 * k] = (float)(short)(w[k] & 0xffffu) * (1.0f / 32768.0f);
            f[2 * k + 1] = (float)(short)(w[k] >> 16) * (1.0f / 32768.0f);
          }
          *reinterpret_cast<float4*>(dst + 8 * i) = make_float4(f[0], f[1], f[2], f[3]);
          *reinterpret_cast<float4*>(dst + 8 * i + 4) = make_float4(f[4], f[5], f[6], f[7]);
        }
      } else {
        for (int i = tid; i < kTile / 4; i += kLmThreads) cp_async16(dst + 4 * i, x + base + 4 * i);  // base % 8 == 0
      }
    } else {
      for (int i = tid; i < kTile; i += kLmThreads) {
        const long long j = lm::reflect_index(base + i, n_valid, n_total);
        if constexpr (kPcm16) dst[i] = (j >= 0) ? (float)__ldg(x + j) * (1.0f / 32768.0f) : 0.0f;
        else dst[i] = (j >= 0) ? __ldg(x + j) : 0.0f;
      }
    }
    cp_async_commit();
  };

  // K1b fused (done_tiles != null; cooperative launch, every CTA resident): the clamp needs the maximum over the WHOLE
  // audio, so a tile can only be normalised once all tiles of its audio are stored.  A CTA therefore normalises the tile
  // it produced two rounds earlier -- by then its audio is complete (the wait practically never spins) and the 16 KB
  // are still in L2: no second pass over HBM.  done_tiles[a] counts the stored tiles of audio a.  The work hides in
  // the tile loop: warp 7 (idle in phase 1) waits for the count and fetches the clamp floor; after phase 3 the rows
  // stream back into the free part of the FFT work area with cp.async; after the row stores every thread clamps the
  // chunks it fetched itself and writes them out.
  __shared__ float s_floor;
  float* s_back = reinterpret_cast<float*>(s_work) + kFrames * kOutStride;  // behind the staged rows: 16.5 KB .. 33 KB
  static_assert((kFrames * kOutStride * sizeof(float)) % 16 == 0, "cp.async destination alignment");
  static_assert(sizeof(float) * (kFrames * kOutStride + kFrames * 128) <= sizeof(lm::cpx) * kPairs * kPairStride, "fits the work area");
  auto tile_rows = [&](int t, int& n_floats) -> float* {
    const int a = t / tiles_per_audio;
    const int tf0 = (t - a * tiles_per_audio) * kFrames;
    n_floats = min(kFrames, n_frames - tf0) * NM;  // contiguous, 16-byte aligned
    return out + ((long long)a * n_frames + tf0) * NM;
  };
  auto wait_floor = [&](int t) {  // one thread: the audio of tile t is complete -> its clamp floor
    const int a = t / tiles_per_audio;
    unsigned int spins = 0;
    while (*reinterpret_cast<volatile int*>(done_tiles + a) < tiles_per_audio) {
      if (++spins > (1u << 28)) __trap();  // a lost CTA must not hang the GPU
    }
    __threadfence();
    s_floor = __ldcg(gmax + a) - 8.0f;
  };
  auto clamp4 = [](float4 v, float fl) {
    v.x = (fmaxf(v.x, fl) + 4.0f) * 0.25f;
    v.y = (fmaxf(v.y, fl) + 4.0f) * 0.25f;
    v.z = (fmaxf(v.z, fl) + 4.0f) * 0.25f;
    v.w = (fmaxf(v.w, fl) + 4.0f) * 0.25f;
    return v;
  };
  auto finalize_tile = [&](int t) {  // unpipelined form, for the tiles left over when the loop ends
    if (tid == 0) wait_floor(t);
    __syncthreads();
    const float fl = s_floor;
    int n;
    float4* o4 = reinterpret_cast<float4*>(tile_rows(t, n));
    for (int i = tid; i < n / 4; i += kLmThreads) o4[i] = clamp4(__ldcg(o4 + i), fl);
    __syncthreads();  // s_floor is rewritten by the next call
  };
  int pend0 = -1, pend1 = -1;  // this CTA's stored tiles that are not normalised yet (newest, older)

  int tile = blockIdx.x;
  if (tile < n_tiles) stage_tile(buf_cur, tile);
  for (; tile < n_tiles; tile += gridDim.x) {
  const int audio = tile / tiles_per_audio;
  const int f0 = (tile - audio * tiles_per_audio) * kFrames;
  float* s_samples = buf_cur;
  float* s_power = buf_cur;
  // thread t < 200 owns column n2 = t % 25 for pairs t / 25 and t / 25 + 8 (two each: no warp waits at the barrier for
  // a straggler with an extra pair).  Its window and twiddle values are fetched BEFORE the wait for the samples: with
  // ~20 KB of L1 left beside 205 KB of shared memory these table reads often come from L2, and their latency then
  // hides behind the cp.async wait and the barrier instead of stalling the first butterflies (ncu r02: long-scoreboard
  // stalls were 18 % of the samples of this phase).
  const int n2 = tid % 25, g = tid / 25;
  float hw[16];
  float4 tw[8];
  if (tid < 200) {
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) hw[n1] = __ldg(tb.hann + 25 * n1 + n2);
    const float4* twp = reinterpret_cast<const float4*>(tb.tw400 + n2 * 16);
#pragma unroll
    for (int i = 0; i < 8; ++i) tw[i] = __ldg(twp + i);
  }
  cp_async_wait_all();
  __syncthreads();  // samples of this tile have landed; the previous tile's staged rows / power are no longer read

  // ---- phase 1: 25 column DFT-16 per pair + W400 twiddle ------------------------------------------
  if (done_tiles != nullptr && pend1 >= 0 && tid == kLmThreads - 32) wait_floor(pend1);  // (warp 7 has no column work)
  if (tid < 200) {
#pragma unroll 1
    for (int p = g; p < kPairs; p += 8) {
      const float* fa = s_samples + (2 * p) * kHop + n2;
      const float* fb = fa + kHop;
      lm::cpx a[16];
#pragma unroll
      for (int n1 = 0; n1 < 16; ++n1) {
        a[n1].re = hw[n1] * fa[25 * n1];
        a[n1].im = hw[n1] * fb[25 * n1];
      }
      lm::dft16(a);
      lm::cpx* dst = s_work + p * kPairStride + n2;
#pragma unroll
      for (int k1 = 0; k1 < 16; k1 += 2) {
        dst[(k1 + 0) * 25] = lm::cmul(a[k1 + 0], lm::cpx{tw[k1 >> 1].x, tw[k1 >> 1].y});
        dst[(k1 + 1) * 25] = lm::cmul(a[k1 + 1], lm::cpx{tw[k1 >> 1].z, tw[k1 >> 1].w});
      }
    }
  }
  __syncthreads();
  if (tile + (int)gridDim.x < n_tiles) stage_tile(buf_nxt, tile + gridDim.x);  // overlaps phases 3-5

  // ---- phase 3: 16 row DFT-25 per pair; split the two real spectra in registers -------------------------
  // thread (p, k1) ends with Z[k1 + 16*k2], k2 < 25.  The conjugate partner Z[400 - k] of its bins k <= 200 lives
  // in lane (16 - k1) % 16 of the same half-warp at register 24 - k2 (k1 == 0: own register 25 - k2), so
  // |Xa|^2, |Xb|^2 need one shuffle pair per bin instead of a round trip through shared memory.
  {
    static_assert(kPairs * 16 == kLmThreads, "one DFT-25 per thread");
    const int p = tid >> 4, k1 = tid & 15;
    const lm::cpx* row = s_work + p * kPairStride + k1 * 25;
    lm::cpx a[25];
#pragma unroll
    for (int n2 = 0; n2 < 25; ++n2) a[n2] = row[n2];
    lm::dft25(a);
    const int partner = (lane & 16) | ((16 - k1) & 15);
    float* pa = s_power + (2 * p) * kPowStride + k1;
    float* pb = pa + kPowStride;
#pragma unroll
    for (int k2 = 0; k2 <= 12; ++k2) {
      const float sre = __shfl_sync(0xffffffffu, a[24 - k2].re, partner);
      const float sim = __shfl_sync(0xffffffffu, a[24 - k2].im, partner);
      const lm::cpx own = a[(25 - k2) % 25];
      const float zr = (k1 == 0) ? own.re : sre, zi = (k1 == 0) ? own.im : sim;
      const float ar = a[k2].re + zr, ai = a[k2].im - zi;
      const float br = a[k2].re - zr, bi = a[k2].im + zi;
      if (k1 + 16 * k2 <= 200) {
        pa[16 * k2] = ar * ar + ai * ai;  // 4 |Xa|^2: the factor 1/4 is folded into the mel weights (exact)
        pb[16 * k2] = br * br + bi * bi;
      }
    }
  }
  __syncthreads();

  if (done_tiles != nullptr && pend1 >= 0) {  // the FFT work area is free behind the staged rows
    int n;
    const float* src = tile_rows(pend1, n);
    for (int i = tid; i < n / 4; i += kLmThreads) cp_async16(s_back + 4 * i, src + 4 * i);
    cp_async_commit();
  }

  // ---- phase 5: sparse mel with one lane per frame, log10, staged rows, running max ---------------------
  // every lane of a warp walks the same (mel, bin) sequence: no divergence, weights are warp-uniform loads
  float vmax = -INFINITY;
  {
    const float* pw = s_power + lane * kPowStride;  // lane == frame within the tile
    float* so = s_out + lane * kOutStride;
    switch (warp) {
      case 0: mel_for_warp<NM, 0>(pw, so, vmax); break;
      case 1: mel_for_warp<NM, 1>(pw, so, vmax); break;
      case 2: mel_for_warp<NM, 2>(pw, so, vmax); break;
      case 3: mel_for_warp<NM, 3>(pw, so, vmax); break;
      case 4: mel_for_warp<NM, 4>(pw, so, vmax); break;
      case 5: mel_for_warp<NM, 5>(pw, so, vmax); break;
      case 6: mel_for_warp<NM, 6>(pw, so, vmax); break;
      default: mel_for_warp<NM, 7>(pw, so, vmax); break;
    }
    if (f0 + lane >= n_frames) vmax = -INFINITY;  // frames past the end of the signal do not count
  }
  __syncthreads();
  const int valid_frames = min(kFrames, n_frames - f0);
  float* o = out + ((long long)audio * n_frames + f0) * n_mels;
  {
    // thread -> (column m, first row r0), rows r0, r0 + kRows, ...: immediate offsets, conflict-free LDS, coalesced STG
    constexpr int kRows = kLmThreads / NM;  // 2 rows per pass (128 mels) or 3 (80 mels; 16 threads idle)
    const int m = tid % NM, r0 = tid / NM;
    if (r0 < kRows) {
      const float* sp = s_out + r0 * kOutStride + m;
      float* op = o + r0 * NM + m;
      if (valid_frames == kFrames) {
#pragma unroll
        for (int i = 0; i * kRows < kFrames; ++i)
          if (i * kRows + kRows <= kFrames || r0 + i * kRows < kFrames) op[i * kRows * NM] = sp[i * kRows * kOutStride];
      } else {
        for (int r = r0; r < valid_frames; r += kRows) o[r * NM + m] = s_out[r * kOutStride + m];
      }
    }
  }
  vmax = warp_max(vmax);
  if (lane == 0) s_red[warp] = vmax;
  __syncthreads();
  if (tid == 0) {
    float m = s_red[0];
#pragma unroll
    for (int i = 1; i < kLmThreads / 32; ++i) m = fmaxf(m, s_red[i]);
    if (m > -INFINITY) atomic_max_float(gmax + audio, m);
    if (done_tiles != nullptr) {
      __threadfence();  // rows (made visible to this thread by the barrier above) and the maximum before the count
      atomicAdd(done_tiles + audio, 1);
    }
  }
  if (done_tiles != nullptr) {
    if (pend1 >= 0) {  // every thread clamps the chunks it fetched itself: only its own copies need to have landed
      cp_async_wait_all();
      const float fl = s_floor;  // written in phase 1, barriers since
      int n;
      float4* o4 = reinterpret_cast<float4*>(tile_rows(pend1, n));
      for (int i = tid; i < n / 4; i += kLmThreads) o4[i] = clamp4(*reinterpret_cast<const float4*>(s_back + 4 * i), fl);
    }
    pend1 = pend0;
    pend0 = tile;
  }
  float* t = buf_cur;
  buf_cur = buf_nxt;
  buf_nxt = t;
  }
  if (done_tiles != nullptr) {
    __syncthreads();  // s_floor of the last pipelined clamp has been read by everyone
    if (pend1 >= 0) finalize_tile(pend1);
    if (pend0 >= 0) finalize_tile(pend0);
  }
}

__global__ void fill_f32_kernel(float* p, float v, int* counters, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    p[i] = v;
    if (counters != nullptr) counters[i] = 0;
  }
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
  static_assert(sizeof(lm::cpx) * kPairs * kPairStride >= sizeof(float) * kFrames * kOutStride, "staged rows fit the work area");
  return 2 * sizeof(float) * kRegion0Floats + sizeof(lm::cpx) * kPairs * kPairStride;
}

int init_logmel() {
  static bool done = false;
  if (done) return kOk;
  B200W_CUDA_OK(cudaFuncSetAttribute(logmel_kernel<80, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)logmel_smem_bytes()));
  B200W_CUDA_OK(cudaFuncSetAttribute(logmel_kernel<128, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)logmel_smem_bytes()));
  B200W_CUDA_OK(cudaFuncSetAttribute(logmel_kernel<80, int16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)logmel_smem_bytes()));
  B200W_CUDA_OK(cudaFuncSetAttribute(logmel_kernel<128, int16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)logmel_smem_bytes()));
  done = true;
  return kOk;
}

template <typename PcmT>
static int launch_logmel_t(const PcmT* pcm, int n_audio, long long audio_stride, long long n_valid, long long n_total,
                           int n_mels, const float* hann, const float* tw400, float* out, float* gmax, int* done_tiles,
                           cudaStream_t stream) {
  // the reflect pad works on the zero-extended signal (UPSTREAM pads first): only its total length must exceed the pad
  B200W_CHECK_ARG(n_audio > 0 && n_valid >= 0 && n_total >= n_valid && n_total > kNfft / 2,
                  "logmel: the (zero-extended) signal must be longer than the 200-sample reflect pad");
  B200W_CHECK_ARG(n_mels == 80 || n_mels == 128, "logmel: n_mels must be 80 or 128, got %d", n_mels);
  const long long n_frames_ll = n_total / kHop;
  B200W_CHECK_ARG(n_frames_ll > 0 && n_frames_ll < (1ll << 31) / 128, "logmel: frame count out of range");
  int n_frames = (int)n_frames_ll;
  const size_t smem = logmel_smem_bytes();
  B200W_TRY(init_logmel());
  ProfScope prof_("logmel", stream);
  fill_f32_kernel<<<ceil_div(n_audio, 256), 256, 0, stream>>>(gmax, -INFINITY, done_tiles, n_audio);
  B200W_LAUNCH_OK();
  LogmelTables tb{hann, reinterpret_cast<const float2*>(tw400)};
  int tiles_per_audio = ceil_div(n_frames, kFrames);
  const long long n_tiles_ll = (long long)tiles_per_audio * n_audio;
  B200W_CHECK_ARG(n_tiles_ll < (1ll << 31), "logmel: too many tiles");
  int n_tiles = (int)n_tiles_ll;
  const int grid = n_tiles < 2 * device_sm_count() ? n_tiles : 2 * device_sm_count();
  auto kernel = n_mels == 80 ? logmel_kernel<80, PcmT> : logmel_kernel<128, PcmT>;
  if (done_tiles == nullptr) {
    kernel<<<grid, kLmThreads, smem, stream>>>(pcm, audio_stride, n_valid, n_total, n_frames, tiles_per_audio, n_tiles, tb, out,
                                               gmax, done_tiles);
  } else {
    // fused normalisation: CTAs wait for one another's tiles, so all of them must be resident -- cooperative launch
    void* args[] = {(void*)&pcm, (void*)&audio_stride, (void*)&n_valid, (void*)&n_total, (void*)&n_frames,
                    (void*)&tiles_per_audio, (void*)&n_tiles, (void*)&tb, (void*)&out, (void*)&gmax, (void*)&done_tiles};
    B200W_CUDA_OK(cudaLaunchCooperativeKernel((const void*)kernel, dim3(grid), dim3(kLmThreads), args, smem, stream));
  }
  B200W_LAUNCH_OK();
  count_launch(2);
  return kOk;
}

int launch_logmel(const float* pcm, int n_audio, long long audio_stride, long long n_valid, long long n_total,
                  int n_mels, const float* hann, const float* tw400, float* out_unclamped, float* gmax,
                  cudaStream_t stream, int* done_tiles) {
  return launch_logmel_t(pcm, n_audio, audio_stride, n_valid, n_total, n_mels, hann, tw400, out_unclamped, gmax, done_tiles, stream);
}

int launch_logmel_pcm16(const int16_t* pcm, int n_audio, long long audio_stride, long long n_valid, long long n_total,
                        int n_mels, const float* hann, const float* tw400, float* out_unclamped, float* gmax,
                        cudaStream_t stream, int* done_tiles) {
  return launch_logmel_t(pcm, n_audio, audio_stride, n_valid, n_total, n_mels, hann, tw400, out_unclamped, gmax, done_tiles, stream);
}

int launch_logmel_finalize(float* x, const float* gmax, int n_audio, long long per_audio, cudaStream_t stream) {
  const long long total = per_audio * n_audio;
  const long long nthreads = ceil_div_ll(total, 4);
  ProfScope prof_("logmel_finalize", stream);
  logmel_finalize_kernel<<<(unsigned)ceil_div_ll(nthreads, 256), 256, 0, stream>>>(x, gmax, per_audio, total);
  B200W_LAUNCH_OK();
  count_launch();
  return kOk;
}

int launch_mel_windows(const float* mel, const float* gmax, const long long* row0, const int* size, const int* gidx,
                       int n_windows, int n_mels, __nv_bfloat16* dst, cudaStream_t stream) {
  B200W_CHECK_ARG(n_mels % 2 == 0 && n_windows > 0 && n_windows <= 65535, "mel_windows: bad sizes");
  dim3 grid(ceil_div(3002 * n_mels / 2, 256), n_windows);
  ProfScope prof_("mel_windows", stream);
  mel_window_kernel<<<grid, 256, 0, stream>>>(mel, gmax, row0, size, gidx, n_mels, dst);
  B200W_LAUNCH_OK();
  count_launch();
  return kOk;
}

}  // namespace b200w
