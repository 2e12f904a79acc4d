// K13: a whole single-token decoder step for a SMALL batch (<= 5 sequences) in ONE cooperative launch.
//
// The reference's `transcribe()` default -- what `./run` hits (/root/reference/run:3-6) -- decodes one 30 s window at a
// time (UPSTREAM transcribe.py seek loop -> decoding.py::DecodingTask._main_loop -> whisper.py::TextDecoder), i.e. up
// to 224 sequential steps at batch 1 (batch 5 for the `best_of` fallback).  Such a step streams the whole decoder
// (1.81 GB for large-v3) for a few kFLOP per weight: its floor is HBM time (0.3 ms) and what it actually costs is
// latency -- r01 ran it as 132 launches (attention kernels + tcgen05 chains of <= 128-row GEMMs padded from 1 row) at
// 2.59 ms.  This kernel is built for that regime instead:
//   * one CTA per SM, resident for the whole step; every phase boundary is a grid barrier (8 per layer);
//   * projections are matrix-VECTOR products on the FP32 pipes: output columns are dealt out to the CTAs, each CTA owns
//     its columns' full dot products (no split-K slabs, no cross-CTA reduction, deterministic);
//   * a producer warp streams the CTA's weight rows through a 7-stage shared-memory ring with cp.async.bulk
//     (TMA 1-D) in program order for the WHOLE step: it runs ahead of the consumers across phase boundaries and grid
//     barriers, so HBM keeps streaming while the consumers sit in a barrier or in an attention phase;
//   * LayerNorm is recomputed by every CTA from the fp32 residual stream (a few KB from L2) instead of being a phase of
//     its own; self- and cross-attention are phases of the same launch (keys of a (sequence, head) cut over the idle
//     CTAs, merged by the last one to arrive, as K8 does for small batches).
// Arithmetic (storage bf16, accumulation fp32, bf16 probabilities, gelu_fast) follows the large-batch path; only the
// summation order of the dot products differs.
#include "common.cuh"
#include "kernels.h"

namespace b200w {

constexpr int kSmWarps = 12;                      // consumer warps (+ 1 producer warp = 416 threads: 128 registers each)
constexpr int kSmConsumers = kSmWarps * 32;
constexpr int kSmThreads = kSmConsumers + 32;     // + the producer warp
constexpr int kSmStageBytes = 30720;              // 12 rows of K = 1280 or 3 rows of K = 5120
constexpr int kSmStages = 5;
constexpr int kSmHd = 64;
constexpr int kSmMaxKeys = 1536;
constexpr int kSmMaxK = 5120;
constexpr int kSmMaxPages = 32;
constexpr float kSmLog2e = 1.4426950408889634f;
constexpr int kSmPartFloats = 2 + kSmHd;
constexpr int kSmOwnMax = 16;  // columns of the residual stream one CTA owns (d / grid, rounded up)

constexpr int kSmRingBytes = kSmStages * kSmStageBytes;
constexpr int kSmActBytes = kSmallMaxBatch * kSmMaxK * 2;
constexpr int kSmSpBytes = kSmMaxKeys * 4;
constexpr int kSmPartBytes = kSmWarps * kSmHd * 4;
constexpr int kSmMiscBytes = 4096;
constexpr int kSmSmemBytes = kSmRingBytes + kSmActBytes + kSmSpBytes + kSmPartBytes + kSmMiscBytes + 128;
static_assert(kSmSmemBytes <= 232448, "shared memory budget of one CTA per SM");
// float offsets inside the misc area
constexpr int kMiscRed = 0;      // [B][12]
constexpr int kMiscPart = 96;    // [2][4 rows][4 parts][B]
constexpr int kMiscAttRed = 256; // [12]
constexpr int kMiscXown = 320;   // [B][kSmOwnMax]
constexpr int kMiscKvRow = 400;  // [B] long long (8-byte aligned)
constexpr int kMiscBias = 416;   // [48]: bias of this CTA's output columns in the current phase
constexpr int kMiscBt = 512;     // [B][kSmMaxPages] int
constexpr int kSmBiasMax = 48;
constexpr int kMiscProf = 680;   // [8] long long (development aid)
constexpr int kMiscBars = 992;   // mbarriers: last 128 bytes
static_assert(kMiscAttRed + 2 * kSmWarps <= kMiscXown && kSmallMaxBatch * kSmWarps <= kMiscPart && kMiscPart + 2 * 4 * 4 * kSmallMaxBatch <= kMiscAttRed &&
              kMiscAttRed + kSmWarps <= kMiscXown && kMiscXown + kSmallMaxBatch * kSmOwnMax <= kMiscKvRow &&
              kMiscKvRow + 2 * kSmallMaxBatch <= kMiscBias && kMiscBias + kSmBiasMax <= kMiscBt &&
              kMiscBt + kSmallMaxBatch * kSmMaxPages <= kMiscBars, "misc layout");

enum SmIn { kInLayerNorm = 0, kInVector = 1, kInMerge = 2 };
enum SmEpi { kEpiQkv = 0, kEpiResid = 1, kEpiBf16 = 2, kEpiGelu = 3, kEpiLogits = 4 };

__device__ __forceinline__ void sm_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kSmConsumers) : "memory"); }

__device__ __forceinline__ unsigned int sm_ld_acquire(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ unsigned long long sm_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// All consumer threads of all CTAs (the grid is cooperative: every CTA is resident).
// `timeline` (development aid, normally null): CTA 0 stamps the time at which it arrives at / leaves every barrier.
__device__ __forceinline__ void sm_grid_barrier(unsigned int* counter, unsigned int& epoch, unsigned long long* timeline) {
  sm_sync();
  if (threadIdx.x == 0) {
    if (timeline != nullptr && blockIdx.x == 0) timeline[2 * epoch + 1] = sm_globaltimer();
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
    const unsigned int target = (epoch + 1) * gridDim.x;
    unsigned int spins = 0;
    while (sm_ld_acquire(counter) < target) {
      if (++spins > (1u << 26)) __trap();  // a lost CTA must not hang the GPU
    }
    if (timeline != nullptr && blockIdx.x == 0) timeline[2 * epoch + 2] = sm_globaltimer();
  }
  ++epoch;
  sm_sync();
}

__device__ __forceinline__ void sm_bulk_load(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ float sm_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void sm_unpack8(const uint4& u, float (&f)[8]) {
  f[0] = __uint_as_float(u.x << 16); f[1] = __uint_as_float(u.x & 0xffff0000u);
  f[2] = __uint_as_float(u.y << 16); f[3] = __uint_as_float(u.y & 0xffff0000u);
  f[4] = __uint_as_float(u.z << 16); f[5] = __uint_as_float(u.z & 0xffff0000u);
  f[6] = __uint_as_float(u.w << 16); f[7] = __uint_as_float(u.w & 0xffff0000u);
}

struct SmRing {
  int stage;
  uint32_t phase;
  __device__ __forceinline__ void advance() {
    if (++stage == kSmStages) {
      stage = 0;
      phase ^= 1;
    }
  }
};

// rows of a weight matrix dealt to this CTA, and how many of them ride one ring stage: every consumer warp takes one
// row of a stage (K == d: 12 warps) or a quarter of one row (K == 4d: 3 rows x 4 warps)
__device__ __forceinline__ void sm_my_rows(int N, int& n0, int& n1) {
  n0 = (int)((long long)N * blockIdx.x / gridDim.x);
  n1 = (int)((long long)N * (blockIdx.x + 1) / gridDim.x);
}
__device__ __forceinline__ int sm_rows_per_stage(int K, int d) { return K == d ? kSmWarps : kSmWarps / 4; }

// ---- producer: the CTA's rows of W (N, K) bf16, chunk by chunk, into the ring ---------------------------------------
__device__ __forceinline__ void sm_produce(const void* W, int N, int K, int d, unsigned char* ring, uint64_t* full, uint64_t* empty,
                                           SmRing& rg) {
  int n0, n1;
  sm_my_rows(N, n0, n1);
  const int R = sm_rows_per_stage(K, d);
  const unsigned char* base = static_cast<const unsigned char*>(W);
  for (int r = n0; r < n1; r += R) {
    const int rows = min(R, n1 - r);
    const uint32_t bytes = (uint32_t)rows * (uint32_t)K * 2u;
    mbar_wait(&empty[rg.stage], rg.phase ^ 1);
    mbar_expect_tx(&full[rg.stage], bytes);
    sm_bulk_load(ring + rg.stage * kSmStageBytes, base + (size_t)r * K * 2, bytes, &full[rg.stage]);
    rg.advance();
  }
}

// ---- consumer: dot product of (a segment range of) one weight row with the B activation rows -------------------------
// S segments of 256 elements; every load is issued before the first product; 12 warps keep 3 such chains per scheduler.
template <int B, int S>
__device__ __forceinline__ void sm_dot1(const uint4* __restrict__ wp, const __nv_bfloat16* __restrict__ act, int act_ld, int lane,
                                        float (&acc)[B]) {
  uint4 w[S];
#pragma unroll
  for (int i = 0; i < S; ++i) w[i] = wp[i * 32 + lane];
#pragma unroll
  for (int b = 0; b < B; ++b) {
    uint4 x[S];
#pragma unroll
    for (int i = 0; i < S; ++i) x[i] = *reinterpret_cast<const uint4*>(act + (size_t)b * act_ld + i * 256 + lane * 8);
    float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
#pragma unroll
    for (int i = 0; i < S; ++i) {
      float xf[8], wf[8];
      sm_unpack8(x[i], xf);
      sm_unpack8(w[i], wf);
      s0 = fmaf(wf[0], xf[0], s0);
      s1 = fmaf(wf[1], xf[1], s1);
      s2 = fmaf(wf[2], xf[2], s2);
      s3 = fmaf(wf[3], xf[3], s3);
      s0 = fmaf(wf[4], xf[4], s0);
      s1 = fmaf(wf[5], xf[5], s1);
      s2 = fmaf(wf[6], xf[6], s2);
      s3 = fmaf(wf[7], xf[7], s3);
    }
    acc[b] = (s0 + s1) + (s2 + s3);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int b = 0; b < B; ++b) acc[b] += __shfl_xor_sync(0xffffffffu, acc[b], o);
  }
}

// B == 1: the lane's 8 * S activation values stay in registers (as f32) for all rows of a phase -- every row of every
// stage multiplies the same elements, so the 2 * S shared-memory loads and 8 * S unpack instructions per row shrink to
// S loads and 8 * S unpacks of the weights alone.
template <int S>
__device__ __forceinline__ void sm_load_x(const __nv_bfloat16* __restrict__ act, int lane, float (&xr)[S * 8]) {
#pragma unroll
  for (int i = 0; i < S; ++i) {
    float f[8];
    sm_unpack8(*reinterpret_cast<const uint4*>(act + i * 256 + lane * 8), f);
#pragma unroll
    for (int e = 0; e < 8; ++e) xr[i * 8 + e] = f[e];
  }
}
template <int S>
__device__ __forceinline__ float sm_dot1_reg(const uint4* __restrict__ wp, const float (&xr)[S * 8], int lane) {
  uint4 w[S];
#pragma unroll
  for (int i = 0; i < S; ++i) w[i] = wp[i * 32 + lane];
  float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
#pragma unroll
  for (int i = 0; i < S; ++i) {
    float wf[8];
    sm_unpack8(w[i], wf);
    s0 = fmaf(wf[0], xr[i * 8 + 0], s0);
    s1 = fmaf(wf[1], xr[i * 8 + 1], s1);
    s2 = fmaf(wf[2], xr[i * 8 + 2], s2);
    s3 = fmaf(wf[3], xr[i * 8 + 3], s3);
    s0 = fmaf(wf[4], xr[i * 8 + 4], s0);
    s1 = fmaf(wf[5], xr[i * 8 + 5], s1);
    s2 = fmaf(wf[6], xr[i * 8 + 6], s2);
    s3 = fmaf(wf[7], xr[i * 8 + 7], s3);
  }
  float acc = (s0 + s1) + (s2 + s3);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  return acc;
}

struct SmCtx {
  const SmallArgs* a;
  unsigned char* ring;
  __nv_bfloat16* act;
  float* s_p;
  float* s_part;   // [8][64]
  float* s_misc;   // reductions, partial sums, per-step tables, mbarriers
  float* s_xown;   // [B][kSmOwnMax]: this CTA's columns of the fp32 residual stream, kept on chip for the whole step
  long long* s_kvrow;  // [B]: row of this step's position in the paged self K/V cache
  int* s_bt;       // [B][kSmMaxPages]: the sequences' page tables
  uint64_t *full, *empty;
  SmRing rg;
  unsigned int epoch;
  int tid, warp, lane;
  int own_n0, ca_splits;
  float4 ln_g, ln_b;  // gamma / beta of the NEXT LayerNorm phase (four features of this thread), fetched a phase early
  long long* s_prof;  // development aid (timeline != null): SM cycles of thread 0 in LN / wait / dot / epilogue / SA / CA / input
  bool prof;
};

__device__ __forceinline__ long long sm_tick(const SmCtx& c) { return (c.prof && c.tid == 0) ? clock64() : 0; }
__device__ __forceinline__ void sm_tock(const SmCtx& c, int slot, long long& t0) {
  if (c.prof && c.tid == 0) {
    const long long t1 = clock64();
    c.s_prof[slot] += t1 - t0;
    t0 = t1;
  }
}

// acts <- LayerNorm(x[b]) * gamma + beta as bf16, for every sequence (each CTA recomputes it: x is a few KB in L2).
// Thread t < d / 4 owns four consecutive features.  gamma / beta of THIS phase were fetched one phase ago (they come
// from HBM: nothing else keeps them in L2 between steps); the next LayerNorm's are requested here.
template <int B>
__device__ __forceinline__ void sm_input_layernorm(SmCtx& c, int d, const float* next_gamma, const float* next_beta) {
  const SmallArgs& a = *c.a;
  const bool on = c.tid < d / 4;
  float4 v[B];
  float* red = c.s_misc + kMiscRed;
#pragma unroll
  for (int b = 0; b < B; ++b) v[b] = on ? __ldcg(reinterpret_cast<const float4*>(a.x + (size_t)b * d) + c.tid) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 g = c.ln_g, be = c.ln_b;
  if (on && next_gamma != nullptr) {
    c.ln_g = __ldg(reinterpret_cast<const float4*>(next_gamma) + c.tid);
    c.ln_b = __ldg(reinterpret_cast<const float4*>(next_beta) + c.tid);
  }
#pragma unroll
  for (int b = 0; b < B; ++b) {
    float s = warp_sum((v[b].x + v[b].y) + (v[b].z + v[b].w));
    if (c.lane == 0) red[b * kSmWarps + c.warp] = s;
  }
  sm_sync();
  float mean[B];
#pragma unroll
  for (int b = 0; b < B; ++b) {
    float s = 0.0f;
#pragma unroll
    for (int w = 0; w < kSmWarps; ++w) s += red[b * kSmWarps + w];
    mean[b] = s / (float)d;
  }
  sm_sync();
#pragma unroll
  for (int b = 0; b < B; ++b) {
    if (on) {
      v[b].x -= mean[b]; v[b].y -= mean[b]; v[b].z -= mean[b]; v[b].w -= mean[b];
    }
    float s = warp_sum((v[b].x * v[b].x + v[b].y * v[b].y) + (v[b].z * v[b].z + v[b].w * v[b].w));
    if (c.lane == 0) red[b * kSmWarps + c.warp] = s;
  }
  sm_sync();
#pragma unroll
  for (int b = 0; b < B; ++b) {
    float s = 0.0f;
#pragma unroll
    for (int w = 0; w < kSmWarps; ++w) s += red[b * kSmWarps + w];
    const float rstd = rsqrtf(s / (float)d + 1e-5f);
    if (on)
      *reinterpret_cast<uint2*>(c.act + (size_t)b * d + c.tid * 4) =
          make_uint2(pack_bf16x2(v[b].x * rstd * g.x + be.x, v[b].y * rstd * g.y + be.y),
                     pack_bf16x2(v[b].z * rstd * g.z + be.z, v[b].w * rstd * g.w + be.w));
  }
  sm_sync();
}

// acts <- src (B, K) bf16 written by other CTAs in the previous phase
template <int B>
__device__ __forceinline__ void sm_input_vector(SmCtx& c, const __nv_bfloat16* src, int K) {
  const int n16 = B * K / 8;  // 16-byte pieces
  constexpr int kMax = (B * kSmMaxK / 8 + kSmConsumers - 1) / kSmConsumers;  // per thread: 2 (B = 1) .. 9 (B = 5)
  const uint4* s = reinterpret_cast<const uint4*>(src);
  uint4* dst = reinterpret_cast<uint4*>(c.act);
  uint4 u[kMax];
#pragma unroll
  for (int k = 0; k < kMax; ++k) {
    const int i = c.tid + k * kSmConsumers;
    if (i < n16) u[k] = __ldcg(s + i);
  }
#pragma unroll
  for (int k = 0; k < kMax; ++k) {
    const int i = c.tid + k * kSmConsumers;
    if (i < n16) dst[i] = u[k];
  }
  sm_sync();
}

// acts <- cross-attention output merged from the key-split partials (max, sum, unnormalised output) that the CTAs of the
// previous phase left in global memory: every CTA merges for itself, so the attention phase needs no arrival counter,
// no "last CTA" and no second round trip -- the partials are a few tens of KB in L2.
template <int B>
__device__ __forceinline__ void sm_input_merge(SmCtx& c) {
  const SmallArgs& a = *c.a;
  const int H = a.n_head, d = a.d, splits = c.ca_splits;
  for (int unit = c.warp; unit < B * H; unit += kSmWarps) {  // one (sequence, head) per warp, two dims per lane
    const float* all = a.ca_part + (size_t)unit * splits * kSmPartFloats;
    float mk[8], lk[8], o0[8], o1[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const bool on = k < splits;
      mk[k] = on ? __ldcg(all + k * kSmPartFloats) : -INFINITY;
      lk[k] = on ? __ldcg(all + k * kSmPartFloats + 1) : 0.0f;
      o0[k] = on ? __ldcg(all + k * kSmPartFloats + 2 + c.lane) : 0.0f;
      o1[k] = on ? __ldcg(all + k * kSmPartFloats + 2 + 32 + c.lane) : 0.0f;
    }
    float M = mk[0];
#pragma unroll
    for (int k = 1; k < 8; ++k) M = fmaxf(M, mk[k]);
    float Lsum = 0.0f, a0 = 0.0f, a1 = 0.0f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float w = (k < splits) ? sm_exp2(mk[k] - M) : 0.0f;
      Lsum = fmaf(lk[k], w, Lsum);
      a0 = fmaf(o0[k], w, a0);
      a1 = fmaf(o1[k], w, a1);
    }
    const int b = unit / H, h = unit - b * H;
    c.act[(size_t)b * d + h * kSmHd + c.lane] = __float2bfloat16(a0 / Lsum);
    c.act[(size_t)b * d + h * kSmHd + 32 + c.lane] = __float2bfloat16(a1 / Lsum);
  }
  sm_sync();
}

// value v (bias already added) of output column n for sequence b
__device__ __forceinline__ void sm_epilogue(const SmCtx& c, int epi, int layer, int n, int b, float v, void* out, int N) {
  const SmallArgs& a = *c.a;
  const int d = a.d;
  if (epi == kEpiLogits) {
    a.logits[(size_t)b * a.logits_ld + n] = v;
  } else if (epi == kEpiResid) {
    // this CTA owns column n of the residual stream in every layer: the running value never leaves the chip
    float* own = c.s_xown + b * kSmOwnMax + (n - c.own_n0);
    const float xn = *own + v;
    *own = xn;
    a.x[(size_t)b * d + n] = xn;  // for the other CTAs' LayerNorms
  } else if (epi == kEpiBf16) {
    static_cast<__nv_bfloat16*>(out)[(size_t)b * N + n] = __float2bfloat16(v);
  } else if (epi == kEpiGelu) {
    static_cast<__nv_bfloat16*>(out)[(size_t)b * N + n] = __float2bfloat16(gelu_fast(v));
  } else {  // fused q | k | v: q to the query buffer, k / v appended to the paged cache at this step's position
    if (n < d) {
      a.q[(size_t)b * d + n] = __float2bfloat16(v);
    } else {
      __nv_bfloat16* pages = (n < 2 * d ? a.k_pages : a.v_pages) + (size_t)layer * a.layer_page_stride;
      pages[c.s_kvrow[b] * d + (n < 2 * d ? n - d : n - 2 * d)] = __float2bfloat16(v);
    }
  }
}

template <int B, int S>
__device__ __forceinline__ void sm_gemv_chunks(SmCtx& c, int N, int K, int epi, int layer, bool has_bias, void* out) {
  static_assert(B <= 32, "one epilogue lane per sequence");
  int n0, n1;
  sm_my_rows(N, n0, n1);
  const bool wide = K != c.a->d;       // K == 4d: four warps share a row, each takes a quarter of K
  const int wpr = wide ? 4 : 1;
  const int R = kSmWarps / wpr;        // rows per stage
  const int row_in = c.warp / wpr, kpart = c.warp - row_in * wpr;
  const int klen = K / wpr;            // = S * 256
  float* part = c.s_misc + kMiscPart;  // [2][4 rows][4 parts][B]: double-buffered by chunk parity
  const float* s_bias = c.s_misc + kMiscBias;
  int parity = 0;
  float xr[B == 1 ? S * 8 : 1];
  if constexpr (B == 1) sm_load_x<S>(c.act + kpart * klen, c.lane, xr);
  for (int r = n0; r < n1; r += R) {
    const int rows = min(R, n1 - r);
    const bool mine_row = row_in < rows;
    const int my_row = r + row_in;
    long long t0 = sm_tick(c);
    mbar_wait(&c.full[c.rg.stage], c.rg.phase);
    sm_tock(c, 1, t0);
    float acc[B];
    if (mine_row) {
      const unsigned char* wrow = c.ring + c.rg.stage * kSmStageBytes + ((size_t)row_in * K + (size_t)kpart * klen) * 2;
      if constexpr (B == 1) acc[0] = sm_dot1_reg<S>(reinterpret_cast<const uint4*>(wrow), xr, c.lane);
      else sm_dot1<B, S>(reinterpret_cast<const uint4*>(wrow), c.act + kpart * klen, K, c.lane, acc);
    }
    __syncwarp();
    if (c.lane == 0) mbar_arrive(&c.empty[c.rg.stage]);  // the stage can be refilled while the results are written
    c.rg.advance();
    float mine = 0.0f;
#pragma unroll
    for (int b = 0; b < B; ++b)
      if (c.lane == b) mine = acc[b];
    sm_tock(c, 2, t0);
    const float bv = (has_bias && mine_row) ? s_bias[my_row - n0] : 0.0f;
    if (!wide) {
      if (mine_row && c.lane < B) sm_epilogue(c, epi, layer, my_row, c.lane, mine + bv, out, N);
    } else {
      float* pp = part + parity * (4 * 4 * kSmallMaxBatch);
      if (mine_row && c.lane < B) pp[(row_in * 4 + kpart) * kSmallMaxBatch + c.lane] = mine;
      sm_sync();
      if (kpart == 0 && mine_row && c.lane < B) {
        const float* q = pp + row_in * 4 * kSmallMaxBatch + c.lane;
        const float v = (q[0] + q[kSmallMaxBatch]) + (q[2 * kSmallMaxBatch] + q[3 * kSmallMaxBatch]);
        sm_epilogue(c, epi, layer, my_row, c.lane, v + bv, out, N);
      }
      parity ^= 1;
    }
    sm_tock(c, 3, t0);
  }
}

// one projection phase: input vector(s) -> shared memory, then this CTA's output columns.
// (next_g, next_b): parameters of the LayerNorm phase after this one, requested now if this phase is a LayerNorm phase.
template <int B, int S>
__device__ __forceinline__ void sm_gemv(SmCtx& c, int in_kind, const float* next_g, const float* next_b, const __nv_bfloat16* vec, int N,
                                        int K, int epi, int layer, const float* bias, void* out) {
  long long t0 = sm_tick(c);
  if (bias != nullptr) {  // this CTA's slice of the bias (from HBM) travels under the input stage
    int n0, n1;
    sm_my_rows(N, n0, n1);
    if (c.tid < n1 - n0) c.s_misc[kMiscBias + c.tid] = __ldg(bias + n0 + c.tid);
  }
  if (in_kind == kInLayerNorm) sm_input_layernorm<B>(c, K, next_g, next_b);
  else if (in_kind == kInMerge) sm_input_merge<B>(c);
  else sm_input_vector<B>(c, vec, K);
  sm_tock(c, in_kind == kInLayerNorm ? 0 : 6, t0);
  sm_gemv_chunks<B, S>(c, N, K, epi, layer, bias != nullptr, out);
}

// ---- attention over `T` key rows of 64 dims for one query head ------------------------------------------------------
// 8 lanes per key row (16 B each), 4 rows per warp per load instruction, 5 + 5 loads in flight per thread (240 keys per
// CTA sweep); the first sweep of V is issued together with K (it does not depend on the probabilities), so a unit of up to
// 256 keys costs one round trip to L2 / HBM.  Returns the (max, sum) of the key range and the unnormalised output of
// dim `tid` (< 64).
template <typename KRow, typename VRow>
__device__ __forceinline__ void sm_attend(SmCtx& c, const float (&qv)[8], int T, KRow krow, VRow vrow, float& mx_out, float& sum_out,
                                          float& o_out) {
  const int sub = c.lane & 7, kg = c.lane >> 3;
  constexpr int kU = 5, kStep = kSmWarps * 4, kSweep = kU * kStep;  // 240 keys per sweep
  float mx = -INFINITY;
  uint4 v0[kU];
  const int first = c.warp * 4;
  {
    uint4 u[kU];
#pragma unroll
    for (int i = 0; i < kU; ++i) u[i] = __ldcg(reinterpret_cast<const uint4*>(krow(min(first + kg + i * kStep, T - 1))) + sub);
#pragma unroll
    for (int i = 0; i < kU; ++i) v0[i] = __ldcg(reinterpret_cast<const uint4*>(vrow(min(first + kg + i * kStep, T - 1))) + sub);
    auto score = [&](const uint4 (&uu)[kU], int j0) {
#pragma unroll
      for (int i = 0; i < kU; ++i) {
        const int j = j0 + kg + i * kStep;
        float f[8];
        sm_unpack8(uu[i], f);
        float s = f[0] * qv[0];
#pragma unroll
        for (int e = 1; e < 8; ++e) s = fmaf(f[e], qv[e], s);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        if (j < T) {
          if (sub == 0) c.s_p[j] = s;
          mx = fmaxf(mx, s);
        }
      }
    };
    score(u, first);
    for (int j0 = first + kSweep; j0 < T + first; j0 += kSweep) {  // warp-uniform trip count: the shuffles need every lane
      if (j0 - first >= T) break;
#pragma unroll
      for (int i = 0; i < kU; ++i) u[i] = __ldcg(reinterpret_cast<const uint4*>(krow(min(j0 + kg + i * kStep, T - 1))) + sub);
      score(u, j0);
    }
  }
  float* red = c.s_misc + kMiscAttRed;
  mx = warp_max(mx);
  if (c.lane == 0) red[c.warp] = mx;
  sm_sync();
  mx = red[0];
#pragma unroll
  for (int i = 1; i < kSmWarps; ++i) mx = fmaxf(mx, red[i]);
  sm_sync();
  float sum = 0.0f;
  for (int j = c.tid; j < T; j += kSmConsumers) {
    const float p = sm_exp2(c.s_p[j] - mx);
    sum += p;
    c.s_p[j] = __bfloat162float(__float2bfloat16(p));  // bf16 probabilities, as in the tensor-core path
  }
  sum = warp_sum(sum);
  if (c.lane == 0) red[c.warp] = sum;
  sm_sync();
  sum = 0.0f;
#pragma unroll
  for (int i = 0; i < kSmWarps; ++i) sum += red[i];
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.0f;
  auto accumulate = [&](const uint4 (&uu)[kU], int j0) {
#pragma unroll
    for (int i = 0; i < kU; ++i) {
      const int j = j0 + kg + i * kStep;
      const float p = (j < T) ? c.s_p[j] : 0.0f;
      float f[8];
      sm_unpack8(uu[i], f);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = fmaf(p, f[e], acc[e]);
    }
  };
  accumulate(v0, first);
  for (int j0 = first + kSweep; j0 - first < T; j0 += kSweep) {
    uint4 u[kU];
#pragma unroll
    for (int i = 0; i < kU; ++i) u[i] = __ldcg(reinterpret_cast<const uint4*>(vrow(min(j0 + kg + i * kStep, T - 1))) + sub);
    accumulate(u, j0);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 8);
    acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 16);
  }
  if (kg == 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) c.s_part[c.warp * kSmHd + sub * 8 + i] = acc[i];
  }
  sm_sync();
  float o = 0.0f;
  if (c.tid < kSmHd) {
#pragma unroll
    for (int w = 0; w < kSmWarps; ++w) o += c.s_part[w * kSmHd + c.tid];
  }
  mx_out = mx;
  sum_out = sum;
  o_out = o;
  sm_sync();  // s_p / s_part / red are reused by the next unit
}

__device__ __forceinline__ void sm_load_q(const SmCtx& c, const __nv_bfloat16* q, float (&qv)[8]) {
  float f[8];
  sm_unpack8(__ldcg(reinterpret_cast<const uint4*>(q) + (c.lane & 7)), f);
  const float s = 0.125f * kSmLog2e;
#pragma unroll
  for (int i = 0; i < 8; ++i) qv[i] = f[i] * s;
}

__device__ __forceinline__ void sm_self_attention(SmCtx& c, int layer) {
  const SmallArgs& a = *c.a;
  const int d = a.d, H = a.n_head;
  const int pshift = __ffs(a.page_size) - 1;
  for (int u = blockIdx.x; u < a.B * H; u += gridDim.x) {
    const int b = u / H, h = u - b * H;
    if (a.finished != nullptr && a.finished[b]) continue;  // (CTA-uniform)
    const int n_keys = __ldg(a.pos + b) + 1;               // the row of this step was appended by the QKV phase
    const int* bt = c.s_bt + b * kSmMaxPages;              // page ids: in shared memory since the start of the step
    const __nv_bfloat16* kp = a.k_pages + (size_t)layer * a.layer_page_stride + h * kSmHd;
    const __nv_bfloat16* vp = a.v_pages + (size_t)layer * a.layer_page_stride + h * kSmHd;
    float qv[8];
    sm_load_q(c, a.q + (size_t)b * d + h * kSmHd, qv);
    auto krow = [&](int j) { return kp + ((long long)bt[j >> pshift] * a.page_size + (j & (a.page_size - 1))) * d; };
    auto vrow = [&](int j) { return vp + ((long long)bt[j >> pshift] * a.page_size + (j & (a.page_size - 1))) * d; };
    float mx, sum, o;
    sm_attend(c, qv, n_keys, krow, vrow, mx, sum, o);
    if (c.tid < kSmHd) a.att[(size_t)b * d + h * kSmHd + c.tid] = __float2bfloat16(o / sum);
  }
}

// Keys of a (sequence, head) are cut into `ca_splits` ranges handled by different CTAs (B x H units alone would leave
// most SMs idle); a range leaves (max, sum, unnormalised output) for the merge in the next phase's input stage.
__device__ __forceinline__ void sm_cross_attention(SmCtx& c, int layer) {
  const SmallArgs& a = *c.a;
  const int d = a.d, H = a.n_head, T_all = a.T;
  const int units = a.B * H, splits = c.ca_splits;
  const int per = ((T_all + splits - 1) / splits + 31) & ~31;
  const long long ld = 2ll * d;
  for (int w = blockIdx.x; w < units * splits; w += gridDim.x) {
    const int unit = w / splits, chunk = w - unit * splits;
    const int b = unit / H, h = unit - b * H;
    if (a.finished != nullptr && a.finished[b]) continue;
    const int k0 = min(chunk * per, T_all), T = min(per, T_all - k0);
    const __nv_bfloat16* kb = a.cross_kv + (size_t)layer * a.cross_layer_stride + (long long)__ldg(a.cross_slot + b) * a.cross_seq_stride +
                              (long long)k0 * ld + h * kSmHd;
    float mx = -INFINITY, sum = 0.0f, o = 0.0f;
    if (T > 0) {
      float qv[8];
      sm_load_q(c, a.qc + (size_t)b * d + h * kSmHd, qv);
      auto krow = [&](int j) { return kb + j * ld; };
      auto vrow = [&](int j) { return kb + d + j * ld; };
      sm_attend(c, qv, T, krow, vrow, mx, sum, o);
    }
    if (splits == 1) {
      if (c.tid < kSmHd) a.att[(size_t)b * d + h * kSmHd + c.tid] = __float2bfloat16(o / sum);
    } else {
      float* mine = a.ca_part + ((size_t)unit * splits + chunk) * kSmPartFloats;
      if (c.tid < kSmHd) mine[2 + c.tid] = o;
      if (c.tid == 0) {
        mine[0] = mx;  // -inf for an empty range: weight 0 in the merge
        mine[1] = sum;
      }
    }
  }
}

template <int B, int S>
__device__ __forceinline__ void sm_consumer(SmCtx& c) {
  const SmallArgs& a = *c.a;
  const int d = a.d, L = a.n_layer;
  {  // per-step constants of this CTA: its columns of the residual stream, page tables, cache rows of this step's position
    int n1;
    sm_my_rows(d, c.own_n0, n1);
    const int cnt = n1 - c.own_n0;
    for (int i = c.tid; i < B * cnt; i += kSmConsumers) {
      const int b = i / cnt, j = i - b * cnt;
      c.s_xown[b * kSmOwnMax + j] = a.x[(size_t)b * d + c.own_n0 + j];
    }
    for (int i = c.tid; i < B * a.max_pages; i += kSmConsumers) {
      const int b = i / a.max_pages, j = i - b * a.max_pages;
      c.s_bt[b * kSmMaxPages + j] = a.block_table[b * a.max_pages + j];
    }
    if (c.tid < B) {
      const int p = a.pos[c.tid];
      const int pshift = __ffs(a.page_size) - 1;
      c.s_kvrow[c.tid] = (long long)a.block_table[c.tid * a.max_pages + (p >> pshift)] * a.page_size + (p & (a.page_size - 1));
    }
    int splits = (int)gridDim.x / (B * a.n_head);
    c.ca_splits = splits < 1 ? 1 : (splits > 8 ? 8 : splits);
    if (c.tid < d / 4) {  // parameters of the first LayerNorm
      c.ln_g = __ldg(reinterpret_cast<const float4*>(a.layers[0].attn_ln_g) + c.tid);
      c.ln_b = __ldg(reinterpret_cast<const float4*>(a.layers[0].attn_ln_b) + c.tid);
    }
    sm_sync();
  }
  if (a.timeline != nullptr && blockIdx.x == 0 && c.tid == 0) a.timeline[0] = sm_globaltimer();
  const int cross_in = c.ca_splits > 1 ? kInMerge : kInVector;
  for (int l = 0; l < L; ++l) {
    const b200w_dec_layer& W = a.layers[l];
    sm_gemv<B, S>(c, kInLayerNorm, W.cross_ln_g, W.cross_ln_b, nullptr, 3 * d, d, kEpiQkv, l, W.b_qkv, nullptr);
    sm_grid_barrier(a.counter, c.epoch, a.timeline);
    {
      long long t0 = sm_tick(c);
      sm_self_attention(c, l);
      sm_tock(c, 4, t0);
    }
    sm_grid_barrier(a.counter, c.epoch, a.timeline);
    sm_gemv<B, S>(c, kInVector, nullptr, nullptr, a.att, d, d, kEpiResid, l, W.b_out, nullptr);
    sm_grid_barrier(a.counter, c.epoch, a.timeline);
    sm_gemv<B, S>(c, kInLayerNorm, W.mlp_ln_g, W.mlp_ln_b, nullptr, d, d, kEpiBf16, l, W.b_cq, a.qc);
    sm_grid_barrier(a.counter, c.epoch, a.timeline);
    {
      long long t0 = sm_tick(c);
      sm_cross_attention(c, l);
      sm_tock(c, 5, t0);
    }
    sm_grid_barrier(a.counter, c.epoch, a.timeline);
    sm_gemv<B, S>(c, cross_in, nullptr, nullptr, a.att, d, d, kEpiResid, l, W.b_cout, nullptr);
    sm_grid_barrier(a.counter, c.epoch, a.timeline);
    sm_gemv<B, S>(c, kInLayerNorm, l + 1 < L ? a.layers[l + 1].attn_ln_g : a.dec_ln_g, l + 1 < L ? a.layers[l + 1].attn_ln_b : a.dec_ln_b,
               nullptr, 4 * d, d, kEpiGelu, l, W.b_mlp1, a.mlp);
    sm_grid_barrier(a.counter, c.epoch, a.timeline);
    sm_gemv<B, S>(c, kInVector, nullptr, nullptr, a.mlp, d, 4 * d, kEpiResid, l, W.b_mlp2, nullptr);
    sm_grid_barrier(a.counter, c.epoch, a.timeline);
  }
  // final LayerNorm + tied logits of the (single) new token of every sequence
  sm_gemv<B, S>(c, kInLayerNorm, nullptr, nullptr, nullptr, a.n_vocab, d, kEpiLogits, 0, nullptr, nullptr);
  if (a.timeline != nullptr && blockIdx.x == 0 && c.tid == 0) {
    a.timeline[2 * c.epoch + 1] = sm_globaltimer();
    for (int i = 0; i < 8; ++i) a.timeline[2 * c.epoch + 2 + i] = (unsigned long long)c.s_prof[i];
  }
}

template <int B, int S>
__global__ void __launch_bounds__(kSmThreads, 1) decode_small_kernel(const __grid_constant__ SmallArgs a) {
  extern __shared__ unsigned char sm_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(sm_raw) + 127) & ~static_cast<uintptr_t>(127));
  SmCtx c;
  c.a = &a;
  c.ring = smem;
  c.act = reinterpret_cast<__nv_bfloat16*>(smem + kSmRingBytes);
  c.s_p = reinterpret_cast<float*>(smem + kSmRingBytes + kSmActBytes);
  c.s_part = reinterpret_cast<float*>(smem + kSmRingBytes + kSmActBytes + kSmSpBytes);
  c.s_misc = reinterpret_cast<float*>(smem + kSmRingBytes + kSmActBytes + kSmSpBytes + kSmPartBytes);
  c.s_xown = c.s_misc + kMiscXown;
  c.s_kvrow = reinterpret_cast<long long*>(c.s_misc + kMiscKvRow);
  c.s_bt = reinterpret_cast<int*>(c.s_misc + kMiscBt);
  uint64_t* bars = reinterpret_cast<uint64_t*>(c.s_misc + kMiscBars);
  c.full = bars;
  c.empty = bars + kSmStages;
  c.rg = SmRing{0, 0};
  c.epoch = 0;
  c.tid = threadIdx.x;
  c.warp = threadIdx.x >> 5;
  c.lane = threadIdx.x & 31;
  c.own_n0 = 0;
  c.ca_splits = 1;
  c.ln_g = c.ln_b = make_float4(0.f, 0.f, 0.f, 0.f);
  c.s_prof = reinterpret_cast<long long*>(c.s_misc + kMiscProf);
  c.prof = a.timeline != nullptr && blockIdx.x == 0;
  if (threadIdx.x < 8) c.s_prof[threadIdx.x] = 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kSmStages; ++i) {
      mbar_init(&c.full[i], 1);
      mbar_init(&c.empty[i], kSmWarps);
    }
    fence_barrier_init();
  }
  __syncthreads();

  const int d = a.d, L = a.n_layer;
  if (c.warp == kSmWarps) {
    // ---- producer warp: every weight row this CTA will need during the step, in program order ----
    if (c.lane == 0) {
      for (int l = 0; l < L; ++l) {
        const b200w_dec_layer& W = a.layers[l];
        sm_produce(W.w_qkv, 3 * d, d, d, c.ring, c.full, c.empty, c.rg);
        sm_produce(W.w_out, d, d, d, c.ring, c.full, c.empty, c.rg);
        sm_produce(W.w_cq, d, d, d, c.ring, c.full, c.empty, c.rg);
        sm_produce(W.w_cout, d, d, d, c.ring, c.full, c.empty, c.rg);
        sm_produce(W.w_mlp1, 4 * d, d, d, c.ring, c.full, c.empty, c.rg);
        sm_produce(W.w_mlp2, d, 4 * d, d, c.ring, c.full, c.empty, c.rg);
      }
      sm_produce(a.tok_emb, a.n_vocab, d, d, c.ring, c.full, c.empty, c.rg);
    }
  } else {
    sm_consumer<B, S>(c);
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------- host
static bool g_small_ready = false;
static unsigned long long* g_small_timeline = nullptr;  // development aid (tools/profile_small.py)
void set_decode_small_timeline(unsigned long long* dev) { g_small_timeline = dev; }

typedef void (*SmallKernel)(const SmallArgs);
template <int B>
static SmallKernel small_kernel_for_s(int s) {
  switch (s) {
    case 3: return decode_small_kernel<B, 3>;
    case 4: return decode_small_kernel<B, 4>;
    case 5: return decode_small_kernel<B, 5>;
    default: return nullptr;
  }
}
static SmallKernel small_kernel(int B, int s) {
  switch (B) {
    case 1: return small_kernel_for_s<1>(s);
    case 2: return small_kernel_for_s<2>(s);
    case 3: return small_kernel_for_s<3>(s);
    case 4: return small_kernel_for_s<4>(s);
    case 5: return small_kernel_for_s<5>(s);
    default: return nullptr;
  }
}

int init_decode_small() {
  if (g_small_ready) return kOk;
  for (int B = 1; B <= kSmallMaxBatch; ++B)
    for (int s2 = 3; s2 <= 5; ++s2)
      B200W_CUDA_OK(cudaFuncSetAttribute(small_kernel(B, s2), cudaFuncAttributeMaxDynamicSharedMemorySize, kSmSmemBytes));
  g_small_ready = true;
  return kOk;
}

bool decode_small_applicable(const b200w_dims& dm, int n_seq, int n_q) {
  const int d = dm.n_text_state;
  return n_q == 1 && n_seq >= 1 && n_seq <= kSmallMaxBatch && d % 256 == 0 && d / 256 <= 5 && d / 4 <= kSmConsumers && 4 * d <= kSmMaxK && (d + 147) / 148 <= kSmOwnMax &&
         dm.n_audio_ctx <= kSmMaxKeys && dm.n_text_ctx <= kSmMaxKeys && n_seq * dm.n_text_head <= device_sm_count();
}

int launch_decode_small(const SmallArgs& a, cudaStream_t stream) {
  B200W_CHECK_ARG(a.layers && a.tok_emb && a.x && a.q && a.att && a.qc && a.mlp && a.logits && a.ca_part && a.ca_cnt && a.counter,
                  "decode_small: null pointer");
  B200W_CHECK_ARG(a.page_size > 0 && (a.page_size & (a.page_size - 1)) == 0, "decode_small: page_size must be a power of two");
  B200W_TRY(init_decode_small());
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(device_sm_count());
  cfg.blockDim = dim3(kSmThreads);
  cfg.dynamicSmemBytes = kSmSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;  // all CTAs co-resident: the grid barriers cannot deadlock
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  ProfScope prof_("decode_small", stream);
  SmallArgs args = a;
  args.timeline = g_small_timeline;
  SmallKernel kernel = small_kernel(a.B, a.d / 256);
  B200W_CHECK_ARG(kernel != nullptr, "decode_small: unsupported batch %d / width %d", a.B, a.d);
  B200W_CUDA_OK(cudaLaunchKernelEx(&cfg, kernel, args));
  count_launch();
  return kOk;
}

}  // namespace b200w
