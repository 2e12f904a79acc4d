// K13: a whole single-token decoder step for a SMALL batch (<= 6 sequences) in ONE cooperative launch.
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

constexpr int kSmWarps = 8;                       // consumer warps
constexpr int kSmConsumers = kSmWarps * 32;
constexpr int kSmThreads = kSmConsumers + 32;     // + the producer warp
constexpr int kSmStageBytes = 20480;              // 8 rows of K = 1280 or 2 rows of K = 5120
constexpr int kSmStages = 7;
constexpr int kSmHd = 64;
constexpr int kSmMaxKeys = 1536;
constexpr int kSmMaxK = 5120;
constexpr float kSmLog2e = 1.4426950408889634f;
constexpr int kSmPartFloats = 2 + kSmHd;

constexpr int kSmRingBytes = kSmStages * kSmStageBytes;
constexpr int kSmActBytes = kSmallMaxBatch * kSmMaxK * 2;
constexpr int kSmSpBytes = kSmMaxKeys * 4;
constexpr int kSmPartBytes = kSmWarps * kSmHd * 4;
constexpr int kSmMiscBytes = 2048;
constexpr int kSmSmemBytes = kSmRingBytes + kSmActBytes + kSmSpBytes + kSmPartBytes + kSmMiscBytes + 128;

enum SmIn { kInLayerNorm = 0, kInVector = 1 };
enum SmEpi { kEpiQkv = 0, kEpiResid = 1, kEpiBf16 = 2, kEpiGelu = 3, kEpiLogits = 4 };

__device__ __forceinline__ void sm_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kSmConsumers) : "memory"); }

__device__ __forceinline__ unsigned int sm_ld_acquire(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// all consumer threads of all CTAs (the grid is cooperative: every CTA is resident)
__device__ __forceinline__ void sm_grid_barrier(unsigned int* counter, unsigned int& epoch) {
  sm_sync();
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
    const unsigned int target = (epoch + 1) * gridDim.x;
    unsigned int spins = 0;
    while (sm_ld_acquire(counter) < target) {
      if (++spins > (1u << 26)) __trap();  // a lost CTA must not hang the GPU
    }
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

// rows of a weight matrix dealt to this CTA, and how many of them ride one ring stage
__device__ __forceinline__ void sm_my_rows(int N, int& n0, int& n1) {
  n0 = (int)((long long)N * blockIdx.x / gridDim.x);
  n1 = (int)((long long)N * (blockIdx.x + 1) / gridDim.x);
}
__device__ __forceinline__ int sm_rows_per_stage(int K, int d) { return K == d ? 8 : 2; }

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

// ---- consumer: dot products of one weight row segment with the B activation rows ------------------------------------
template <int B>
__device__ __forceinline__ void sm_dot(const uint4* __restrict__ wrow, const __nv_bfloat16* __restrict__ act, int act_ld, int n_seg,
                                       int lane, float (&acc)[kSmallMaxBatch]) {
#pragma unroll
  for (int b = 0; b < B; ++b) acc[b] = 0.0f;
#pragma unroll 1
  for (int i = 0; i < n_seg; ++i) {
    float wf[8];
    sm_unpack8(wrow[i * 32 + lane], wf);
#pragma unroll
    for (int b = 0; b < B; ++b) {
      float xf[8];
      sm_unpack8(*reinterpret_cast<const uint4*>(act + (size_t)b * act_ld + i * 256 + lane * 8), xf);
      float s = acc[b];
#pragma unroll
      for (int e = 0; e < 8; ++e) s = fmaf(wf[e], xf[e], s);
      acc[b] = s;
    }
  }
#pragma unroll
  for (int b = 0; b < B; ++b) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[b] += __shfl_xor_sync(0xffffffffu, acc[b], o);
  }
}

struct SmCtx {
  const SmallArgs* a;
  unsigned char* ring;
  __nv_bfloat16* act;
  float* s_p;
  float* s_part;   // [8][64]
  float* s_misc;   // 512 floats
  uint64_t *full, *empty;
  SmRing rg;
  unsigned int epoch;
  int tid, warp, lane;
};

// acts <- LayerNorm(x[b]) * gamma + beta as bf16, for every sequence (each CTA recomputes it: x is a few KB in L2)
__device__ __forceinline__ void sm_input_layernorm(SmCtx& c, const float* gamma, const float* beta, int d) {
  const SmallArgs& a = *c.a;
  const int per = d / kSmConsumers;  // 3, 4 or 5
  float v[kSmallMaxBatch][5];
  float* red = c.s_misc;  // [B][8]
#pragma unroll
  for (int b = 0; b < kSmallMaxBatch; ++b) {
    if (b < a.B) {
      float s = 0.0f;
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        v[b][i] = (i < per) ? __ldcg(a.x + (size_t)b * d + c.tid + i * kSmConsumers) : 0.0f;
        s += v[b][i];
      }
      s = warp_sum(s);
      if (c.lane == 0) red[b * 8 + c.warp] = s;
    }
  }
  sm_sync();
  float mean[kSmallMaxBatch];
#pragma unroll
  for (int b = 0; b < kSmallMaxBatch; ++b) {
    if (b < a.B) {
      float s = 0.0f;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += red[b * 8 + w];
      mean[b] = s / (float)d;
    }
  }
  sm_sync();
#pragma unroll
  for (int b = 0; b < kSmallMaxBatch; ++b) {
    if (b < a.B) {
      float s = 0.0f;
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        if (i < per) {
          v[b][i] -= mean[b];
          s += v[b][i] * v[b][i];
        }
      }
      s = warp_sum(s);
      if (c.lane == 0) red[b * 8 + c.warp] = s;
    }
  }
  sm_sync();
#pragma unroll
  for (int b = 0; b < kSmallMaxBatch; ++b) {
    if (b < a.B) {
      float s = 0.0f;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += red[b * 8 + w];
      const float rstd = rsqrtf(s / (float)d + 1e-5f);
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        if (i < per) {
          const int k = c.tid + i * kSmConsumers;
          c.act[(size_t)b * d + k] = __float2bfloat16(v[b][i] * rstd * __ldg(gamma + k) + __ldg(beta + k));
        }
      }
    }
  }
  sm_sync();
}

// acts <- src (B, K) bf16 written by other CTAs in the previous phase
__device__ __forceinline__ void sm_input_vector(SmCtx& c, const __nv_bfloat16* src, int K) {
  const int n16 = c.a->B * K / 8;
  const uint4* s = reinterpret_cast<const uint4*>(src);
  uint4* dst = reinterpret_cast<uint4*>(c.act);
  for (int i = c.tid; i < n16; i += kSmConsumers) dst[i] = __ldcg(s + i);
  sm_sync();
}

__device__ __forceinline__ void sm_epilogue(const SmCtx& c, int epi, int layer, int n, int b, float v, const float* bias, void* out,
                                            int N) {
  const SmallArgs& a = *c.a;
  const int d = a.d;
  if (epi == kEpiLogits) {
    a.logits[(size_t)b * a.logits_ld + n] = v;
    return;
  }
  v += __ldg(bias + n);
  if (epi == kEpiResid) {
    float* px = a.x + (size_t)b * d + n;
    *px = __ldcg(px) + v;
  } else if (epi == kEpiBf16) {
    static_cast<__nv_bfloat16*>(out)[(size_t)b * N + n] = __float2bfloat16(v);
  } else if (epi == kEpiGelu) {
    static_cast<__nv_bfloat16*>(out)[(size_t)b * N + n] = __float2bfloat16(gelu_fast(v));
  } else {  // fused q | k | v: q to the query buffer, k / v appended to the paged cache at this step's position
    if (n < d) {
      a.q[(size_t)b * d + n] = __float2bfloat16(v);
    } else {
      const int p = __ldg(a.pos + b);
      const int pshift = __ffs(a.page_size) - 1;
      const long long row = (long long)__ldg(a.block_table + b * a.max_pages + (p >> pshift)) * a.page_size + (p & (a.page_size - 1));
      __nv_bfloat16* pages = (n < 2 * d ? a.k_pages : a.v_pages) + (size_t)layer * a.layer_page_stride;
      pages[row * d + (n < 2 * d ? n - d : n - 2 * d)] = __float2bfloat16(v);
    }
  }
}

template <int B>
__device__ __forceinline__ void sm_gemv_chunks(SmCtx& c, int N, int K, int epi, int layer, const float* bias, void* out) {
  const SmallArgs& a = *c.a;
  int n0, n1;
  sm_my_rows(N, n0, n1);
  const int R = sm_rows_per_stage(K, a.d);
  const int wpr = kSmWarps / R;        // warps sharing a row (1 or 4)
  const int row_in = c.warp / wpr, kpart = c.warp - row_in * wpr;
  const int klen = K / wpr, n_seg = klen / 256;
  float* part = c.s_misc;              // [2][2 rows][4 parts][B]: double-buffered by chunk parity
  int parity = 0;
  for (int r = n0; r < n1; r += R) {
    const int rows = min(R, n1 - r);
    mbar_wait(&c.full[c.rg.stage], c.rg.phase);
    float acc[kSmallMaxBatch];
    if (row_in < rows) {
      const unsigned char* wrow = c.ring + c.rg.stage * kSmStageBytes + ((size_t)row_in * K + (size_t)kpart * klen) * 2;
      sm_dot<B>(reinterpret_cast<const uint4*>(wrow), c.act + kpart * klen, K, n_seg, c.lane, acc);
    }
    __syncwarp();
    if (c.lane == 0) mbar_arrive(&c.empty[c.rg.stage]);  // the stage can be refilled while the results are written
    c.rg.advance();
    float mine = 0.0f;
#pragma unroll
    for (int b = 0; b < B; ++b)
      if (c.lane == b) mine = acc[b];
    if (wpr == 1) {
      if (row_in < rows && c.lane < B) sm_epilogue(c, epi, layer, r + row_in, c.lane, mine, bias, out, N);
    } else {
      float* pp = part + parity * (2 * 4 * kSmallMaxBatch);
      if (row_in < rows && c.lane < B) pp[(row_in * 4 + kpart) * kSmallMaxBatch + c.lane] = mine;
      sm_sync();
      if (kpart == 0 && row_in < rows && c.lane < B) {
        const float* q = pp + row_in * 4 * kSmallMaxBatch + c.lane;
        const float v = (q[0] + q[kSmallMaxBatch]) + (q[2 * kSmallMaxBatch] + q[3 * kSmallMaxBatch]);
        sm_epilogue(c, epi, layer, r + row_in, c.lane, v, bias, out, N);
      }
      parity ^= 1;
    }
  }
}

// one projection phase: input vector(s) -> shared memory, then this CTA's output columns
__device__ __forceinline__ void sm_gemv(SmCtx& c, int in_kind, const float* ln_g, const float* ln_b, const __nv_bfloat16* vec, int N,
                                        int K, int epi, int layer, const float* bias, void* out) {
  if (in_kind == kInLayerNorm) sm_input_layernorm(c, ln_g, ln_b, K);
  else sm_input_vector(c, vec, K);
  switch (c.a->B) {
    case 1: sm_gemv_chunks<1>(c, N, K, epi, layer, bias, out); break;
    case 2: sm_gemv_chunks<2>(c, N, K, epi, layer, bias, out); break;
    case 3: sm_gemv_chunks<3>(c, N, K, epi, layer, bias, out); break;
    case 4: sm_gemv_chunks<4>(c, N, K, epi, layer, bias, out); break;
    case 5: sm_gemv_chunks<5>(c, N, K, epi, layer, bias, out); break;
    default: sm_gemv_chunks<6>(c, N, K, epi, layer, bias, out); break;
  }
}

// ---- attention over `T` key rows of 64 dims (K rows at kbase + j * ld, V rows at vbase + j * ld) for one query head --------
// 8 lanes per key row (16 B each), 4 rows per warp per load instruction, 8 loads in flight per thread; returns the
// chunk's (max, sum) and leaves the unnormalised output of dim `tid` (< 64) in s_part-reduced form in `o_out`.
template <typename KRow, typename VRow>
__device__ __forceinline__ void sm_attend(SmCtx& c, const float (&qv)[8], int T, KRow krow, VRow vrow, float& mx_out, float& sum_out,
                                          float& o_out) {
  const int sub = c.lane & 7, kg = c.lane >> 3;
  constexpr int kU = 8, kStep = kSmWarps * 4;
  float mx = -INFINITY;
  for (int j0 = c.warp * 4; j0 < T; j0 += kU * kStep) {  // warp-uniform trip count: the shuffles need every lane
    uint4 u[kU];
#pragma unroll
    for (int i = 0; i < kU; ++i) {
      const int j = j0 + kg + i * kStep;
      u[i] = __ldcg(reinterpret_cast<const uint4*>(krow(min(j, T - 1))) + sub);
    }
#pragma unroll
    for (int i = 0; i < kU; ++i) {
      const int j = j0 + kg + i * kStep;
      float f[8];
      sm_unpack8(u[i], f);
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
  }
  float* red = c.s_misc + 256;
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
  for (int j0 = c.warp * 4; j0 < T; j0 += kU * kStep) {
    uint4 u[kU];
    float p[kU];
#pragma unroll
    for (int i = 0; i < kU; ++i) {
      const int j = j0 + kg + i * kStep;
      u[i] = __ldcg(reinterpret_cast<const uint4*>(vrow(min(j, T - 1))) + sub);
      p[i] = (j < T) ? c.s_p[j] : 0.0f;
    }
#pragma unroll
    for (int i = 0; i < kU; ++i) {
      float f[8];
      sm_unpack8(u[i], f);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = fmaf(p[i], f[e], acc[e]);
    }
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
    const int* bt = a.block_table + b * a.max_pages;
    const __nv_bfloat16* kp = a.k_pages + (size_t)layer * a.layer_page_stride + h * kSmHd;
    const __nv_bfloat16* vp = a.v_pages + (size_t)layer * a.layer_page_stride + h * kSmHd;
    float qv[8];
    sm_load_q(c, a.q + (size_t)b * d + h * kSmHd, qv);
    auto krow = [&](int j) { return kp + ((long long)__ldg(bt + (j >> pshift)) * a.page_size + (j & (a.page_size - 1))) * d; };
    auto vrow = [&](int j) { return vp + ((long long)__ldg(bt + (j >> pshift)) * a.page_size + (j & (a.page_size - 1))) * d; };
    float mx, sum, o;
    sm_attend(c, qv, n_keys, krow, vrow, mx, sum, o);
    if (c.tid < kSmHd) a.att[(size_t)b * d + h * kSmHd + c.tid] = __float2bfloat16(o / sum);
  }
}

__device__ __forceinline__ void sm_cross_attention(SmCtx& c, int layer) {
  const SmallArgs& a = *c.a;
  const int d = a.d, H = a.n_head, T_all = a.T;
  const int units = a.B * H;
  int splits = (int)gridDim.x / units;
  splits = splits < 1 ? 1 : (splits > 8 ? 8 : splits);
  const int per = ((T_all + splits - 1) / splits + 31) & ~31;
  const long long ld = 2ll * d;
  int* s_last = reinterpret_cast<int*>(c.s_misc + 300);
  for (int w = blockIdx.x; w < units * splits; w += gridDim.x) {
    const int unit = w / splits, chunk = w - unit * splits;
    const int b = unit / H, h = unit - b * H;
    if (a.finished != nullptr && a.finished[b]) continue;
    const int k0 = min(chunk * per, T_all), T = min(per, T_all - k0);
    const __nv_bfloat16* kb = a.cross_kv + (size_t)layer * a.cross_layer_stride + (long long)__ldg(a.cross_slot + b) * a.cross_seq_stride +
                              (long long)k0 * ld + h * kSmHd;
    float qv[8];
    sm_load_q(c, a.qc + (size_t)b * d + h * kSmHd, qv);
    float mx = -INFINITY, sum = 0.0f, o = 0.0f;
    if (T > 0) {
      auto krow = [&](int j) { return kb + j * ld; };
      auto vrow = [&](int j) { return kb + d + j * ld; };
      sm_attend(c, qv, T, krow, vrow, mx, sum, o);
    }
    if (splits == 1) {
      if (c.tid < kSmHd) a.att[(size_t)b * d + h * kSmHd + c.tid] = __float2bfloat16(o / sum);
      continue;
    }
    // (max, sum, unnormalised output) of this chunk; the last chunk of the unit to arrive merges them
    float* mine = a.ca_part + ((size_t)unit * splits + chunk) * kSmPartFloats;
    if (c.tid < kSmHd) mine[2 + c.tid] = o;
    if (c.tid == 0) {
      mine[0] = mx;
      mine[1] = sum;
    }
    __threadfence();
    sm_sync();
    if (c.tid == 0) *s_last = (atomicAdd(a.ca_cnt + unit, 1) == splits - 1) ? 1 : 0;
    sm_sync();
    if (*s_last) {
      __threadfence();
      if (c.tid < kSmHd) {
        const float* all = a.ca_part + (size_t)unit * splits * kSmPartFloats;
        float M = -INFINITY;
        for (int k = 0; k < splits; ++k) M = fmaxf(M, __ldcg(all + k * kSmPartFloats));
        float L = 0.0f, acc = 0.0f;
        for (int k = 0; k < splits; ++k) {
          const float w2 = sm_exp2(__ldcg(all + k * kSmPartFloats) - M);
          L = fmaf(__ldcg(all + k * kSmPartFloats + 1), w2, L);
          acc = fmaf(__ldcg(all + k * kSmPartFloats + 2 + c.tid), w2, acc);
        }
        a.att[(size_t)b * d + h * kSmHd + c.tid] = __float2bfloat16(acc / L);
      }
      if (c.tid == 0) a.ca_cnt[unit] = 0;  // ready for the next layer
    }
    sm_sync();
  }
}

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
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSmRingBytes + kSmActBytes + kSmSpBytes + kSmPartBytes + kSmMiscBytes - 128);
  c.full = bars;
  c.empty = bars + kSmStages;
  c.rg = SmRing{0, 0};
  c.epoch = 0;
  c.tid = threadIdx.x;
  c.warp = threadIdx.x >> 5;
  c.lane = threadIdx.x & 31;
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
    for (int l = 0; l < L; ++l) {
      const b200w_dec_layer& W = a.layers[l];
      sm_gemv(c, kInLayerNorm, W.attn_ln_g, W.attn_ln_b, nullptr, 3 * d, d, kEpiQkv, l, W.b_qkv, nullptr);
      sm_grid_barrier(a.counter, c.epoch);
      sm_self_attention(c, l);
      sm_grid_barrier(a.counter, c.epoch);
      sm_gemv(c, kInVector, nullptr, nullptr, a.att, d, d, kEpiResid, l, W.b_out, nullptr);
      sm_grid_barrier(a.counter, c.epoch);
      sm_gemv(c, kInLayerNorm, W.cross_ln_g, W.cross_ln_b, nullptr, d, d, kEpiBf16, l, W.b_cq, a.qc);
      sm_grid_barrier(a.counter, c.epoch);
      sm_cross_attention(c, l);
      sm_grid_barrier(a.counter, c.epoch);
      sm_gemv(c, kInVector, nullptr, nullptr, a.att, d, d, kEpiResid, l, W.b_cout, nullptr);
      sm_grid_barrier(a.counter, c.epoch);
      sm_gemv(c, kInLayerNorm, W.mlp_ln_g, W.mlp_ln_b, nullptr, 4 * d, d, kEpiGelu, l, W.b_mlp1, a.mlp);
      sm_grid_barrier(a.counter, c.epoch);
      sm_gemv(c, kInVector, nullptr, nullptr, a.mlp, d, 4 * d, kEpiResid, l, W.b_mlp2, nullptr);
      sm_grid_barrier(a.counter, c.epoch);
    }
    // final LayerNorm + tied logits of the (single) new token of every sequence
    sm_gemv(c, kInLayerNorm, a.dec_ln_g, a.dec_ln_b, nullptr, a.n_vocab, d, kEpiLogits, 0, nullptr, nullptr);
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------- host
static bool g_small_ready = false;

int init_decode_small() {
  if (g_small_ready) return kOk;
  B200W_CUDA_OK(cudaFuncSetAttribute(decode_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmSmemBytes));
  g_small_ready = true;
  return kOk;
}

bool decode_small_applicable(const b200w_dims& dm, int n_seq, int n_q) {
  const int d = dm.n_text_state;
  return n_q == 1 && n_seq >= 1 && n_seq <= kSmallMaxBatch && d % kSmConsumers == 0 && d / kSmConsumers <= 5 && 4 * d <= kSmMaxK &&
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
  B200W_CUDA_OK(cudaLaunchKernelEx(&cfg, decode_small_kernel, a));
  count_launch();
  return kOk;
}

}  // namespace b200w
