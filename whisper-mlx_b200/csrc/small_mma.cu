// K13m: the one-launch single-token decoder step (K13, small.cu) for batches of 4 .. 16 sequences, with the projections
// on the tensor cores.
//
// K13 wins at <= 3 sequences because a step there is a chain of small phases whose weights are already on chip (a
// producer warp streams every weight row the CTA will need through a shared-memory ring, across phase boundaries) and
// whose cost is the exchange between phases, not the arithmetic.  Its matrix-vector products run on the FP32 pipes and
// grow by ~0.5 ms per extra sequence; the large-batch path (K11 chains + attention kernels) costs 2.6 ms per step at
// ANY batch <= 16 -- what an 8-GPU shard of the reference's hour (15 windows per GPU) or a `best_of` fallback pays.
// K13m keeps K13's structure and gives the dot products to `mma.sync.m16n8k16` (legacy warp-level MMA on purpose: a
// 12-row x 1280 weight chunk against <= 16 activation rows is 80 k-steps of one m16n8k16 -- far too small for a
// tcgen05 tile, and what it needs is few instructions per weight byte, not FLOPs):
//   * A operand = the weight rows of a ring stage (12 rows, row-major, padded by 16 bytes so that the 8 rows a
//     fragment touches fall into different banks), B operand = the activation rows (n = sequence, padded likewise);
//     the 80 k-steps of a stage are dealt to the 12 warps, partial accumulators meet in shared memory;
//   * K = 4d projections (MLP2) walk the row's four quarters as four stages accumulating into the same fragment, with
//     the activation quarter reloaded in between (activations of 16 sequences x 5120 would not fit beside the ring);
//   * LayerNorm: one warp per sequence (warp-level statistics only), gamma / beta prefetched a phase early into shared
//     memory with cp.async; attention phases, barriers, residual ownership, epilogues as in K13.
#include "common.cuh"
#include "kernels.h"

namespace b200w {

constexpr int kMmMaxBatch = 16;
constexpr int kMmWarps = 12;
constexpr int kMmConsumers = kMmWarps * 32;
constexpr int kMmThreads = kMmConsumers + 32;  // + the producer warp
constexpr int kMmMaxD = 1280;
constexpr int kMmPad = 16;                                  // bytes of padding per weight / activation row
constexpr int kMmRowStrideMax = kMmMaxD * 2 + kMmPad;       // 2576
constexpr int kMmStageRows = 12;
constexpr int kMmStageBytes = kMmStageRows * kMmRowStrideMax;  // 30912
constexpr int kMmStages = 4;
constexpr int kMmHd = 64;
constexpr int kMmMaxKeys = 1536;
constexpr int kMmMaxPages = 32;
constexpr float kMmLog2e = 1.4426950408889634f;
constexpr int kMmOwnMax = 16;

constexpr int kMmRingBytes = kMmStages * kMmStageBytes;
constexpr int kMmActBytes = kMmMaxBatch * kMmRowStrideMax;
constexpr int kMmRedBytes = kMmWarps * 2 * 128 * 4;  // [warp][n-tile][16 x 8] f32
constexpr int kMmLnBytes = 2 * kMmMaxD * 4;          // gamma | beta of the next LayerNorm
constexpr int kMmSpBytes = kMmMaxKeys * 4;
constexpr int kMmPartBytes = kMmWarps * kMmHd * 4;
constexpr int kMmMiscBytes = 8192;
constexpr int kMmSmemBytes = kMmRingBytes + kMmActBytes + kMmRedBytes + kMmLnBytes + kMmSpBytes + kMmPartBytes + kMmMiscBytes + 128;
static_assert(kMmSmemBytes <= 232448, "shared memory budget of one CTA per SM");
// float offsets inside the misc area
constexpr int kMmAttRed = 0;     // [12]
constexpr int kMmXown = 32;      // [16][kMmOwnMax]
constexpr int kMmKvRow = 288;    // [16] long long
constexpr int kMmBias = 320;     // [48]
constexpr int kMmBiasMax = 48;
constexpr int kMmBt = 384;       // [16][kMmMaxPages] int
constexpr int kMmBars = 1024;    // mbarriers
static_assert(kMmXown + kMmMaxBatch * kMmOwnMax <= kMmKvRow && kMmKvRow + 2 * kMmMaxBatch <= kMmBias && kMmBias + kMmBiasMax <= kMmBt &&
              kMmBt + kMmMaxBatch * kMmMaxPages <= kMmBars && kMmBars + 32 <= kMmMiscBytes / 4, "misc layout");

enum MmIn { kMmInLayerNorm = 0, kMmInVector = 1 };
enum MmEpi { kMmEpiQkv = 0, kMmEpiResid = 1, kMmEpiBf16 = 2, kMmEpiGelu = 3, kMmEpiLogits = 4 };

__device__ __forceinline__ void mm_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kMmConsumers) : "memory"); }

__device__ __forceinline__ unsigned int mm_ld_acquire(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// all consumer threads of all CTAs (cooperative launch: every CTA is resident)
__device__ __forceinline__ void mm_grid_barrier(unsigned int* counter, unsigned int& epoch) {
  mm_sync();
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
    const unsigned int target = (epoch + 1) * gridDim.x;
    unsigned int spins = 0;
    while (mm_ld_acquire(counter) < target) {
      if (++spins > (1u << 26)) __trap();  // a lost CTA must not hang the GPU
    }
  }
  ++epoch;
  mm_sync();
}

__device__ __forceinline__ void mm_bulk_load(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mm_cp16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void mm_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void mm_cp_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ float mm_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void mm_unpack8(const uint4& u, float (&f)[8]) {
  f[0] = __uint_as_float(u.x << 16); f[1] = __uint_as_float(u.x & 0xffff0000u);
  f[2] = __uint_as_float(u.y << 16); f[3] = __uint_as_float(u.y & 0xffff0000u);
  f[4] = __uint_as_float(u.z << 16); f[5] = __uint_as_float(u.z & 0xffff0000u);
  f[6] = __uint_as_float(u.w << 16); f[7] = __uint_as_float(u.w & 0xffff0000u);
}
__device__ __forceinline__ void mm_mma(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

struct MmRing {
  int stage;
  uint32_t phase;
  __device__ __forceinline__ void advance() {
    if (++stage == kMmStages) {
      stage = 0;
      phase ^= 1;
    }
  }
};

__device__ __forceinline__ void mm_my_rows(int N, int& n0, int& n1) {
  n0 = (int)((long long)N * blockIdx.x / gridDim.x);
  n1 = (int)((long long)N * (blockIdx.x + 1) / gridDim.x);
}

// ---- producer: this CTA's rows of W (N, K) bf16 in chunks of d elements: K == d one chunk per row, K == 4d the row's
// four quarters as four consecutive stages of the same row group --------------------------------------------------------
__device__ __forceinline__ void mm_produce(const void* W, int N, int K, int d, unsigned char* ring, uint64_t* full, uint64_t* empty,
                                           MmRing& rg) {
  int n0, n1;
  mm_my_rows(N, n0, n1);
  const int nq = K / d, rs = d * 2 + kMmPad;
  const uint32_t chunk = (uint32_t)d * 2u;
  const unsigned char* base = static_cast<const unsigned char*>(W);
  for (int r = n0; r < n1; r += kMmStageRows) {
    const int rows = min(kMmStageRows, n1 - r);
    for (int q = 0; q < nq; ++q) {
      mbar_wait(&empty[rg.stage], rg.phase ^ 1);
      mbar_expect_tx(&full[rg.stage], (uint32_t)rows * chunk);
      unsigned char* dst = ring + rg.stage * kMmStageBytes;
      for (int i = 0; i < rows; ++i)
        mm_bulk_load(dst + i * rs, base + ((size_t)(r + i) * K + (size_t)q * d) * 2, chunk, &full[rg.stage]);
      rg.advance();
    }
  }
}

struct MmCtx {
  const SmallArgs* a;
  unsigned char* ring;
  unsigned char* act;  // [16][d * 2 + 16 bytes] bf16 activation rows
  float* s_red;        // [12][NT][128]
  float* s_ln;         // gamma[d] | beta[d] of the next LayerNorm phase
  float* s_p;
  float* s_part;
  float* s_misc;
  float* s_xown;
  long long* s_kvrow;
  int* s_bt;
  uint64_t *full, *empty;
  MmRing rg;
  unsigned int epoch;
  int tid, warp, lane;
  int own_n0, B;
};

// the LayerNorm parameters of the phase after this one, into shared memory (the copies are awaited by that phase)
__device__ __forceinline__ void mm_prefetch_ln(MmCtx& c, int d, const float* g, const float* b) {
  if (g != nullptr && c.tid < d / 4) {
    mm_cp16(c.s_ln + c.tid * 4, g + c.tid * 4);
    mm_cp16(c.s_ln + kMmMaxD + c.tid * 4, b + c.tid * 4);
  }
  mm_cp_commit();
}

// act <- LayerNorm(x[b]) * gamma + beta as bf16: one warp per sequence, V = d / 128 float4 per lane
template <int V>
__device__ __forceinline__ void mm_input_layernorm(MmCtx& c, int d, const float* next_g, const float* next_b) {
  const SmallArgs& a = *c.a;
  const int rs = d * 2 + kMmPad;
  mm_cp_wait_all();
  mm_sync();  // every thread's share of gamma / beta has landed
  for (int b = c.warp; b < c.B; b += kMmWarps) {
    float4 v[V];
#pragma unroll
    for (int i = 0; i < V; ++i) v[i] = __ldcg(reinterpret_cast<const float4*>(a.x + (size_t)b * d) + i * 32 + c.lane);
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < V; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    const float mean = warp_sum(s) / (float)d;
    float q = 0.0f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
      q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)d + 1e-5f);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float4 g = *reinterpret_cast<const float4*>(c.s_ln + (i * 32 + c.lane) * 4);
      const float4 be = *reinterpret_cast<const float4*>(c.s_ln + kMmMaxD + (i * 32 + c.lane) * 4);
      *reinterpret_cast<uint2*>(c.act + (size_t)b * rs + (i * 32 + c.lane) * 8) =
          make_uint2(pack_bf16x2(v[i].x * rstd * g.x + be.x, v[i].y * rstd * g.y + be.y),
                     pack_bf16x2(v[i].z * rstd * g.z + be.z, v[i].w * rstd * g.w + be.w));
    }
  }
  mm_sync();
  mm_prefetch_ln(c, d, next_g, next_b);
}

// act <- src[b][0 .. d) (rows ld_src elements apart) written by other CTAs in an earlier phase
__device__ __forceinline__ void mm_input_vector(MmCtx& c, const __nv_bfloat16* src, long long ld_src, int d) {
  const int per = d / 8, rs = d * 2 + kMmPad;
  for (int i = c.tid; i < c.B * per; i += kMmConsumers) {
    const int b = i / per, j = i - b * per;
    *reinterpret_cast<uint4*>(c.act + (size_t)b * rs + j * 16) = __ldcg(reinterpret_cast<const uint4*>(src + (size_t)b * ld_src) + j);
  }
  mm_sync();
}

// value v (bias already added) of output column n for sequence b
__device__ __forceinline__ void mm_epilogue(const MmCtx& c, int epi, int layer, int n, int b, float v, void* out, int N) {
  const SmallArgs& a = *c.a;
  const int d = a.d;
  if (epi == kMmEpiLogits) {
    a.logits[(size_t)b * a.logits_ld + n] = v;
  } else if (epi == kMmEpiResid) {
    float* own = c.s_xown + b * kMmOwnMax + (n - c.own_n0);
    const float xn = *own + v;
    *own = xn;
    a.x[(size_t)b * d + n] = xn;
  } else if (epi == kMmEpiBf16) {
    static_cast<__nv_bfloat16*>(out)[(size_t)b * N + n] = __float2bfloat16(v);
  } else if (epi == kMmEpiGelu) {
    static_cast<__nv_bfloat16*>(out)[(size_t)b * N + n] = __float2bfloat16(gelu_fast(v));
  } else {  // fused q | k | v
    if (n < d) {
      a.q[(size_t)b * d + n] = __float2bfloat16(v);
    } else {
      __nv_bfloat16* pages = (n < 2 * d ? a.k_pages : a.v_pages) + (size_t)layer * a.layer_page_stride;
      pages[c.s_kvrow[b] * d + (n < 2 * d ? n - d : n - 2 * d)] = __float2bfloat16(v);
    }
  }
}

// one projection phase.  NT = n-tiles of 8 sequences, S = d / 256 (16 * S k-steps per chunk of d).
template <int NT, int S>
__device__ __forceinline__ void mm_gemv(MmCtx& c, int in_kind, const float* next_g, const float* next_b, const __nv_bfloat16* vec, int N, int K,
                                        int epi, int layer, const float* bias, void* out) {
  const int d = c.a->d, rs = d * 2 + kMmPad;
  int n0, n1;
  mm_my_rows(N, n0, n1);
  if (bias != nullptr && c.tid < n1 - n0) c.s_misc[kMmBias + c.tid] = __ldg(bias + n0 + c.tid);
  const int nq = K / d;
  if (nq == 1) {
    if (in_kind == kMmInLayerNorm) mm_input_layernorm<2 * S>(c, d, next_g, next_b);
    else mm_input_vector(c, vec, K, d);
  }
  const int g = c.lane >> 2, tg = c.lane & 3;
  const float* s_bias = c.s_misc + kMmBias;
  for (int r = n0; r < n1; r += kMmStageRows) {
    const int rows = min(kMmStageRows, n1 - r);
    float acc[NT][4];
#pragma unroll
    for (int t = 0; t < NT; ++t)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[t][i] = 0.0f;
    for (int q = 0; q < nq; ++q) {
      if (nq > 1) {
        mm_sync();  // the previous quarter's fragments have been read
        mm_input_vector(c, vec + (size_t)q * d, K, d);
      }
      mbar_wait(&c.full[c.rg.stage], c.rg.phase);
      const unsigned char* wa = c.ring + c.rg.stage * kMmStageBytes + g * rs + tg * 4;
      const int hi = min(g + 8, kMmStageRows - 1) - g;  // rows 12 .. 15 of the fragment do not exist: re-read row 11
      const unsigned char* xb = c.act + g * rs + tg * 4;
      for (int ks = c.warp; ks < 16 * S; ks += kMmWarps) {
        const int kb = ks * 32;  // bytes
        const uint32_t a0 = *reinterpret_cast<const uint32_t*>(wa + kb), a1 = *reinterpret_cast<const uint32_t*>(wa + hi * rs + kb);
        const uint32_t a2 = *reinterpret_cast<const uint32_t*>(wa + kb + 16), a3 = *reinterpret_cast<const uint32_t*>(wa + hi * rs + kb + 16);
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          const uint32_t b0 = *reinterpret_cast<const uint32_t*>(xb + t * 8 * rs + kb), b1 = *reinterpret_cast<const uint32_t*>(xb + t * 8 * rs + kb + 16);
          mm_mma(acc[t], a0, a1, a2, a3, b0, b1);
        }
      }
      __syncwarp();
      if (c.lane == 0) mbar_arrive(&c.empty[c.rg.stage]);
      c.rg.advance();
    }
    // the 12 warps' partial fragments meet in shared memory: [warp][n-tile][16 rows x 8 sequences]
    float* mine = c.s_red + (c.warp * NT) * 128;
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      *reinterpret_cast<float2*>(mine + t * 128 + g * 8 + tg * 2) = make_float2(acc[t][0], acc[t][1]);
      *reinterpret_cast<float2*>(mine + t * 128 + (g + 8) * 8 + tg * 2) = make_float2(acc[t][2], acc[t][3]);
    }
    mm_sync();
    for (int i = c.tid; i < rows * c.B; i += kMmConsumers) {
      const int row = i / c.B, b = i - row * c.B;
      const float* p = c.s_red + (b >> 3) * 128 + row * 8 + (b & 7);
      float v = 0.0f;
#pragma unroll
      for (int w = 0; w < kMmWarps; ++w) v += p[w * NT * 128];
      if (bias != nullptr) v += s_bias[r + row - n0];
      mm_epilogue(c, epi, layer, r + row, b, v, out, N);
    }
    mm_sync();  // s_red is rewritten by the next row group
  }
}

// ---- attention over `T` key rows of 64 dims for one query head (K13's routine) -----------------------------------------
template <typename KRow, typename VRow>
__device__ __forceinline__ void mm_attend(MmCtx& c, const float (&qv)[8], int T, KRow krow, VRow vrow, float& mx_out, float& sum_out,
                                          float& o_out) {
  const int sub = c.lane & 7, kg = c.lane >> 3;
  constexpr int kU = 5, kStep = kMmWarps * 4, kSweep = kU * kStep;
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
        mm_unpack8(uu[i], f);
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
    for (int j0 = first + kSweep; j0 < T + first; j0 += kSweep) {  // warp-uniform trip count
      if (j0 - first >= T) break;
#pragma unroll
      for (int i = 0; i < kU; ++i) u[i] = __ldcg(reinterpret_cast<const uint4*>(krow(min(j0 + kg + i * kStep, T - 1))) + sub);
      score(u, j0);
    }
  }
  float* red = c.s_misc + kMmAttRed;
  mx = warp_max(mx);
  if (c.lane == 0) red[c.warp] = mx;
  mm_sync();
  mx = red[0];
#pragma unroll
  for (int i = 1; i < kMmWarps; ++i) mx = fmaxf(mx, red[i]);
  mm_sync();
  float sum = 0.0f;
  for (int j = c.tid; j < T; j += kMmConsumers) {
    const float p = mm_exp2(c.s_p[j] - mx);
    sum += p;
    c.s_p[j] = __bfloat162float(__float2bfloat16(p));
  }
  sum = warp_sum(sum);
  if (c.lane == 0) red[c.warp] = sum;
  mm_sync();
  sum = 0.0f;
#pragma unroll
  for (int i = 0; i < kMmWarps; ++i) sum += red[i];
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.0f;
  auto accumulate = [&](const uint4 (&uu)[kU], int j0) {
#pragma unroll
    for (int i = 0; i < kU; ++i) {
      const int j = j0 + kg + i * kStep;
      const float p = (j < T) ? c.s_p[j] : 0.0f;
      float f[8];
      mm_unpack8(uu[i], f);
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
    for (int i = 0; i < 8; ++i) c.s_part[c.warp * kMmHd + sub * 8 + i] = acc[i];
  }
  mm_sync();
  float o = 0.0f;
  if (c.tid < kMmHd) {
#pragma unroll
    for (int w = 0; w < kMmWarps; ++w) o += c.s_part[w * kMmHd + c.tid];
  }
  mx_out = mx;
  sum_out = sum;
  o_out = o;
  mm_sync();
}

__device__ __forceinline__ void mm_load_q(const MmCtx& c, const __nv_bfloat16* q, float (&qv)[8]) {
  float f[8];
  mm_unpack8(__ldcg(reinterpret_cast<const uint4*>(q) + (c.lane & 7)), f);
  const float s = 0.125f * kMmLog2e;
#pragma unroll
  for (int i = 0; i < 8; ++i) qv[i] = f[i] * s;
}

__device__ __forceinline__ void mm_self_attention(MmCtx& c, int layer) {
  const SmallArgs& a = *c.a;
  const int d = a.d, H = a.n_head;
  const int pshift = __ffs(a.page_size) - 1;
  for (int u = blockIdx.x; u < c.B * H; u += gridDim.x) {
    const int b = u / H, h = u - b * H;
    if (a.finished != nullptr && a.finished[b]) continue;  // (CTA-uniform)
    const int n_keys = __ldg(a.pos + b) + 1;
    const int* bt = c.s_bt + b * kMmMaxPages;
    const __nv_bfloat16* kp = a.k_pages + (size_t)layer * a.layer_page_stride + h * kMmHd;
    const __nv_bfloat16* vp = a.v_pages + (size_t)layer * a.layer_page_stride + h * kMmHd;
    float qv[8];
    mm_load_q(c, a.q + (size_t)b * d + h * kMmHd, qv);
    auto krow = [&](int j) { return kp + ((long long)bt[j >> pshift] * a.page_size + (j & (a.page_size - 1))) * d; };
    auto vrow = [&](int j) { return vp + ((long long)bt[j >> pshift] * a.page_size + (j & (a.page_size - 1))) * d; };
    float mx, sum, o;
    mm_attend(c, qv, n_keys, krow, vrow, mx, sum, o);
    if (c.tid < kMmHd) a.att[(size_t)b * d + h * kMmHd + c.tid] = __float2bfloat16(o / sum);
  }
}

__device__ __forceinline__ void mm_cross_attention(MmCtx& c, int layer) {
  const SmallArgs& a = *c.a;
  const int d = a.d, H = a.n_head, T = a.T;
  const long long ld = 2ll * d;
  for (int u = blockIdx.x; u < c.B * H; u += gridDim.x) {
    const int b = u / H, h = u - b * H;
    if (a.finished != nullptr && a.finished[b]) continue;
    const __nv_bfloat16* kb = a.cross_kv + (size_t)layer * a.cross_layer_stride + (long long)__ldg(a.cross_slot + b) * a.cross_seq_stride +
                              h * kMmHd;
    float qv[8];
    mm_load_q(c, a.qc + (size_t)b * d + h * kMmHd, qv);
    auto krow = [&](int j) { return kb + j * ld; };
    auto vrow = [&](int j) { return kb + d + j * ld; };
    float mx, sum, o;
    mm_attend(c, qv, T, krow, vrow, mx, sum, o);
    if (c.tid < kMmHd) a.att[(size_t)b * d + h * kMmHd + c.tid] = __float2bfloat16(o / sum);
  }
}

template <int NT, int S>
__device__ __forceinline__ void mm_consumer(MmCtx& c) {
  const SmallArgs& a = *c.a;
  const int d = a.d, L = a.n_layer, B = c.B;
  {
    int n1;
    mm_my_rows(d, c.own_n0, n1);
    const int cnt = n1 - c.own_n0;
    for (int i = c.tid; i < B * cnt; i += kMmConsumers) {
      const int b = i / cnt, j = i - b * cnt;
      c.s_xown[b * kMmOwnMax + j] = a.x[(size_t)b * d + c.own_n0 + j];
    }
    for (int i = c.tid; i < B * a.max_pages; i += kMmConsumers) {
      const int b = i / a.max_pages, j = i - b * a.max_pages;
      c.s_bt[b * kMmMaxPages + j] = a.block_table[b * a.max_pages + j];
    }
    if (c.tid < B) {
      const int p = a.pos[c.tid];
      const int pshift = __ffs(a.page_size) - 1;
      c.s_kvrow[c.tid] = (long long)a.block_table[c.tid * a.max_pages + (p >> pshift)] * a.page_size + (p & (a.page_size - 1));
    }
    mm_prefetch_ln(c, d, a.layers[0].attn_ln_g, a.layers[0].attn_ln_b);
    mm_sync();
  }
  for (int l = 0; l < L; ++l) {
    const b200w_dec_layer& W = a.layers[l];
    mm_gemv<NT, S>(c, kMmInLayerNorm, W.cross_ln_g, W.cross_ln_b, nullptr, 3 * d, d, kMmEpiQkv, l, W.b_qkv, nullptr);
    mm_grid_barrier(a.counter, c.epoch);
    mm_self_attention(c, l);
    mm_grid_barrier(a.counter, c.epoch);
    mm_gemv<NT, S>(c, kMmInVector, nullptr, nullptr, a.att, d, d, kMmEpiResid, l, W.b_out, nullptr);
    mm_grid_barrier(a.counter, c.epoch);
    mm_gemv<NT, S>(c, kMmInLayerNorm, W.mlp_ln_g, W.mlp_ln_b, nullptr, d, d, kMmEpiBf16, l, W.b_cq, a.qc);
    mm_grid_barrier(a.counter, c.epoch);
    mm_cross_attention(c, l);
    mm_grid_barrier(a.counter, c.epoch);
    mm_gemv<NT, S>(c, kMmInVector, nullptr, nullptr, a.att, d, d, kMmEpiResid, l, W.b_cout, nullptr);
    mm_grid_barrier(a.counter, c.epoch);
    mm_gemv<NT, S>(c, kMmInLayerNorm, l + 1 < L ? a.layers[l + 1].attn_ln_g : a.dec_ln_g, l + 1 < L ? a.layers[l + 1].attn_ln_b : a.dec_ln_b,
                   nullptr, 4 * d, d, kMmEpiGelu, l, W.b_mlp1, a.mlp);
    mm_grid_barrier(a.counter, c.epoch);
    mm_gemv<NT, S>(c, kMmInVector, nullptr, nullptr, a.mlp, d, 4 * d, kMmEpiResid, l, W.b_mlp2, nullptr);
    mm_grid_barrier(a.counter, c.epoch);
  }
  mm_gemv<NT, S>(c, kMmInLayerNorm, nullptr, nullptr, nullptr, a.n_vocab, d, kMmEpiLogits, 0, nullptr, nullptr);
}

template <int NT, int S>
__global__ void __launch_bounds__(kMmThreads, 1) decode_small_mma_kernel(const __grid_constant__ SmallArgs a) {
  extern __shared__ unsigned char mm_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(mm_raw) + 127) & ~static_cast<uintptr_t>(127));
  MmCtx c;
  c.a = &a;
  c.ring = smem;
  c.act = smem + kMmRingBytes;
  c.s_red = reinterpret_cast<float*>(smem + kMmRingBytes + kMmActBytes);
  c.s_ln = reinterpret_cast<float*>(smem + kMmRingBytes + kMmActBytes + kMmRedBytes);
  c.s_p = reinterpret_cast<float*>(smem + kMmRingBytes + kMmActBytes + kMmRedBytes + kMmLnBytes);
  c.s_part = reinterpret_cast<float*>(smem + kMmRingBytes + kMmActBytes + kMmRedBytes + kMmLnBytes + kMmSpBytes);
  c.s_misc = reinterpret_cast<float*>(smem + kMmRingBytes + kMmActBytes + kMmRedBytes + kMmLnBytes + kMmSpBytes + kMmPartBytes);
  c.s_xown = c.s_misc + kMmXown;
  c.s_kvrow = reinterpret_cast<long long*>(c.s_misc + kMmKvRow);
  c.s_bt = reinterpret_cast<int*>(c.s_misc + kMmBt);
  uint64_t* bars = reinterpret_cast<uint64_t*>(c.s_misc + kMmBars);
  c.full = bars;
  c.empty = bars + kMmStages;
  c.rg = MmRing{0, 0};
  c.epoch = 0;
  c.tid = threadIdx.x;
  c.warp = threadIdx.x >> 5;
  c.lane = threadIdx.x & 31;
  c.own_n0 = 0;
  c.B = a.B;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kMmStages; ++i) {
      mbar_init(&c.full[i], 1);
      mbar_init(&c.empty[i], kMmWarps);
    }
    fence_barrier_init();
  }
  __syncthreads();

  const int d = a.d, L = a.n_layer;
  if (c.warp == kMmWarps) {
    if (c.lane == 0) {
      for (int l = 0; l < L; ++l) {
        const b200w_dec_layer& W = a.layers[l];
        mm_produce(W.w_qkv, 3 * d, d, d, c.ring, c.full, c.empty, c.rg);
        mm_produce(W.w_out, d, d, d, c.ring, c.full, c.empty, c.rg);
        mm_produce(W.w_cq, d, d, d, c.ring, c.full, c.empty, c.rg);
        mm_produce(W.w_cout, d, d, d, c.ring, c.full, c.empty, c.rg);
        mm_produce(W.w_mlp1, 4 * d, d, d, c.ring, c.full, c.empty, c.rg);
        mm_produce(W.w_mlp2, d, 4 * d, d, c.ring, c.full, c.empty, c.rg);
      }
      mm_produce(a.tok_emb, a.n_vocab, d, d, c.ring, c.full, c.empty, c.rg);
    }
  } else {
    mm_consumer<NT, S>(c);
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------- host
typedef void (*MmKernel)(const SmallArgs);
static MmKernel mm_kernel(int nt, int s) {
  if (nt == 1) {
    switch (s) {
      case 3: return decode_small_mma_kernel<1, 3>;
      case 4: return decode_small_mma_kernel<1, 4>;
      case 5: return decode_small_mma_kernel<1, 5>;
    }
  } else if (nt == 2) {
    switch (s) {
      case 3: return decode_small_mma_kernel<2, 3>;
      case 4: return decode_small_mma_kernel<2, 4>;
      case 5: return decode_small_mma_kernel<2, 5>;
    }
  }
  return nullptr;
}

static int init_decode_small_mma() {
  static bool done = false;
  if (done) return kOk;
  for (int nt = 1; nt <= 2; ++nt)
    for (int s = 3; s <= 5; ++s)
      B200W_CUDA_OK(cudaFuncSetAttribute(mm_kernel(nt, s), cudaFuncAttributeMaxDynamicSharedMemorySize, kMmSmemBytes));
  done = true;
  return kOk;
}

bool decode_small_mma_applicable(const b200w_dims& dm, int n_seq, int n_q) {
  const int d = dm.n_text_state;
  return n_q == 1 && n_seq >= 1 && n_seq <= kMmMaxBatch && d % 256 == 0 && d / 256 >= 3 && d / 256 <= 5 && (d + 147) / 148 <= kMmOwnMax &&
         (3 * d + 147) / 148 <= kMmBiasMax && (4 * d + 147) / 148 <= kMmBiasMax && dm.n_audio_ctx <= kMmMaxKeys && dm.n_text_ctx <= kMmMaxKeys;
}

int launch_decode_small_mma(const SmallArgs& a, cudaStream_t stream) {
  B200W_CHECK_ARG(a.layers && a.tok_emb && a.x && a.q && a.att && a.qc && a.mlp && a.logits && a.counter, "decode_small_mma: null pointer");
  B200W_CHECK_ARG(a.page_size > 0 && (a.page_size & (a.page_size - 1)) == 0, "decode_small_mma: page_size must be a power of two");
  B200W_CHECK_ARG(a.max_pages <= kMmMaxPages, "decode_small_mma: more than %d pages per sequence", kMmMaxPages);
  B200W_TRY(init_decode_small_mma());
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(device_sm_count());
  cfg.blockDim = dim3(kMmThreads);
  cfg.dynamicSmemBytes = kMmSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;  // all CTAs co-resident: the grid barriers cannot deadlock
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  ProfScope prof_("decode_small_mma", stream);
  MmKernel kernel = mm_kernel((a.B + 7) / 8, a.d / 256);
  B200W_CHECK_ARG(kernel != nullptr, "decode_small_mma: unsupported batch %d / width %d", a.B, a.d);
  B200W_CUDA_OK(cudaLaunchKernelEx(&cfg, kernel, a));
  count_launch();
  return kOk;
}

}  // namespace b200w
