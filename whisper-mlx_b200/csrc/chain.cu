// K11: decode chain -- several small-M phases of one decoder layer in ONE persistent launch.
//
// A greedy decode step runs 11 launches per decoder layer around the two attention kernels, each a <= 128-row GEMM
// or a 120-row residual + LayerNorm whose duration is launch ramp and dependent-latency chain, not work
// (profiles/r01_decode_launches_v2.md: ~86 us per layer outside the cross-attention for ~7 us of weight streaming).
// This kernel strings such phases together behind grid barriers: one cooperative launch of one CTA per SM keeps the
// mbarrier ring, the TMEM accumulators and the warp roles of the K5 GEMM alive across phases
//   GEMM phase : A (rows <= 128, K) bf16 by TMA  x  W (N, K)^T  ->  raw fp32 split-K slabs, or bias + GELU -> bf16
//   LN phase   : x += bias + sum of slabs;  h = LayerNorm(x) as bf16        (one row per CTA, K4's arithmetic)
// so that e.g. [cross-attn out-proj -> LN -> MLP1 -> MLP2 -> LN -> QKV] of mlx_whisper's ResidualAttentionBlock
// (reference call site /root/reference/run:3-6) is one launch instead of six.
// Phase results cross CTAs through L2: producers fence (generic and async proxy) before the barrier, the LayerNorm
// reads slabs with ld.global.cg, the GEMM operands arrive by TMA.
#include "common.cuh"
#include "kernels.h"

namespace b200w {

constexpr int kChBM = 128;
constexpr int kChBN = 64;
constexpr int kChBK = 64;
constexpr int kChThreads = 384;  // 4 role warps + 8 epilogue warps (the K5 layout)
constexpr int kChStages = 8;
constexpr int kChABytes = kChBM * kChBK * 2;
constexpr int kChBBytes = kChBN * kChBK * 2;
constexpr int kChStageBytes = kChABytes + kChBBytes;
constexpr int kChTmemCols = 2 * kChBN;
constexpr int kChEpiBytes = 8 * 32 * 33 * 4;
constexpr int kChSmemBytes = kChStages * kChStageBytes + 1024 + 256 + kChEpiBytes;

__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// All CTAs of the (cooperative, co-resident) grid meet here; `target` = CTAs x barriers passed so far.
__device__ __forceinline__ void chain_grid_barrier(unsigned int* counter, unsigned int target) {
  fence_proxy_async_global();  // this thread's generic-proxy global writes are ordered before later TMA reads
  __syncthreads();
  if (threadIdx.x == 0) {
    // release: the CTA's writes (ordered before this thread by the bar.sync above) are visible device-wide before the
    // arrival; acquire on the poll orders everything after it.  Explicit __threadfence() on both sides measured
    // 1.84 us per barrier (tools/probe_chain.py); the two fences are implied by .release / .acquire.
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");  // no return value to wait for
    unsigned int spins = 0;
    while (ld_acquire_gpu(counter) < target) {
      if (++spins > (1u << 26)) __trap();  // a lost CTA must not hang the GPU
    }
  }
  __syncthreads();
  fence_proxy_async_global();  // (measured free: 1.53 us per barrier with or without it)
}

__global__ void __launch_bounds__(kChThreads, 1)
decode_chain_kernel(const __grid_constant__ ChainMaps maps, const __grid_constant__ ChainParams p) {
  extern __shared__ unsigned char chain_smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(chain_smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kChStages * kChStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kChStages;
  uint64_t* tmem_full_bar = bars + 2 * kChStages;
  uint64_t* tmem_empty_bar = bars + 2 * kChStages + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kChStages + 4);
  __shared__ float s_red[2][kChThreads / 32];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < p.n_gemm; ++i) {
      tma_prefetch_desc(&maps.a[i]);
      tma_prefetch_desc(&maps.b[i]);
    }
    for (int i = 0; i < kChStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], 8);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, kChTmemCols);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // pipeline state, carried across phases: each role thread walks the same tile list in every GEMM phase, so the
  // producer's / issuer's ring position and the issuer's / epilogue's accumulator parity advance in step
  int stage = 0;
  uint32_t ring_phase = 0;
  int acc = 0;
  uint32_t acc_phase = 0;

  // Producer only: the weight tiles of a GEMM phase do not depend on the phase before it, so up to a ring of them is
  // put in flight BEFORE the grid barrier that releases its A operand; the A tiles follow after the barrier.
  int pre_cnt = 0, pre_stage0 = 0;
  auto prefetch_w = [&](const ChainPhase& Q) {
    const CUtensorMap* qb = &maps.b[Q.map];
    const int q_tiles = Q.tiles_n * Q.split_k;
    pre_stage0 = stage;
    pre_cnt = 0;
    for (int tile = blockIdx.x; tile < q_tiles && pre_cnt < kChStages; tile += gridDim.x) {
      const int ks = tile % Q.split_k, nt = tile / Q.split_k;
      const int kb_end = min(Q.num_kb, (ks + 1) * Q.kb_per_split);
      for (int kb = ks * Q.kb_per_split; kb < kb_end && pre_cnt < kChStages; ++kb, ++pre_cnt) {
        mbar_wait(&empty_bar[stage], ring_phase ^ 1);
        mbar_expect_tx(&full_bar[stage], kChStageBytes);
        tma_load_3d(smem + stage * kChStageBytes + kChABytes, qb, &full_bar[stage], kb * kChBK, nt * kChBN, 0);
        if (++stage == kChStages) {
          stage = 0;
          ring_phase ^= 1;
        }
      }
    }
  };

  for (int ph = 0; ph < p.n_phases; ++ph) {
    const ChainPhase& P = p.ph[ph];
    if (P.kind == kChainGemm) {
      const int num_tiles = P.tiles_n * P.split_k;
      const CUtensorMap* ta = &maps.a[P.map];
      const CUtensorMap* tb = &maps.b[P.map];
      if (warp == 0) {
        if (lane == 0) {
          int item = 0;
          for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int ks = tile % P.split_k, nt = tile / P.split_k;
            const int kb_end = min(P.num_kb, (ks + 1) * P.kb_per_split);
            for (int kb = ks * P.kb_per_split; kb < kb_end; ++kb, ++item) {
              if (item < pre_cnt) {  // W tile already in flight (issued before the grid barrier): only A is missing
                const int slot = (pre_stage0 + item) % kChStages;
                tma_load_3d(smem + slot * kChStageBytes, ta, &full_bar[slot], kb * kChBK, 0, 0);
                continue;
              }
              mbar_wait(&empty_bar[stage], ring_phase ^ 1);
              unsigned char* sa = smem + stage * kChStageBytes;
              mbar_expect_tx(&full_bar[stage], kChStageBytes);
              tma_load_3d(sa, ta, &full_bar[stage], kb * kChBK, 0, 0);
              tma_load_3d(sa + kChABytes, tb, &full_bar[stage], kb * kChBK, nt * kChBN, 0);
              if (++stage == kChStages) {
                stage = 0;
                ring_phase ^= 1;
              }
            }
          }
          pre_cnt = 0;
          if (ph + 1 < p.n_phases && p.ph[ph + 1].kind == kChainGemm) prefetch_w(p.ph[ph + 1]);
        }
        __syncwarp();
      } else if (warp == 1) {
        if (lane == 0) {
          constexpr uint32_t idesc = make_idesc_bf16(kChBM, kChBN, 0, 0);
          for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
            tcgen05_fence_after();
            const uint32_t d_tmem = tmem_base + acc * kChBN;
            const int kb_begin = (tile % P.split_k) * P.kb_per_split;
            const int kb_end = min(P.num_kb, kb_begin + P.kb_per_split);
            for (int kb = kb_begin; kb < kb_end; ++kb) {
              mbar_wait(&full_bar[stage], ring_phase);
              tcgen05_fence_after();
              const uint32_t sa = smem_u32(smem + stage * kChStageBytes);
              const uint64_t a_desc = make_sw128_desc(sa);
              const uint64_t b_desc = make_sw128_desc(sa + kChABytes);
#pragma unroll
              for (int k = 0; k < kChBK / 16; ++k)
                umma_f16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb > kb_begin || k != 0) ? 1u : 0u);
              umma_commit(&empty_bar[stage]);
              if (++stage == kChStages) {
                stage = 0;
                ring_phase ^= 1;
              }
            }
            umma_commit(&tmem_full_bar[acc]);
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
          }
        }
        __syncwarp();
      } else if (warp >= 4) {
        // the K5 epilogue: TMEM lane quarter e % 4, 32-column chunks e / 4 and e / 4 + 2... (BN = 64: one chunk per warp)
        const int e = warp - 4;
        const int q = e & 3, hh = e >> 2;
        float* stg = reinterpret_cast<float*>(smem + kChStages * kChStageBytes + 256) + e * (32 * 33);
        const int t0 = q * 32;
        const int rows_here = min(32, p.rows - t0);
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
          const int ks = tile % P.split_k, nt = tile / P.split_k;
          mbar_wait(&tmem_full_bar[acc], acc_phase);
          tcgen05_fence_after();
          const uint32_t t_base = tmem_base + ((uint32_t)(q * 32) << 16) + acc * kChBN;
          const int col = nt * kChBN + hh * 32;
          uint32_t r[32];
          tmem_ld_32x32(t_base + hh * 32, r);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; ++j) stg[lane * 33 + j] = __uint_as_float(r[j]);
          __syncwarp();
          if (!P.gelu) {
            float* op = reinterpret_cast<float*>(P.out) + ks * P.split_stride + (long long)t0 * P.ldc + col + lane;
#pragma unroll
            for (int rr = 0; rr < 32; ++rr)
              if (rr < rows_here) op[(long long)rr * P.ldc] = stg[rr * 33 + lane];
          } else {
            const int l2 = (lane & 15) * 2, hi = lane >> 4;
            const float b0 = __ldg(P.bias + col + l2), b1 = __ldg(P.bias + col + l2 + 1);
            __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(P.out) + (long long)t0 * P.ldc + col + l2;
#pragma unroll 4
            for (int rr = hi; rr < rows_here; rr += 2) {
              const float v0 = gelu_fast(stg[rr * 33 + l2] + b0), v1 = gelu_fast(stg[rr * 33 + l2 + 1] + b1);
              *reinterpret_cast<uint32_t*>(op + (long long)rr * P.ldc) = pack_bf16x2(v0, v1);
            }
          }
          __syncwarp();
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1;
        }
      }
    } else {
      // residual + LayerNorm, one row per CTA (rows <= gridDim.x): thread t < d / 4 owns four consecutive features
      const int row = blockIdx.x;
      if (row < p.rows) {
        const int d = P.d;
        const bool on = tid < d / 4;
        const long long off = (long long)row * d + tid * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (on) {
          v = *reinterpret_cast<const float4*>(P.x + off);
          if (P.n_split > 0) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(P.bias) + tid);
            float4 s[8];
#pragma unroll
            for (int i = 0; i < 8; ++i)
              s[i] = (i < P.n_split) ? __ldcg(reinterpret_cast<const float4*>(P.part + i * P.split_stride + off))
                                     : make_float4(0.f, 0.f, 0.f, 0.f);
            v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              v.x += s[i].x; v.y += s[i].y; v.z += s[i].z; v.w += s[i].w;
            }
            *reinterpret_cast<float4*>(P.x + off) = v;
          }
        }
        float sum = warp_sum((v.x + v.y) + (v.z + v.w));
        if (lane == 0) s_red[0][warp] = sum;
        __syncthreads();
        sum = 0.0f;
#pragma unroll
        for (int i = 0; i < kChThreads / 32; ++i) sum += s_red[0][i];
        const float mean = sum / (float)d;
        const float a = on ? v.x - mean : 0.f, b2 = on ? v.y - mean : 0.f, c = on ? v.z - mean : 0.f, e2 = on ? v.w - mean : 0.f;
        float var = warp_sum((a * a + b2 * b2) + (c * c + e2 * e2));
        if (lane == 0) s_red[1][warp] = var;
        __syncthreads();
        var = 0.0f;
#pragma unroll
        for (int i = 0; i < kChThreads / 32; ++i) var += s_red[1][i];
        const float rstd = rsqrtf(var / (float)d + 1e-5f);
        if (on) {
          const float4 g = __ldg(reinterpret_cast<const float4*>(P.gamma) + tid), bb = __ldg(reinterpret_cast<const float4*>(P.beta) + tid);
          *reinterpret_cast<uint2*>(P.h + off) = make_uint2(pack_bf16x2(a * rstd * g.x + bb.x, b2 * rstd * g.y + bb.y),
                                                           pack_bf16x2(c * rstd * g.z + bb.z, e2 * rstd * g.w + bb.w));
        }
      }
    }
    if (P.kind == kChainLn) {  // (a GEMM phase's producer has already done this at the end of its loop)
      if (tid == 0 && ph + 1 < p.n_phases && p.ph[ph + 1].kind == kChainGemm) prefetch_w(p.ph[ph + 1]);
      __syncwarp();
    }
    if (ph + 1 < p.n_phases) chain_grid_barrier(p.counter, (unsigned int)(ph + 1) * gridDim.x);
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, kChTmemCols);
  }
}

int init_chain() {
  static bool done = false;
  if (done) return kOk;
  B200W_CUDA_OK(cudaFuncSetAttribute(decode_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kChSmemBytes));
  done = true;
  return kOk;
}

// ---------------------------------------------------------------------------------------------- host
int chain_add_gemm(ChainMaps* maps, ChainParams* p, const void* A, long long lda, const void* W, int N, int K, int split_k,
                   void* out, long long ldc, long long split_stride, const float* bias, bool gelu) {
  B200W_CHECK_ARG(p->n_phases < kChainMaxPhases && p->n_gemm < kChainMaxGemm, "chain: too many phases");
  B200W_CHECK_ARG(N % kChBN == 0 && K % kChBK == 0 && p->rows > 0 && p->rows <= kChBM, "chain: unsupported GEMM shape");
  B200W_CHECK_ARG(!gelu || (split_k <= 1 && bias != nullptr), "chain: GELU phase cannot be split");
  ChainPhase& P = p->ph[p->n_phases++];
  P = ChainPhase{};
  P.kind = kChainGemm;
  P.map = p->n_gemm++;
  B200W_TRY(make_tmap_a(&maps->a[P.map], A, 1, p->rows, K, lda, (long long)p->rows * lda));
  B200W_TRY(make_tmap_w(&maps->b[P.map], W, N, K, kChBN));
  P.tiles_n = N / kChBN;
  P.num_kb = K / kChBK;
  if (split_k < 1) split_k = 1;
  P.kb_per_split = ceil_div(P.num_kb, split_k);
  P.split_k = ceil_div(P.num_kb, P.kb_per_split);
  P.gelu = gelu ? 1 : 0;
  P.bias = bias;
  P.out = out;
  P.ldc = ldc;
  P.split_stride = split_stride;
  return kOk;
}

int chain_add_ln(ChainParams* p, float* x, const float* part, int n_split, long long split_stride, const float* bias,
                 const float* gamma, const float* beta, int d, __nv_bfloat16* h) {
  B200W_CHECK_ARG(p->n_phases < kChainMaxPhases, "chain: too many phases");
  B200W_CHECK_ARG(d % 4 == 0 && d / 4 <= kChThreads && n_split >= 0 && n_split <= 8, "chain: unsupported LayerNorm shape");
  B200W_CHECK_ARG(n_split == 0 || (part && bias), "chain: partial slabs without bias");
  ChainPhase& P = p->ph[p->n_phases++];
  P = ChainPhase{};
  P.kind = kChainLn;
  P.x = x;
  P.part = part;
  P.n_split = n_split;
  P.split_stride = split_stride;
  P.bias = bias;
  P.gamma = gamma;
  P.beta = beta;
  P.d = d;
  P.h = h;
  return kOk;
}

int launch_chain(const ChainMaps& maps, const ChainParams& p, cudaStream_t stream) {
  B200W_CHECK_ARG(p.n_phases > 0 && p.counter != nullptr, "chain: empty chain or no barrier counter");
  B200W_CHECK_ARG(p.rows <= device_sm_count(), "chain: one LayerNorm row per CTA needs rows <= SM count");
  B200W_TRY(init_chain());
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(device_sm_count());
  cfg.blockDim = dim3(kChThreads);
  cfg.dynamicSmemBytes = kChSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;  // all CTAs co-resident: the grid barriers cannot deadlock
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  ProfScope prof_("dec_chain", stream);
  B200W_CUDA_OK(cudaLaunchKernelEx(&cfg, decode_chain_kernel, maps, p));
  count_launch();
  return kOk;
}

}  // namespace b200w
