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
#ifndef B200W_CHAIN_STAGES
#define B200W_CHAIN_STAGES 8
#endif
constexpr int kChStages = B200W_CHAIN_STAGES;
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

__device__ __forceinline__ uint32_t chain_cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void chain_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA tile load delivered to the same shared-memory offset of every CTA in `mask`, completing on each one's barrier
__device__ __forceinline__ void tma_load_3d_mc(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                               uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4, %5}], "
      "[%2], %6;" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "h"(mask)
      : "memory");
}
// arrive on the barrier at this offset in every CTA of `mask` once this thread's prior MMAs have completed
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}

// All CTAs of the (cooperative, co-resident) grid meet here; `target` = CTAs x barriers passed so far.  `between` runs in
// thread 0 after the CTA's arrival and before the poll: work that does not depend on the barrier (the weight prefetch of the
// next GEMM phase) stays off every other CTA's critical path -- issued before the arrival it delayed the whole grid by the
// ~1.7 us it takes one thread to put 8 TMA loads in flight (tools/probe_chain_timeline.py).
template <typename F>
__device__ __forceinline__ void chain_grid_barrier(unsigned int* counter, unsigned int target, F&& between) {
  fence_proxy_async_global();  // this thread's generic-proxy global writes are ordered before later TMA reads
  __syncthreads();
  if (threadIdx.x == 0) {
    // release: the CTA's writes (ordered before this thread by the bar.sync above) are visible device-wide before the
    // arrival; acquire on the poll orders everything after it.  Explicit __threadfence() on both sides measured
    // 1.84 us per barrier (tools/probe_chain.py); the two fences are implied by .release / .acquire.
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");  // no return value to wait for
    between();
    unsigned int spins = 0;
    while (ld_acquire_gpu(counter) < target) {
      if (++spins > (1u << 26)) __trap();  // a lost CTA must not hang the GPU
    }
  }
  __syncthreads();
  fence_proxy_async_global();  // (measured free: 1.53 us per barrier with or without it)
}

// MC: clusters of kChCluster CTAs take the kChCluster neighbouring column tiles of one K range and share the A operand:
// every CTA loads a quarter of the 128 x 64 A tile and multicasts it to the cluster, so a CTA ingests 4 + 8 KB per K
// block instead of 16 + 8 KB (the <= 128-row GEMMs are bound by the per-SM operand ingest; the same activations used
// to be fetched by every CTA).  A stage is released to its four producers by the four MMA issuers (multicast commit).
#ifndef B200W_CHAIN_CLUSTER
#define B200W_CHAIN_CLUSTER 4
#endif
constexpr int kChCluster = B200W_CHAIN_CLUSTER;

template <bool MC>
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
  const int crank = MC ? (int)chain_cluster_rank() : 0;
  constexpr int kGroup = MC ? kChCluster : 1;            // CTAs that walk one work list together
  const int worker = blockIdx.x / kGroup, n_workers = gridDim.x / kGroup;
  constexpr uint16_t kMask = (1u << kChCluster) - 1;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < p.n_gemm; ++i) {
      tma_prefetch_desc(MC ? &maps.a4[i] : &maps.a[i]);
      tma_prefetch_desc(&maps.b[i]);
    }
    for (int i = 0; i < kChStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], kGroup);
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
  if constexpr (MC) chain_cluster_sync();  // every CTA's barriers exist before a peer's multicast can signal them
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
    const int q_tiles = (Q.tiles_n / kGroup) * p.tiles_m * Q.split_k;
    pre_stage0 = stage;
    pre_cnt = 0;
    for (int tile = worker; tile < q_tiles && pre_cnt < kChStages; tile += n_workers) {
      const int ks = tile % Q.split_k, nt = (tile / Q.split_k / p.tiles_m) * kGroup + crank;
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

  const int tl_cta = (p.timeline == nullptr) ? -1 : (blockIdx.x == 0 ? 0 : (blockIdx.x == gridDim.x / 2 ? 1 : -1));
  auto stamp = [&](int ph, int k) {
    if (tl_cta >= 0) p.timeline[(tl_cta * kChainMaxPhases + ph) * 8 + k] = clock64();
  };
  for (int ph = 0; ph < p.n_phases; ++ph) {
    const ChainPhase& P = p.ph[ph];
    if (tid == 0) stamp(ph, 0);  // phase start
    if (P.kind == kChainGemm) {
      // work items of this CTA's group: (column tile [group], row tile, K slice), K slice fastest
      const int num_tiles = (P.tiles_n / kGroup) * p.tiles_m * P.split_k;
      const CUtensorMap* ta = MC ? &maps.a4[P.map] : &maps.a[P.map];
      const CUtensorMap* tb = &maps.b[P.map];
      if (warp == 0) {
        if (lane == 0) {
          int item = 0;
          auto load_a = [&](int slot, int kb, int mt) {  // the whole A tile, or this CTA's quarter of it for the cluster
            unsigned char* sa = smem + slot * kChStageBytes;
            if constexpr (MC)
              tma_load_3d_mc(sa + crank * (kChABytes / kChCluster), ta, &full_bar[slot], kb * kChBK,
                             mt * kChBM + crank * (kChBM / kChCluster), 0, kMask);
            else
              tma_load_3d(sa, ta, &full_bar[slot], kb * kChBK, mt * kChBM, 0);
          };
          for (int tile = worker; tile < num_tiles; tile += n_workers) {
            const int ks = tile % P.split_k, rest = tile / P.split_k;
            const int mt = rest % p.tiles_m, nt = (rest / p.tiles_m) * kGroup + crank;
            const int kb_end = min(P.num_kb, (ks + 1) * P.kb_per_split);
            for (int kb = ks * P.kb_per_split; kb < kb_end; ++kb, ++item) {
              if (item < pre_cnt) {  // W tile already in flight (issued before the grid barrier): only A is missing
                load_a((pre_stage0 + item) % kChStages, kb, mt);
                continue;
              }
              mbar_wait(&empty_bar[stage], ring_phase ^ 1);
              unsigned char* sa = smem + stage * kChStageBytes;
              mbar_expect_tx(&full_bar[stage], kChStageBytes);
              load_a(stage, kb, mt);
              tma_load_3d(sa + kChABytes, tb, &full_bar[stage], kb * kChBK, nt * kChBN, 0);
              if (++stage == kChStages) {
                stage = 0;
                ring_phase ^= 1;
              }
            }
          }
          stamp(ph, 1);  // this phase's loads are issued
          pre_cnt = 0;
          if (ph + 1 < p.n_phases && p.ph[ph + 1].kind == kChainGemm) prefetch_w(p.ph[ph + 1]);
        }
        __syncwarp();
      } else if (warp == 1) {
        if (lane == 0) {
          constexpr uint32_t idesc = make_idesc_bf16(kChBM, kChBN, 0, 0);
          for (int tile = worker; tile < num_tiles; tile += n_workers) {
            mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
            tcgen05_fence_after();
            const uint32_t d_tmem = tmem_base + acc * kChBN;
            const int kb_begin = (tile % P.split_k) * P.kb_per_split;
            const int kb_end = min(P.num_kb, kb_begin + P.kb_per_split);
            for (int kb = kb_begin; kb < kb_end; ++kb) {
              mbar_wait(&full_bar[stage], ring_phase);
              tcgen05_fence_after();
              if (kb == kb_begin && tile == worker) stamp(ph, 2);  // first operands landed
              const uint32_t sa = smem_u32(smem + stage * kChStageBytes);
              const uint64_t a_desc = make_sw128_desc(sa);
              const uint64_t b_desc = make_sw128_desc(sa + kChABytes);
#pragma unroll
              for (int k = 0; k < kChBK / 16; ++k)
                umma_f16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb > kb_begin || k != 0) ? 1u : 0u);
              if constexpr (MC) umma_commit_mc(&empty_bar[stage], kMask);  // the slot is refilled by all four producers
              else umma_commit(&empty_bar[stage]);
              if (++stage == kChStages) {
                stage = 0;
                ring_phase ^= 1;
              }
            }
            umma_commit(&tmem_full_bar[acc]);
            if (tile + n_workers >= num_tiles) stamp(ph, 3);  // last MMA issued
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
        for (int tile = worker; tile < num_tiles; tile += n_workers) {
          const int ks = tile % P.split_k, rest = tile / P.split_k;
          const int mt = rest % p.tiles_m, nt = (rest / p.tiles_m) * kGroup + crank;
          const int t0 = mt * kChBM + q * 32;
          const int rows_here = min(32, p.rows - t0);
          mbar_wait(&tmem_full_bar[acc], acc_phase);
          tcgen05_fence_after();
          if (warp == 4 && lane == 0 && tile + n_workers >= num_tiles) stamp(ph, 4);  // last accumulator complete
          const uint32_t t_base = tmem_base + ((uint32_t)(q * 32) << 16) + acc * kChBN;
          const int col = nt * kChBN + hh * 32;
          uint32_t r[32];
          tmem_ld_32x32(t_base + hh * 32, r);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; ++j) stg[lane * 33 + j] = __uint_as_float(r[j]);
          __syncwarp();
          if (!P.gelu) {
            // 16-byte stores: a lane owns four columns of rows rq, rq + 4, ... (conflict-free reads of the padded tile)
            const int sub = lane & 7, rq = lane >> 3;
            float* op = reinterpret_cast<float*>(P.out) + ks * P.split_stride + (long long)t0 * P.ldc + col + 4 * sub;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int rr = 4 * i + rq;
              const float* sp = stg + rr * 33 + 4 * sub;
              if (rr < rows_here) *reinterpret_cast<float4*>(op + (long long)rr * P.ldc) = make_float4(sp[0], sp[1], sp[2], sp[3]);
            }
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
      // residual + LayerNorm, a row per CTA and round (rows <= 2 x grid): thread t < d / 4 owns four consecutive features
      for (int row = blockIdx.x; row < p.rows; row += gridDim.x) {
        __syncthreads();  // s_red is reused by the next row
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
    if (warp == 4 && lane == 0) stamp(ph, 5);  // epilogue warp 4 done with the phase
    if (ph + 1 < p.n_phases) {
      if (tid == 0) stamp(ph, 6);  // arrives at the grid barrier (before the CTA-wide sync)
      // after a LayerNorm phase the producer thread puts the next GEMM's weight tiles in flight while it waits (a GEMM
      // phase's producer has already done this at the end of its loop, under the phase's own MMAs)
      const bool pre = P.kind == kChainLn && p.ph[ph + 1].kind == kChainGemm;
      chain_grid_barrier(p.counter, (unsigned int)(ph + 1) * gridDim.x, [&]() {
        if (pre) prefetch_w(p.ph[ph + 1]);
      });
      if (tid == 0) stamp(ph, 7);  // released
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if constexpr (MC) chain_cluster_sync();  // no CTA leaves while a peer may still signal its barriers
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, kChTmemCols);
  }
}

static long long* g_chain_timeline = nullptr;
static int g_chain_timeline_countdown = -1;
void set_chain_timeline(long long* dev, int launch_index) {
  g_chain_timeline = dev;
  g_chain_timeline_countdown = dev ? launch_index : -1;
}

static int g_chain_mc_grid = 0;  // CTAs of the multicast form that can be co-resident (0: unavailable)

int init_chain() {
  static bool done = false;
  if (done) return kOk;
  B200W_CUDA_OK(cudaFuncSetAttribute(decode_chain_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kChSmemBytes));
  B200W_CUDA_OK(cudaFuncSetAttribute(decode_chain_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kChSmemBytes));
  // the largest shared-memory carve-out whatever kChStages is: kernels that are to share an SM need the SAME carve-out
  cudaFuncSetAttribute(decode_chain_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  cudaFuncSetAttribute(decode_chain_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  // B200W_CHAIN_MC=1 opts into the cluster-multicast form.  Measured on B200 (large-v3, 120 sequences): 7.46 ms per step
  // against 7.44 ms without it -- only 33 clusters of 4 fit at once (132 of 148 SMs), the four CTAs of a cluster advance
  // in lockstep, and a phase is dominated by its fixed ~4.5 us, not by operand ingest.  Off by default.
  const char* e = getenv("B200W_CHAIN_MC");
  if (e != nullptr && e[0] == '1') {
    // clusters of 4 must sit inside one GPC: ask how many fit at once, the cooperative grid cannot be larger
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(device_sm_count() / kChCluster * kChCluster);
    cfg.blockDim = dim3(kChThreads);
    cfg.dynamicSmemBytes = kChSmemBytes;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kChCluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n_clusters = 0;
    if (cudaOccupancyMaxActiveClusters(&n_clusters, decode_chain_kernel<true>, &cfg) == cudaSuccess && n_clusters > 0) {
      const int cap = device_sm_count() / kChCluster;
      g_chain_mc_grid = (n_clusters < cap ? n_clusters : cap) * kChCluster;
    } else {
      (void)cudaGetLastError();
    }
  }
  done = true;
  return kOk;
}

// ---------------------------------------------------------------------------------------------- host
int chain_mc_cluster() { return kChCluster; }

int chain_mc_grid() {
  if (init_chain() != kOk) return -1;
  return g_chain_mc_grid;
}

int chain_add_gemm(ChainMaps* maps, ChainParams* p, const void* A, long long lda, const void* W, int N, int K, int split_k,
                   void* out, long long ldc, long long split_stride, const float* bias, bool gelu) {
  B200W_CHECK_ARG(p->n_phases < kChainMaxPhases && p->n_gemm < kChainMaxGemm, "chain: too many phases");
  B200W_CHECK_ARG(N % kChBN == 0 && K % kChBK == 0 && p->rows > 0 && p->rows <= 2 * kChBM, "chain: unsupported GEMM shape");
  p->tiles_m = ceil_div(p->rows, kChBM);
  B200W_CHECK_ARG(!gelu || (split_k <= 1 && bias != nullptr), "chain: GELU phase cannot be split");
  B200W_CHECK_ARG(gelu || ((ldc & 3) == 0 && (split_stride & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0),
                  "chain: partial slabs must be 16-byte aligned");
  ChainPhase& P = p->ph[p->n_phases++];
  P = ChainPhase{};
  P.kind = kChainGemm;
  P.map = p->n_gemm++;
  B200W_TRY(make_tmap_a(&maps->a[P.map], A, 1, p->rows, K, lda, (long long)p->rows * lda));
  {  // the same operand in quarter-tile boxes (32 rows) for the multicast form
    uint64_t dims[3] = {(uint64_t)K, (uint64_t)p->rows, 1};
    uint64_t strides[2] = {(uint64_t)lda * 2, (uint64_t)p->rows * lda * 2};
    uint32_t box[3] = {kChBK, kChBM / kChCluster, 1};
    B200W_TRY(encode_tmap_bf16(&maps->a4[P.map], A, 3, dims, strides, box));
  }
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
  B200W_TRY(init_chain());
  // the multicast form needs every GEMM phase to have a multiple of 4 column tiles and a grid that holds the rows
  bool mc = g_chain_mc_grid > 0;
  for (int i = 0; i < p.n_phases; ++i)
    if (p.ph[i].kind == kChainGemm && p.ph[i].tiles_n % kChCluster != 0) mc = false;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(mc ? g_chain_mc_grid : device_sm_count());
  cfg.blockDim = dim3(kChThreads);
  cfg.dynamicSmemBytes = kChSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeCooperative;  // all CTAs co-resident: the grid barriers cannot deadlock
  attr[0].val.cooperative = 1;
  attr[1].id = cudaLaunchAttributeClusterDimension;
  attr[1].val.clusterDim.x = kChCluster;
  attr[1].val.clusterDim.y = 1;
  attr[1].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = mc ? 2 : 1;
  // B200W_CHAIN_NONCOOP=1 (experiment): a plain launch -- co-residency of the 148 CTAs is then only what the sizes of the
  // kernels sharing the GPU make it (one CTA per SM by shared memory); a barrier that cannot complete traps
  static const bool noncoop = [] { const char* e = getenv("B200W_CHAIN_NONCOOP"); return e != nullptr && e[0] == '1'; }();
  if (noncoop && !mc) cfg.numAttrs = 0;
  ProfScope prof_("dec_chain", stream);
  ChainParams pl = p;
  pl.timeline = nullptr;
  if (g_chain_timeline_countdown >= 0 && g_chain_timeline_countdown-- == 0) pl.timeline = g_chain_timeline;
  if (mc)
    B200W_CUDA_OK(cudaLaunchKernelEx(&cfg, decode_chain_kernel<true>, maps, pl));
  else
    B200W_CUDA_OK(cudaLaunchKernelEx(&cfg, decode_chain_kernel<false>, maps, pl));
  count_launch();
  return kOk;
}

}  // namespace b200w
