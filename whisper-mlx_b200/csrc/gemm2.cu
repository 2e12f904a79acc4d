// K5 (2-CTA form): bf16 GEMM with tcgen05.mma.cta_group::2 -- two SMs of a cluster work on one 256 x 256 tile.
//
// The single-CTA kernel (gemm.cu) moves 48 KB of operands into shared memory per 128x256x64 MMA block and is bound by
// the per-SM operand ingest (~67 B/clk measured), not by the tensor pipe.  Here each CTA of a pair supplies its own
// 128 rows of A and HALF of the W tile (128 of the 256 columns); the pair's tensor cores read the other half from the
// peer's shared memory, so every CTA ingests 32 KB per K block for the same amount of math.  Accumulators: rows
// 0..127 of the tile in the leader's TMEM, rows 128..255 in the peer's, double buffered (2 x 256 columns each).
//
// Protocol (barriers have the same shared-memory offsets in both CTAs):
//   full[s]       leader only.  Both producers' TMA loads complete_tx on the LEADER's barrier (peer bit masked off the
//                 mbarrier address, cp.async.bulk.tensor ... cta_group::2); the leader arms it with the pair's bytes.
//   empty[s]      one per CTA; the leader's tcgen05.commit multicasts the arrive to both CTAs.
//   tmem_full[a]  one per CTA, multicast commit after the last K block of a tile.
//   tmem_empty[a] leader only, 16 arrivals: the 8 epilogue warps of each CTA (the peer's arrive remotely).
// Used for plain (single-slab) GEMMs with N % 256 == 0 and many row tiles: the encoder projections and the cross-K/V.
#include "common.cuh"
#include "kernels.h"

namespace b200w {

constexpr int k2BM = 128;           // rows per CTA (256 per pair)
constexpr int k2BN = 256;           // columns per pair tile; each CTA stages 128 of them
constexpr int k2BK = 64;
constexpr int k2Stages = 5;
constexpr int k2Threads = 384;
constexpr int k2ABytes = k2BM * k2BK * 2;          // 16 KB
constexpr int k2BBytes = (k2BN / 2) * k2BK * 2;    // 16 KB
constexpr int k2StageBytes = k2ABytes + k2BBytes;
constexpr int k2EpiBytes = 8 * 32 * 33 * 4;
constexpr int k2SmemBytes = k2Stages * k2StageBytes + 1024 + 256 + k2EpiBytes;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// TMA load whose completion bytes are credited to the pair leader's mbarrier
__device__ __forceinline__ void tma_load_3d_2sm(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  const uint32_t leader_bar = smem_u32(bar) & 0xFEFFFFFFu;
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at this shared-memory offset in BOTH CTAs once all prior MMAs of the pair have completed
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"((unsigned short)3)
               : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(k2Threads, 1)
gemm2_bf16_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, const GemmParams p) {
  extern __shared__ unsigned char gemm2_smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(gemm2_smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + k2Stages * k2StageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + k2Stages;
  uint64_t* tmem_full_bar = bars + 2 * k2Stages;
  uint64_t* tmem_empty_bar = bars + 2 * k2Stages + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * k2Stages + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int n_clusters = gridDim.x >> 1, cluster_id = blockIdx.x >> 1;

  const int tiles_m = p.tiles_m_per_batch;  // 256-row pair tiles
  const int num_tiles = tiles_m * p.tiles_n;
  const int num_kb = (p.K + k2BK - 1) / k2BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    for (int i = 0; i < k2Stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], 16);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc2(tmem_slot, 512);
    tmem_relinquish2();
  }
  tcgen05_fence_before();
  cluster_sync_all();  // barrier inits and TMEM allocation of both CTAs are visible to the pair
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto decode_tile = [&](int id, int& mt, int& nt) {
    const int per_group = p.group_m * p.tiles_n;
    const int g = id / per_group;
    const int first_m = g * p.group_m;
    const int gsize = min(p.group_m, tiles_m - first_m);
    const int within = id - g * per_group;
    nt = within / gsize;
    mt = first_m + (within - nt * gsize);
  };

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += n_clusters) {
        int mt, nt;
        decode_tile(tile, mt, nt);
        const int m0 = mt * 2 * k2BM + (int)rank * k2BM;    // this CTA's 128 rows of A
        const int n0 = nt * k2BN + (int)rank * (k2BN / 2);  // this CTA's half of the W tile
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          unsigned char* sa = smem + stage * k2StageBytes;
          unsigned char* sb = sa + k2ABytes;
          if (leader) mbar_expect_tx(&full_bar[stage], 2 * k2StageBytes);
          tma_load_3d_2sm(sa, &tma_a, &full_bar[stage], kb * k2BK, m0, 0);
          tma_load_3d_2sm(sb, &tma_b, &full_bar[stage], kb * k2BK, n0, 0);
          if (++stage == k2Stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (leader && lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(256, k2BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += n_clusters) {
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * k2BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(smem + stage * k2StageBytes);
          const uint64_t a_desc = make_sw128_desc(sa);
          const uint64_t b_desc = make_sw128_desc(sa + k2ABytes);
#pragma unroll
          for (int k = 0; k < k2BK / 16; ++k)
            umma2_f16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          umma2_commit_mc(&empty_bar[stage]);
          if (++stage == k2Stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma2_commit_mc(&tmem_full_bar[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // same epilogue as gemm.cu: TMEM -> padded smem transpose -> coalesced bias / GELU / residual / store
    const int e = warp - 4;
    const int q = e & 3, hh = e >> 2;
    float* stage_buf = reinterpret_cast<float*>(smem + k2Stages * k2StageBytes + 256) + e * (32 * 33);
    const uint32_t leader_empty0 = mapa_shared(smem_u32(&tmem_empty_bar[0]), 0);
    int acc = 0;
    uint32_t acc_phase = 0;
    const bool vec_bf16 = !p.out_f32 && (p.ldc & 7) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0;
    for (int tile = cluster_id; tile < num_tiles; tile += n_clusters) {
      int mt, nt;
      decode_tile(tile, mt, nt);
      const int t0 = mt * 2 * k2BM + (int)rank * k2BM + q * 32;  // first row of this warp
      const int n0 = nt * k2BN;
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tcgen05_fence_after();
      const uint32_t t_base = tmem_base + ((uint32_t)(q * 32) << 16) + acc * k2BN;
      const int rows_here = min(32, p.rows_per_batch - t0);
#pragma unroll 1
      for (int c = hh; c < k2BN / 32; c += 2) {
        const int col = n0 + c * 32;
        float rsd[32];
        if (p.out_f32 && p.resid != nullptr) {
          const float* rp = p.resid + (long long)t0 * p.resid_ld + col + lane;
#pragma unroll
          for (int rr = 0; rr < 32; ++rr) rsd[rr] = (rr < rows_here) ? rp[(long long)rr * p.resid_ld] : 0.0f;
        } else {
#pragma unroll
          for (int rr = 0; rr < 32; ++rr) rsd[rr] = 0.0f;
        }
        uint32_t r[32];
        tmem_ld_32x32(t_base + c * 32, r);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; ++j) stage_buf[lane * 33 + j] = __uint_as_float(r[j]);
        __syncwarp();
        if (p.out_f32) {
          const float bias_l = (p.bias != nullptr) ? __ldg(p.bias + col + lane) : 0.0f;
          float* op = reinterpret_cast<float*>(p.out) + (long long)t0 * p.ldc + col + lane;
#pragma unroll
          for (int rr = 0; rr < 32; ++rr) {
            float v = stage_buf[rr * 33 + lane] + bias_l;
            if (p.gelu) v = gelu_fast(v);
            v += rsd[rr];
            if (rr < rows_here) op[(long long)rr * p.ldc] = v;
          }
        } else if (vec_bf16) {
          // 16-byte stores: a lane owns eight consecutive columns of rows r8, r8 + 8, ... (as in gemm.cu)
          const int s4 = lane & 3, r8 = lane >> 2;
          float bs[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) bs[j] = (p.bias != nullptr) ? __ldg(p.bias + col + 8 * s4 + j) : 0.0f;
          __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + (long long)t0 * p.ldc + col + 8 * s4;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int rr = 8 * i + r8;
            const float* sp = stage_buf + rr * 33 + 8 * s4;
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              v[j] = sp[j] + bs[j];
              if (p.gelu) v[j] = gelu_fast(v[j]);
            }
            if (rr < rows_here)
              *reinterpret_cast<uint4*>(op + (long long)rr * p.ldc) =
                  make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
          }
        } else {
          const int l2 = (lane & 15) * 2, hi = lane >> 4;
          float b0 = 0.0f, b1 = 0.0f;
          if (p.bias != nullptr) {
            b0 = __ldg(p.bias + col + l2);
            b1 = __ldg(p.bias + col + l2 + 1);
          }
          __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + (long long)t0 * p.ldc + col + l2;
#pragma unroll 4
          for (int rr = hi; rr < rows_here; rr += 2) {
            float v0 = stage_buf[rr * 33 + l2] + b0, v1 = stage_buf[rr * 33 + l2 + 1] + b1;
            if (p.gelu) {
              v0 = gelu_fast(v0);
              v1 = gelu_fast(v1);
            }
            *reinterpret_cast<uint32_t*>(op + (long long)rr * p.ldc) = pack_bf16x2(v0, v1);
          }
        }
        __syncwarp();
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(leader_empty0 + acc * 8);  // the leader's own warps use the same cluster address
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tcgen05_fence_before();
  cluster_sync_all();  // nobody tears down TMEM / exits while the peer may still read this CTA's shared memory
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc2(tmem_base, 512);
  }
}

int init_gemm2() {
  static bool done = false;
  if (done) return kOk;
  B200W_CUDA_OK(cudaFuncSetAttribute(gemm2_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, k2SmemBytes));
  done = true;
  return kOk;
}

// Measured on B200 (tools/probe_gemm2.py): +4-5 % on the bf16-output projections (QKV, MLP1, cross-K/V); the
// f32-residual GEMMs are epilogue-bound and run slightly slower in pairs, so they stay on the single-CTA kernel.
bool gemm2_applicable(const GemmParams& p) {
  return p.n_batch == 1 && p.resid == nullptr && p.split_k <= 1 && p.n_store % k2BN == 0 && p.N == p.n_store && p.resid_mod == 0 &&
         (p.out_batch_rows <= 0 || p.out_batch_rows == p.rows_per_batch) && p.rows_per_batch >= 16 * 256;
}

// A: (M, K) bf16 rows lda apart; W: (N, K).  Tensor maps: A box (64, 128), W box (64, 128) -- one half tile per CTA.
int launch_gemm2(const void* A, long long lda, const void* W, GemmParams p, cudaStream_t stream) {
  B200W_TRY(init_gemm2());
  CUtensorMap ta, tb;
  B200W_TRY(make_tmap_a(&ta, A, 1, p.rows_per_batch, p.K, lda, (long long)p.rows_per_batch * lda));
  B200W_TRY(make_tmap_w(&tb, W, p.N, p.K, k2BN / 2));
  p.tiles_m_per_batch = ceil_div(p.rows_per_batch, 2 * k2BM);
  p.tiles_n = p.n_store / k2BN;
  long long g = (32ll << 20) / ((long long)2 * k2BM * p.K * 2);
  p.group_m = (int)(g < 4 ? 4 : (g > 74 ? 74 : g));
  const long long tiles = (long long)p.tiles_m_per_batch * p.tiles_n;
  int clusters = device_sm_count() / 2;
  if (tiles < clusters) clusters = (int)tiles;
  ProfScope prof_(p.tag ? p.tag : "gemm2", stream);
  gemm2_bf16_kernel<<<2 * clusters, k2Threads, k2SmemBytes, stream>>>(ta, tb, p);
  B200W_LAUNCH_OK();
  count_launch();
  return kOk;
}

}  // namespace b200w
