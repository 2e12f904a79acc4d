// K5: bf16 GEMM on the 5th-generation tensor cores.  C[M,N] = epilogue(A[M,K] * W[N,K]^T).
//
// Replaces every `nn.Linear` / `nn.Conv1d` / `token_embedding.as_linear` matmul that mlx_whisper's
// AudioEncoder / TextDecoder issue (reference call site
// /root/reference/run:3-6).
//
// Structure (persistent, warp-specialised, one CTA per SM):
//   warp 0   : TMA producer   - cp.async.bulk.tensor tiles of A (128 x 64) and W (BN x 64), 128-byte swizzle,
//                               kStages-deep mbarrier ring
//   warp 1   : MMA issuer     - one lane issues tcgen05.mma.cta_group::1.kind::f16 (128 x BN x 16), fp32
//                               accumulators live in TMEM, double buffered across tiles
//   warp 2   : TMEM allocator
//   warps 4-11: epilogue      - tcgen05.ld 32x32b -> shared-memory transpose -> bias / GELU / fp32 residual (or
//                               positional table) -> fully coalesced bf16 or fp32 global stores
// A is addressed through a 3-D tensor map (k, row-in-batch, batch) so that the conv stem runs as an
// implicit GEMM: with a (T+2, C) zero-padded NLC slab the im2col row of output t is the contiguous
// 3*C span starting at padded row stride*t, i.e. just a tensor map with an overlapping row stride.
#include "common.cuh"
#include "kernels.h"

namespace b200w {

constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kGemmThreads = 384;            // 4 role warps + 8 epilogue warps
constexpr int kEpiStageBytes = 8 * 32 * 33 * 4;  // per-warp padded 32x32 f32 transpose tiles

template <int BN>
struct GemmCfg {
  static constexpr int kABytes = kBM * kBK * 2;
  static constexpr int kBBytes = BN * kBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  // BN <= 64 serves the decode GEMMs: TMA latency (~1.2 us) over ~0.3 us per K block needs the deep ring
  static constexpr int kStages = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int kTmemCols = (2 * BN < 32) ? 32 : 2 * BN;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/ + kEpiStageBytes;
};

template <int BN>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                 const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ unsigned char gemm_smem_raw[];
  // 1024-byte alignment for the 128B-swizzle atoms
  unsigned char* smem = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(gemm_smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + Cfg::kStages;
  uint64_t* tmem_full_bar = bars + 2 * Cfg::kStages;
  uint64_t* tmem_empty_bar = bars + 2 * Cfg::kStages + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * Cfg::kStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int tiles_m = p.n_batch * p.tiles_m_per_batch;
  const int num_tiles = tiles_m * p.tiles_n * p.split_k;
  const int num_kb = (p.K + kBK - 1) / kBK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    for (int i = 0; i < Cfg::kStages; ++i) {
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
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // everything above overlapped the previous kernel's tail; operands and outputs are touched below
  pdl_launch_dependents();  // after the wait: at most one dependent grid is resident ahead of the running one

  // tile id -> (m tile, n tile): groups of group_m row tiles sweep all n tiles so the A rows of a group
  // stay L2-resident while W (small) is re-read from L2.
  // split-K: the K range of one (m, n) tile is cut into split_k slices handled by consecutive tile ids; each
  // slice stores its raw fp32 partial to its own slab and the consumer kernel sums the slabs.
  auto decode_tile = [&](int id, int& mt, int& nt, int& ks) {
    ks = id % p.split_k;
    id /= p.split_k;
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
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int mt, nt, ks;
        decode_tile(tile, mt, nt, ks);
        const int b = mt / p.tiles_m_per_batch;
        const int t0 = (mt - b * p.tiles_m_per_batch) * kBM;
        const int n0 = nt * BN;
        const int kb_end = min(num_kb, (ks + 1) * p.kb_per_split);
        for (int kb = ks * p.kb_per_split; kb < kb_end; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          unsigned char* sa = smem + stage * Cfg::kStageBytes;
          unsigned char* sb = sa + Cfg::kABytes;
          mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          tma_load_3d(sa, &tma_a, &full_bar[stage], kb * kBK, t0, b);
          tma_load_3d(sb, &tma_b, &full_bar[stage], kb * kBK, n0, 0);
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
    __syncwarp();
  } else if (__shfl_sync(0xffffffffu, warp, 0) == 1) {
    // The whole warp walks the loop and one elected lane issues (elect_one(), common.cuh): the descriptors then live in
    // uniform registers and the four MMAs of a K block issue back to back (issued from a divergent single-thread branch
    // each cost a 54-cycle ELECT / R2UR loop -- as long as a 128 x 128 x 16 MMA itself).
    {
      constexpr uint32_t idesc = make_idesc_bf16(kBM, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const uint32_t s_base = smem_u32(smem);
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        const int kb_begin = (tile % p.split_k) * p.kb_per_split;
        const int kb_end = min(num_kb, kb_begin + p.kb_per_split);
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint32_t sa = s_base + stage * Cfg::kStageBytes;
          const uint64_t a_desc = make_sw128_desc(sa);
          const uint64_t b_desc = make_sw128_desc(sa + Cfg::kABytes);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            // +32 bytes along the swizzled row per K=16 slice -> +2 in the (addr >> 4) field
            if (elect_one()) umma_f16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb > kb_begin || k != 0) ? 1u : 0u);
          }
          if (elect_one()) umma_commit(&empty_bar[stage]);
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (elect_one()) umma_commit(&tmem_full_bar[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // Epilogue: 8 warps.  Warp e reads TMEM lane quarter e % 4 (a hardware constraint: warp id % 4) and the 32-column
    // chunks c = e / 4, e / 4 + 2, ...  A chunk arrives row-per-lane (tcgen05.ld 32x32b), is transposed through a
    // padded per-warp shared-memory tile and leaves column-per-lane, so that every global access of a warp is one
    // full 128-byte row segment (f32) or two 64-byte row segments (bf16): bias, residual and stores are coalesced.
    const int e = warp - 4;
    const int q = e & 3, hh = e >> 2;
    float* stage = reinterpret_cast<float*>(smem + Cfg::kStages * Cfg::kStageBytes + 256) + e * (32 * 33);
    int acc = 0;
    uint32_t acc_phase = 0;
    // the 16-byte f32 epilogue needs aligned rows (whole 32-column chunks are always stored: ldc >= n_store rounded up)
    const bool vec_f32 = p.out_f32 && (p.ldc & 3) == 0 && (p.split_stride & 3) == 0 &&
                         (reinterpret_cast<uintptr_t>(p.out) & 15) == 0 &&
                         (p.resid == nullptr || ((p.resid_ld & 3) == 0 && (reinterpret_cast<uintptr_t>(p.resid) & 15) == 0)) &&
                         (p.bias == nullptr || (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0);
    const bool vec_bf16 = !p.out_f32 && (p.ldc & 7) == 0 && ((p.out_batch_rows * p.ldc) & 7) == 0 &&
                          (reinterpret_cast<uintptr_t>(p.out) & 15) == 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      int mt, nt, ks;
      decode_tile(tile, mt, nt, ks);
      const int b = mt / p.tiles_m_per_batch;
      const int t0 = (mt - b * p.tiles_m_per_batch) * kBM + q * 32;  // first row (within the slab) of this warp
      const int n0 = nt * BN;
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tcgen05_fence_after();
      const uint32_t t_base = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
#pragma unroll 1
      for (int c = hh; c < BN / 32; c += 2) {
        const int col = n0 + c * 32;
        if (col >= p.n_store) break;  // warp-uniform
        const int rows_here = min(32, p.rows_per_batch - t0);  // rows of this warp that exist (may be <= 0)
        if (vec_f32) {
          // f32 output, 16-byte form: a lane owns four consecutive columns (sub) of rows rq, rq + 4, ... so that one
          // instruction moves four full 128-byte row segments -- a quarter of the load / store instructions of the
          // scalar form below, which kept the out-projection (K = d) waiting on its own epilogue.  Reading the padded
          // transpose tile as 4 scalars per lane is conflict-free (row stride 33: the four rows of an instruction start
          // one bank apart, the eight column groups four banks apart).
          const int sub = lane & 7, rq = lane >> 3;
          float4 rs4[8];
          if (p.resid != nullptr) {
            const long long lrow0 = (long long)b * p.rows_per_batch + t0;
            const float* rp = p.resid + col + 4 * sub;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int rr = 4 * i + rq;
              long long ridx = lrow0 + rr;
              if (p.resid_mod > 0) ridx %= p.resid_mod;
              rs4[i] = (rr < rows_here) ? *reinterpret_cast<const float4*>(rp + ridx * p.resid_ld) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) rs4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
          uint32_t r[32];
          tmem_ld_32x32(t_base + c * 32, r);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; ++j) stage[lane * 33 + j] = __uint_as_float(r[j]);
          __syncwarp();
          float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (p.bias != nullptr) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col + 4 * sub));
          float* op = reinterpret_cast<float*>(p.out) + ks * p.split_stride + ((long long)b * p.out_batch_rows + t0) * p.ldc + col +
                      4 * sub;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = 4 * i + rq;
            const float* sp = stage + rr * 33 + 4 * sub;
            float4 v = make_float4(sp[0] + b4.x, sp[1] + b4.y, sp[2] + b4.z, sp[3] + b4.w);
            if (p.gelu) {
              v.x = gelu_fast(v.x);
              v.y = gelu_fast(v.y);
              v.z = gelu_fast(v.z);
              v.w = gelu_fast(v.w);
            }
            v.x += rs4[i].x;
            v.y += rs4[i].y;
            v.z += rs4[i].z;
            v.w += rs4[i].w;
            if (rr < rows_here) *reinterpret_cast<float4*>(op + (long long)rr * p.ldc) = v;
          }
          __syncwarp();
          continue;
        }
        // residual rows first: 32 independent coalesced loads in flight while the accumulator is fetched and transposed
        float rsd[32];
        if (p.out_f32 && p.resid != nullptr) {
          const long long lrow0 = (long long)b * p.rows_per_batch + t0;
          long long ridx = (p.resid_mod > 0) ? (lrow0 % p.resid_mod) : lrow0;
          const float* rp = p.resid + col + lane;
#pragma unroll
          for (int rr = 0; rr < 32; ++rr) {
            rsd[rr] = (rr < rows_here) ? rp[ridx * p.resid_ld] : 0.0f;
            if (++ridx == p.resid_mod) ridx = 0;
          }
        } else {
#pragma unroll
          for (int rr = 0; rr < 32; ++rr) rsd[rr] = 0.0f;
        }
        uint32_t r[32];
        tmem_ld_32x32(t_base + c * 32, r);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; ++j) stage[lane * 33 + j] = __uint_as_float(r[j]);
        __syncwarp();
        if (p.out_f32) {
          const float bias_l = (p.bias != nullptr) ? __ldg(p.bias + col + lane) : 0.0f;
          float* op = reinterpret_cast<float*>(p.out) + ks * p.split_stride + ((long long)b * p.out_batch_rows + t0) * p.ldc +
                      col + lane;
#pragma unroll
          for (int rr = 0; rr < 32; ++rr) {
            float v = stage[rr * 33 + lane] + bias_l;
            if (p.gelu) v = gelu_fast(v);
            v += rsd[rr];
            if (rr < rows_here) op[(long long)rr * p.ldc] = v;
          }
        } else if (vec_bf16) {
          // bf16 output, 16-byte form: a lane owns eight consecutive columns of rows r8, r8 + 8, ...: one instruction
          // stores eight full 64-byte row segments (conflict-free reads: bank = r8 + 8 * s4 + j mod 32)
          const int s4 = lane & 3, r8 = lane >> 2;
          float bs[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) bs[j] = (p.bias != nullptr) ? __ldg(p.bias + col + 8 * s4 + j) : 0.0f;
          __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + ((long long)b * p.out_batch_rows + t0) * p.ldc + col + 8 * s4;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int rr = 8 * i + r8;
            const float* sp = stage + rr * 33 + 8 * s4;
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
          // two rows per pass: lanes 0-15 the even row, lanes 16-31 the odd one, two adjacent columns per lane
          const int l2 = (lane & 15) * 2, hi = lane >> 4;
          float b0 = 0.0f, b1 = 0.0f;
          if (p.bias != nullptr) {
            b0 = __ldg(p.bias + col + l2);
            b1 = __ldg(p.bias + col + l2 + 1);
          }
          __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + ((long long)b * p.out_batch_rows + t0) * p.ldc + col + l2;
#pragma unroll 4
          for (int rr = hi; rr < rows_here; rr += 2) {
            float v0 = stage[rr * 33 + l2] + b0, v1 = stage[rr * 33 + l2 + 1] + b1;
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
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// Opt every instantiation into its dynamic shared memory size once, up front (never inside a stream capture).
int init_gemm() {
  static bool done = false;
  if (done) return kOk;
  B200W_CUDA_OK(cudaFuncSetAttribute(gemm_bf16_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<256>::kSmemBytes));
  B200W_CUDA_OK(cudaFuncSetAttribute(gemm_bf16_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<128>::kSmemBytes));
  B200W_CUDA_OK(cudaFuncSetAttribute(gemm_bf16_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<64>::kSmemBytes));
  B200W_CUDA_OK(cudaFuncSetAttribute(gemm_bf16_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<32>::kSmemBytes));
  done = true;
  return kOk;
}

template <int BN>
static int launch_gemm_bn(const CUtensorMap& ta, const CUtensorMap& tb, GemmParams p, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  B200W_TRY(init_gemm());
  if (p.out_batch_rows <= 0) p.out_batch_rows = p.rows_per_batch;
  p.tiles_m_per_batch = ceil_div(p.rows_per_batch, kBM);
  p.tiles_n = ceil_div(p.n_store, BN);
  const int num_kb = ceil_div(p.K, kBK);
  if (p.split_k <= 1) {
    p.split_k = 1;
    p.kb_per_split = num_kb;
  } else {
    B200W_CHECK_ARG(p.out_f32 && !p.bias && !p.resid && !p.gelu, "gemm: split-K writes raw fp32 partials only");
    p.kb_per_split = ceil_div(num_kb, p.split_k);
    p.split_k = ceil_div(num_kb, p.kb_per_split);  // no empty slices
  }
  const long long tiles = (long long)p.n_batch * p.tiles_m_per_batch * p.tiles_n * p.split_k;
  B200W_CHECK_ARG(tiles > 0 && tiles < (1ll << 30), "gemm: tile count out of range");
  long long g = (32ll << 20) / ((long long)kBM * p.K * 2);
  p.group_m = (int)(g < 8 ? 8 : (g > 148 ? 148 : g));
  const int grid = (int)(tiles < device_sm_count() ? tiles : device_sm_count());
  ProfScope prof_(p.tag ? p.tag : "gemm", stream);
  B200W_CUDA_OK(launch_k(gemm_bf16_kernel<BN>, dim3(grid), dim3(kGemmThreads), Cfg::kSmemBytes, stream, ta, tb, p));
  count_launch();
  return kOk;
}

int gemm_block_n(int n_batch, int rows_per_batch, int n_store) {
  const long long m_tiles = (long long)n_batch * ceil_div(rows_per_batch, kBM);
  // few row tiles (decode steps): narrow N tiles so that enough CTAs stream the weights
  if (m_tiles <= 2) return (n_store % 64 == 0 || n_store > 4096) ? 64 : 32;
  if (n_store % 256 == 0 && m_tiles * (n_store / 256) >= 2 * 148) return 256;
  return 128;
}

int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int block_n, cudaStream_t stream) {
  B200W_CHECK_ARG(p.K > 0 && p.K % 8 == 0, "gemm: K must be a positive multiple of 8 (got %d)", p.K);
  B200W_CHECK_ARG(p.n_store > 0 && p.ldc >= ((p.n_store + 31) / 32) * 32, "gemm: ldc %lld too small for n_store %d",
                  p.ldc, p.n_store);
  B200W_CHECK_ARG(p.bias == nullptr || p.n_store % 32 == 0, "gemm: bias needs N %% 32 == 0");
  switch (block_n) {
    case 256: return launch_gemm_bn<256>(ta, tb, p, stream);
    case 128: return launch_gemm_bn<128>(ta, tb, p, stream);
    case 64: return launch_gemm_bn<64>(ta, tb, p, stream);
    case 32: return launch_gemm_bn<32>(ta, tb, p, stream);
    default: set_last_error("gemm: unsupported block_n %d", block_n); return kErrInvalidArgument;
  }
}

// Tensor map for the weight operand W (N, K) row-major bf16, box (64, block_n).
int make_tmap_w(CUtensorMap* out, const void* w, int N, int K, int block_n) {
  uint64_t dims[3] = {(uint64_t)K, (uint64_t)N, 1};
  uint64_t strides[2] = {(uint64_t)K * 2, (uint64_t)K * 2 * (uint64_t)N};
  uint32_t box[3] = {kBK, (uint32_t)block_n, 1};
  return encode_tmap_bf16(out, w, 3, dims, strides, box);
}

// Tensor map for the activation operand: `n_batch` slabs, each `rows` logical rows of K contiguous bf16,
// consecutive rows `row_stride` elements apart, slabs `batch_stride` elements apart; box (64, 128, 1).
int make_tmap_a(CUtensorMap* out, const void* a, int n_batch, int rows, int K, long long row_stride,
                long long batch_stride) {
  uint64_t dims[3] = {(uint64_t)K, (uint64_t)rows, (uint64_t)n_batch};
  uint64_t strides[2] = {(uint64_t)row_stride * 2, (uint64_t)batch_stride * 2};
  uint32_t box[3] = {kBK, kBM, 1};
  return encode_tmap_bf16(out, a, 3, dims, strides, box);
}

}  // namespace b200w
