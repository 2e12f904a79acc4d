// K6 / K7 / K8: attention kernels (reference: mlx_whisper/whisper.py::MultiHeadAttention.qkv_attention,
// reached from /root/reference/run:3-6).  Head dim is 64 for every Whisper size; the
// reference scales q and k by hd^-0.25 each, here the product hd^-0.5 is folded into the fp32 scores.
//
// K6 encoder MHA (non-causal, S = 1500): flash-style on tcgen05.  One CTA = 128 queries of one head, key blocks
//    of 64.  Q/K/V tiles arrive by TMA (128-byte swizzle) straight out of the fused QKV activation, K/V two blocks
//    ahead; S = Q K^T runs SS (both operands in shared memory), each of the 128 threads owns one query row (one
//    TMEM lane) so softmax needs no shuffles, P goes back to tensor memory and P V runs TS (A operand from TMEM,
//    V consumed MN-major); the output accumulates in TMEM and is rescaled only when a row maximum moves by > 2^8.
//    Three CTAs per SM overlap one CTA's exp work with the others' MMAs and TMEM round trips.
// K7 decoder self-attention over the paged KV cache (q-len 1, or a short prompt), CUDA cores.
// K8 decoder cross-attention (S = 1500), the HBM-bound hot spot of decoding: 16-byte coalesced loads,
//    8 lanes per key row, each K/V byte read exactly once per step.
#include <algorithm>
#include "common.cuh"
#include "kernels.h"

namespace b200w {

constexpr int kHd = 64;
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ float fast_exp2(float x) {
#ifdef B200W_EXP_PROBE  // timing probe only: how much of K6 is the MUFU pipe?
  return x * 0.001f;
#else
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
#endif
}

// thread-private cp.async rings of the decode attention kernels (K7r, K8r, K8p)
__device__ __forceinline__ void cross_cp16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cross_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cross_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// eight bf16 -> f32 with one integer instruction each (shift / mask)
__device__ __forceinline__ void cross_unpack8(const uint4& u, float (&f)[8]) {
  f[0] = __uint_as_float(u.x << 16); f[1] = __uint_as_float(u.x & 0xffff0000u);
  f[2] = __uint_as_float(u.y << 16); f[3] = __uint_as_float(u.y & 0xffff0000u);
  f[4] = __uint_as_float(u.z << 16); f[5] = __uint_as_float(u.z & 0xffff0000u);
  f[6] = __uint_as_float(u.w << 16); f[7] = __uint_as_float(u.w & 0xffff0000u);
}


// exp2 on the FMA / ALU pipes (Cody-Waite: round to the nearest integer with the 1.5 * 2^23 trick, cubic minimax of 2^f on
// [-0.5, 0.5], integer added into the exponent field): relative error <= 7.5e-5, invisible after the bf16 rounding of P.
// FlashAttention-4's software exp2 for a share of the probabilities: 9 FMA / ALU instructions at 128 lanes per clock against
// one MUFU at 16.  Measured on B200 (K6, 40 windows, same box): none 0.727 ms, every 4th pair 0.772, every 3rd 0.795, every
// 2nd 0.853 -- K6 is bound by instruction issue and the TMEM / MMA round trips (ncu r02: issue 46 %, XU 61 %, FMA 17 %), not
// by the MUFU pipe alone, so the extra instructions cost more than the MUFU slots they free.  Off by default
// (-DB200W_ENC_POLY=n for A/B builds).
__device__ __forceinline__ float poly_exp2(float x) {
#ifdef B200W_EXP_PROBE
  return x * 0.001f;
#else
  x = fmaxf(x, -125.0f);
  const float r = x + 12582912.0f;
  const float f = x - (r - 12582912.0f);
  float p = fmaf(f, 0.0551716685f, 0.2426111251f);
  p = fmaf(p, f, 0.6932609677f);
  p = fmaf(p, f, 0.9999280572f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(r) << 23));
#endif
}
#ifndef B200W_ENC_POLY
#define B200W_ENC_POLY 0  // every B200W_ENC_POLY-th pair of probabilities is computed by poly_exp2 (0: none)
#endif
__device__ __forceinline__ bool enc_poly_pair(int i) { return B200W_ENC_POLY > 0 && (i % (B200W_ENC_POLY > 0 ? B200W_ENC_POLY : 1)) == (B200W_ENC_POLY - 1); }

// descriptor for an MN-major or K-major SW128 tile with explicit leading byte offset
__device__ __forceinline__ uint64_t make_sw128_desc_lbo(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// =============================================================================================== K6
// Template on the key block: BKV = 128 (2 CTAs per SM: 80 KB of tiles, 256 TMEM columns, the 128-wide score row in
// registers) or BKV = 64 (48 KB, 128 TMEM columns, 64-wide score row: up to 4 CTAs per SM).  Every stage of a key block
// is serialised inside a CTA (S MMA -> TMEM read -> exp -> P store -> PV MMA -> TMEM read), so the SM is kept busy by
// CTA-level overlap: more, smaller CTAs hide the MMA / TMEM round trips of each other.
constexpr int kEncThreads = 128;
constexpr int kBQ = 128;
constexpr int kQTileBytes = kBQ * kHd * 2;  // 16 KB: 128 rows x 128 B

// PT: the probabilities go back to tensor memory (tcgen05.st, over the S columns they were computed from) and P V reads
// its A operand from there (the TS form of tcgen05.mma): no P tile in shared memory, so neither its 16-32 KB store nor
// its read per block -- the SS form spends more shared-memory cycles than tensor cycles (A = 4 KB per 128xNx16 MMA).
#ifndef B200W_ENC_CTAS
#define B200W_ENC_CTAS 3
#endif
template <int BKV, bool PT>
struct EncCfg {
  static constexpr int kKvTileBytes = BKV * kHd * 2;          // K or V tile: BKV rows x 128 B
  static constexpr int kPTileBytes = PT ? 0 : kBQ * BKV * 2;  // P: BKV / 64 K-major sub-tiles of 128 rows x 128 B
  static constexpr int kSmem = kQTileBytes + 4 * kKvTileBytes + kPTileBytes + 1024 + 128;  // K and V double-buffered
  static constexpr int kTmemCols = (BKV + kHd <= 128) ? 128 : 256;  // S (BKV; P over its first half) + O (64)
  static constexpr int kCtasPerSm = (BKV == 64) ? (PT ? B200W_ENC_CTAS : 3) : 2;
};

template <int BKV, bool PT>
__global__ void __launch_bounds__(kEncThreads, EncCfg<BKV, PT>::kCtasPerSm)
encoder_attention_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv, int T, int d,
                         __nv_bfloat16* __restrict__ out) {
  using Cfg = EncCfg<BKV, PT>;
  extern __shared__ unsigned char att_smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(att_smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  unsigned char* sQ = smem;
  unsigned char* sK = sQ + kQTileBytes;             // two K tiles (block j lives in buffer j & 1)
  unsigned char* sV = sK + 2 * Cfg::kKvTileBytes;   // two V tiles
  unsigned char* sP = sV + 2 * Cfg::kKvTileBytes;   // BKV / 64 K-major sub-tiles (keys 0-63, 64-127) of 16 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + Cfg::kPTileBytes);
  uint64_t* bar_q = bars + 0;
  uint64_t* bar_k = bars + 1;  // [2]
  uint64_t* bar_v = bars + 3;  // [2]
  uint64_t* bar_s = bars + 5;
  uint64_t* bar_o = bars + 6;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7);

  const int tid = threadIdx.x, warp = tid >> 5;
  const bool w0 = __shfl_sync(0xffffffffu, warp, 0) == 0;  // provably warp-uniform: warp 0 issues every MMA (one elected lane)
  const int q0 = blockIdx.x * kBQ, h = blockIdx.y, b = blockIdx.z;
  const int nkv = (T + BKV - 1) / BKV;

  if (tid == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_kv);
    mbar_init(bar_q, 1);
    mbar_init(&bar_k[0], 1);
    mbar_init(&bar_k[1], 1);
    mbar_init(&bar_v[0], 1);
    mbar_init(&bar_v[1], 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_o, 1);
    fence_barrier_init();
  }
  __syncwarp();
  if (warp == 0) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_s = tmem_base + ((uint32_t)(warp * 32) << 16);
  const uint32_t tmem_o = tmem_s + BKV;
  pdl_wait();
  pdl_launch_dependents();  // after the wait: at most one dependent grid is resident ahead of the running one

  // K / V tiles are requested two blocks ahead: a TMA round trip (~1 us from L2) is longer than one block's work, and
  // with single buffers every block waited for its tiles (the first version: 437 TFLOP/s whatever else was changed)
  auto load_k = [&](int blk) {
    mbar_expect_tx(&bar_k[blk & 1], Cfg::kKvTileBytes);
    tma_load_3d(sK + (blk & 1) * Cfg::kKvTileBytes, &tm_kv, &bar_k[blk & 1], d + h * kHd, blk * BKV, b);
  };
  auto load_v = [&](int blk) {
    mbar_expect_tx(&bar_v[blk & 1], Cfg::kKvTileBytes);
    tma_load_3d(sV + (blk & 1) * Cfg::kKvTileBytes, &tm_kv, &bar_v[blk & 1], 2 * d + h * kHd, blk * BKV, b);
  };
  if (tid == 0) {
    mbar_expect_tx(bar_q, kQTileBytes);
    tma_load_3d(sQ, &tm_q, bar_q, h * kHd, q0, b);
    load_k(0);
    load_v(0);
    if (nkv > 1) {
      load_k(1);
      load_v(1);
    }
  }

  constexpr uint32_t idesc_s = make_idesc_bf16(128, BKV, 0, 0);
  constexpr uint32_t idesc_o = make_idesc_bf16(128, 64, 0, 1);  // B (= V) is MN-major
  const float c = 0.125f * kLog2e;                              // hd^-0.5 in the exp2 domain

  // Software pipeline of one CTA (every MMA is issued by thread 0, every barrier below is CTA-wide):
  //   S(j+1) = Q K(j+1)^T is issued as soon as all four warps hold S(j) in registers, so it runs under softmax(j);
  //   O accumulates in TMEM across blocks (no per-block read-back); PV(j) is only waited for when block j+1 needs
  //   the P buffer again.  The running maximum used for exp2 is allowed to lag the true maximum by up to 2^8 (the
  //   probabilities then reach 256, harmless in bf16 / fp32): O and l are rescaled only when a row's maximum grows by
  //   more than that, which after the first blocks is rare -- the FlashAttention-4 "conditional rescaling".
  float m_used = -INFINITY, l_run = 0.0f;
  const int row = tid;  // query row within the tile == TMEM lane

  if (w0) {
    mbar_wait(bar_q, 0);
    mbar_wait(&bar_k[0], 0);
    tcgen05_fence_after();
    const uint64_t qd = make_sw128_desc(smem_u32(sQ));
    const uint64_t kd = make_sw128_desc(smem_u32(sK));
#pragma unroll
    for (int k = 0; k < kHd / 16; ++k)
      if (elect_one()) umma_f16(tmem_base, qd + 2 * k, kd + 2 * k, idesc_s, k != 0);
    if (elect_one()) umma_commit(bar_s);
  }
  __syncwarp();

  for (int j = 0; j < nkv; ++j) {
    mbar_wait(bar_s, j & 1);
    tcgen05_fence_after();
    float sc[BKV];
    {
      uint32_t r[BKV];
#pragma unroll
      for (int cb = 0; cb < BKV / 32; ++cb) tmem_ld_32x32(tmem_s + cb * 32, reinterpret_cast<uint32_t(&)[32]>(r[cb * 32]));
      tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < BKV; ++i) sc[i] = __uint_as_float(r[i]);
    }
    auto issue_next_s = [&]() {  // all of warp 0
      if (j + 1 < nkv) {
        mbar_wait(&bar_k[(j + 1) & 1], ((j + 1) >> 1) & 1);
        tcgen05_fence_after();
        const uint64_t qd = make_sw128_desc(smem_u32(sQ));
        const uint64_t kd = make_sw128_desc(smem_u32(sK + ((j + 1) & 1) * Cfg::kKvTileBytes));
#pragma unroll
        for (int k = 0; k < kHd / 16; ++k)
          if (elect_one()) umma_f16(tmem_base, qd + 2 * k, kd + 2 * k, idesc_s, k != 0);
        if (elect_one()) umma_commit(bar_s);
      }
    };
    if constexpr (!PT) {
      tcgen05_fence_before();
      __syncthreads();  // every warp holds its S(j) rows: the S columns and K buffer j & 1 are free
      if (w0) {
        tcgen05_fence_after();
        if (tid == 0 && j + 2 < nkv) load_k(j + 2);
        __syncwarp();
        issue_next_s();
      }
    } else {
      if (tid == 0 && j + 2 < nkv) load_k(j + 2);  // S(j) is complete: K buffer j & 1 is free
    }
    __syncwarp();

    const int kv_valid = min(BKV, T - j * BKV);
    if (kv_valid < BKV) {  // only the last key block is ragged (block-uniform branch)
#pragma unroll
      for (int i = 0; i < BKV; ++i)
        if (i >= kv_valid) sc[i] = -INFINITY;
    }
    float mx8[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) mx8[i] = sc[i];
#pragma unroll
    for (int i = 8; i < BKV; ++i) mx8[i & 7] = fmaxf(mx8[i & 7], sc[i]);
    const float mx = fmaxf(fmaxf(fmaxf(mx8[0], mx8[1]), fmaxf(mx8[2], mx8[3])), fmaxf(fmaxf(mx8[4], mx8[5]), fmaxf(mx8[6], mx8[7])));
    const float m_cand = fmaxf(m_used, mx * c);

    if (j > 0) {  // PV(j-1) done: P buffer and V buffer (j-1) & 1 are free, O is consistent
      mbar_wait(bar_o, (j - 1) & 1);
      tcgen05_fence_after();
      if (tid == 0 && j + 1 < nkv) load_v(j + 1);
      __syncwarp();
      if (__any_sync(0xffffffffu, m_cand > m_used + 8.0f)) {  // warp-uniform: tcgen05.ld / st are warp-collective
        const float alpha = fast_exp2(m_used - m_cand);     // 1 for the rows of this warp whose maximum did not move
        m_used = m_cand;
        l_run *= alpha;
        uint32_t r[kHd];
        tmem_ld_32x32(tmem_o, reinterpret_cast<uint32_t(&)[32]>(r[0]));
        tmem_ld_32x32(tmem_o + 32, reinterpret_cast<uint32_t(&)[32]>(r[32]));
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < kHd; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
        tmem_st_32x32(tmem_o, reinterpret_cast<const uint32_t(&)[32]>(r[0]));
        tmem_st_32x32(tmem_o + 32, reinterpret_cast<const uint32_t(&)[32]>(r[32]));
        tmem_wait_st();
      }
    } else {
      m_used = m_cand;
    }

    // probabilities -> swizzled bf16 A operand (exp2(-inf) = 0 for the masked tail)
    float ps4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    if constexpr (PT) {
      // bf16 pairs (keys 2c, 2c + 1) -> TMEM column c of this thread's lane, over the S columns it has just read
      uint32_t pk[BKV / 2];
#pragma unroll
      for (int i = 0; i < BKV / 2; ++i) {
        const float x0 = fmaf(sc[2 * i], c, -m_used), x1 = fmaf(sc[2 * i + 1], c, -m_used);
        const float p0 = enc_poly_pair(i) ? poly_exp2(x0) : fast_exp2(x0);
        const float p1 = enc_poly_pair(i) ? poly_exp2(x1) : fast_exp2(x1);
        ps4[i & 3] += p0 + p1;
        pk[i] = pack_bf16x2(p0, p1);
      }
#pragma unroll
      for (int cb = 0; cb < BKV / 64; ++cb) tmem_st_32x32(tmem_s + cb * 32, reinterpret_cast<const uint32_t(&)[32]>(pk[cb * 32]));
      tmem_wait_st();
    } else {
#pragma unroll
      for (int cb = 0; cb < BKV / 32; ++cb) {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float x0 = fmaf(sc[cb * 32 + 2 * i], c, -m_used), x1 = fmaf(sc[cb * 32 + 2 * i + 1], c, -m_used);
          const float p0 = enc_poly_pair(i) ? poly_exp2(x0) : fast_exp2(x0);
          const float p1 = enc_poly_pair(i) ? poly_exp2(x1) : fast_exp2(x1);
          ps4[i & 3] += p0 + p1;
          pk[i] = pack_bf16x2(p0, p1);
        }
        // columns cb*32 .. +31 -> sub-tile cb/2, 16-byte pieces (cb%2)*4 .. +3, XOR-swizzled with row%8
        unsigned char* base = sP + (cb >> 1) * kQTileBytes + row * 128;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int piece = ((cb & 1) * 4 + q) ^ (row & 7);
          *reinterpret_cast<uint4*>(base + piece * 16) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
        }
      }
      fence_proxy_async_smem();
    }
    l_run += (ps4[0] + ps4[1]) + (ps4[2] + ps4[3]);
    tcgen05_fence_before();
    __syncthreads();  // P(j) (and a rescaled O) are in place
    if (w0) {
      tcgen05_fence_after();
      mbar_wait(&bar_v[j & 1], (j >> 1) & 1);
      tcgen05_fence_after();
      const uint32_t v_base = smem_u32(sV + (j & 1) * Cfg::kKvTileBytes);
#pragma unroll
      for (int k = 0; k < BKV / 16; ++k) {
        const uint64_t vd = make_sw128_desc_lbo(v_base + k * 2048, BKV * 128);
        if constexpr (PT) {
          if (elect_one()) umma_f16_ts(tmem_base + BKV, tmem_base + k * 8, vd, idesc_o, (j > 0 || k != 0) ? 1u : 0u);
        } else {
          const uint64_t pd = make_sw128_desc(smem_u32(sP + (k >> 2) * kQTileBytes)) + 2 * (k & 3);
          if (elect_one()) umma_f16(tmem_base + BKV, pd, vd, idesc_o, (j > 0 || k != 0) ? 1u : 0u);
        }
      }
      if (elect_one()) umma_commit(bar_o);
      // the tensor pipe runs the warp's MMAs in order: S(j+1) may overwrite the P columns only after P V(j) read them
      if constexpr (PT) issue_next_s();
    }
    __syncwarp();
  }

  mbar_wait(bar_o, (nkv - 1) & 1);
  tcgen05_fence_after();
  {  // (rows past T exist only in the last query tile; their lanes still take part in the collective TMEM load)
    uint32_t r[kHd];
    tmem_ld_32x32(tmem_o, reinterpret_cast<uint32_t(&)[32]>(r[0]));
    tmem_ld_32x32(tmem_o + 32, reinterpret_cast<uint32_t(&)[32]>(r[32]));
    tmem_wait_ld();
    if (q0 + row < T) {
      const float inv = 1.0f / l_run;
      uint4* dst = reinterpret_cast<uint4*>(out + ((long long)b * T + q0 + row) * d + h * kHd);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(r[8 * i + e]) * inv;
        dst[i] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// A/B switches: B200W_ENC_BKV=128 selects the wide-block form, B200W_ENC_PT=0 the form with P in shared memory
static int enc_bkv() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B200W_ENC_BKV");
    v = (e != nullptr && atoi(e) == 128) ? 128 : 64;
  }
  return v;
}
static bool enc_pt() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B200W_ENC_PT");
    v = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

int init_attention() {
  static bool done = false;
  if (done) return kOk;
  B200W_CUDA_OK(cudaFuncSetAttribute(encoder_attention_kernel<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, EncCfg<64, false>::kSmem));
  B200W_CUDA_OK(cudaFuncSetAttribute(encoder_attention_kernel<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, EncCfg<128, false>::kSmem));
  B200W_CUDA_OK(cudaFuncSetAttribute(encoder_attention_kernel<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, EncCfg<64, true>::kSmem));
  B200W_CUDA_OK(cudaFuncSetAttribute(encoder_attention_kernel<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, EncCfg<128, true>::kSmem));
  done = true;
  return kOk;
}

int launch_encoder_attention(const __nv_bfloat16* qkv, int n_batch, int T, int n_head, __nv_bfloat16* out,
                             cudaStream_t stream) {
  B200W_CHECK_ARG(n_batch > 0 && n_batch <= 65535 && T > 0 && n_head > 0, "encoder_attention: bad sizes");
  const int d = n_head * kHd;
  const int bkv = enc_bkv();
  CUtensorMap tq, tkv;
  uint64_t dims[3] = {(uint64_t)(3 * d), (uint64_t)T, (uint64_t)n_batch};
  uint64_t strides[2] = {(uint64_t)(3 * d) * 2, (uint64_t)T * 3 * d * 2};
  uint32_t box_q[3] = {kHd, kBQ, 1};
  uint32_t box_kv[3] = {kHd, (uint32_t)bkv, 1};
  B200W_TRY(encode_tmap_bf16(&tq, qkv, 3, dims, strides, box_q));
  B200W_TRY(encode_tmap_bf16(&tkv, qkv, 3, dims, strides, box_kv));
  B200W_TRY(init_attention());
  dim3 grid(ceil_div(T, kBQ), n_head, n_batch);
  ProfScope prof_("encoder_attention", stream);
  const bool pt = enc_pt();
  if (bkv == 64 && pt)
    B200W_CUDA_OK(launch_k(encoder_attention_kernel<64, true>, grid, dim3(kEncThreads), EncCfg<64, true>::kSmem, stream, tq, tkv, T, d, out));
  else if (bkv == 64)
    B200W_CUDA_OK(launch_k(encoder_attention_kernel<64, false>, grid, dim3(kEncThreads), EncCfg<64, false>::kSmem, stream, tq, tkv, T, d, out));
  else if (pt)
    B200W_CUDA_OK(launch_k(encoder_attention_kernel<128, true>, grid, dim3(kEncThreads), EncCfg<128, true>::kSmem, stream, tq, tkv, T, d, out));
  else
    B200W_CUDA_OK(launch_k(encoder_attention_kernel<128, false>, grid, dim3(kEncThreads), EncCfg<128, false>::kSmem, stream, tq, tkv, T, d, out));
  count_launch();
  return kOk;
}

// =============================================================================================== K7
// One CTA of 4 warps per (sequence, head, query), the access pattern of K8: 8 lanes per key row (16 B each), 4 rows
// per warp per load instruction, 8 loads in flight per thread (128 keys per CTA sweep), the first sweep of V issued
// before the softmax barrier.  Sized for <= 448 cached keys.  (The first version used one warp per unit: at 220
// cached keys it ran at 1.6 TB/s, latency-bound by its 2 KB in flight per warp.)
constexpr int kSelfWarps = 4;
constexpr int kSelfThreads = kSelfWarps * 32;
constexpr int kMaxSelfKeys = 448;
constexpr int kSelfUnroll = 8;
constexpr int kSelfSweep = kSelfWarps * 4 * kSelfUnroll;  // keys per CTA sweep

__global__ void __launch_bounds__(kSelfThreads)
decoder_self_attention_kernel(const __nv_bfloat16* __restrict__ qkv, int n_seq, int n_q, int n_head,
                              const int* __restrict__ pos, __nv_bfloat16* k_pages, __nv_bfloat16* v_pages,
                              const int* __restrict__ block_table, int max_pages, int page_size,
                              __nv_bfloat16* __restrict__ out, const float* __restrict__ part, int n_split,
                              long long split_stride, const float* __restrict__ bias, const int* __restrict__ finished) {
  __shared__ float s_p[kMaxSelfKeys];
  __shared__ float s_red[kSelfWarps];
  __shared__ float s_part[kSelfWarps][kHd];
  __shared__ uint4 s_q[8];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int unit = blockIdx.x;  // ((b * n_head) + h) * n_q + qi
  pdl_wait();
  pdl_launch_dependents();  // after the wait: at most one dependent grid is resident ahead of the running one
  const int qi = unit % n_q, h = (unit / n_q) % n_head, b = unit / (n_q * n_head);
  if (finished != nullptr && finished[b]) return;  // the sequence has emitted EOT: its row is ignored from here on
  const int sub = lane & 7, kg = lane >> 3;
  const int d = n_head * kHd;
  const int p0 = pos[b];           // tokens already cached for this sequence
  const int n_keys = p0 + qi + 1;  // causal
  const long long my_row = ((long long)b * n_q + qi) * 3 * d;
  const bool split_mode = n_split > 0;  // qkv arrives as split-K fp32 partial slabs (+ bias), n_q == 1
  const int* bt = block_table + b * max_pages;
  const float c = 0.125f * kLog2e;
  const int pshift = __ffs(page_size) - 1;  // page_size is a power of two (checked by the launcher)
  auto page_row = [&](int j) -> long long {
    return ((long long)bt[j >> pshift] * page_size + (j & (page_size - 1))) * d + h * kHd + sub * 8;
  };

  // ---- this token's q (registers, 8 dims per lane) and k / v rows (appended to the paged cache) ----
  float qv[8];
  {
    uint4 qu;
    if (split_mode) {
      // warp 0: lanes 0-7 reduce q, 8-15 k, 16-23 v (8 consecutive dims each); q goes to the CTA through s_q
      if (warp == 0 && kg < 3) {
        float v8[8];
        const long long col = (long long)kg * d + h * kHd + sub * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) v8[i] = bias[col + i];
        for (int sidx = 0; sidx < n_split; ++sidx) {
          const float* pp = part + sidx * split_stride + (long long)b * 3 * d + col;
          const float4 a0 = *reinterpret_cast<const float4*>(pp), a1 = *reinterpret_cast<const float4*>(pp + 4);
          v8[0] += a0.x; v8[1] += a0.y; v8[2] += a0.z; v8[3] += a0.w;
          v8[4] += a1.x; v8[5] += a1.y; v8[6] += a1.z; v8[7] += a1.w;
        }
        const uint4 packed = make_uint4(pack_bf16x2(v8[0], v8[1]), pack_bf16x2(v8[2], v8[3]), pack_bf16x2(v8[4], v8[5]),
                                        pack_bf16x2(v8[6], v8[7]));
        const long long dst = page_row(p0);
        if (kg == 0) s_q[sub] = packed;  // q rounded to bf16 like the GEMM epilogue would have
        if (kg == 1) *reinterpret_cast<uint4*>(k_pages + dst) = packed;
        if (kg == 2) *reinterpret_cast<uint4*>(v_pages + dst) = packed;
      }
      __syncthreads();  // s_q and this CTA's appended rows are visible to all four warps
      qu = s_q[sub];
    } else {
      qu = *reinterpret_cast<const uint4*>(qkv + my_row + h * kHd + sub * 8);
      if (warp == 0 && (kg == 1 || kg == 2)) {
        const long long dst = page_row(p0 + qi);
        *reinterpret_cast<uint4*>((kg == 1 ? k_pages : v_pages) + dst) =
            *reinterpret_cast<const uint4*>(qkv + my_row + (long long)kg * d + h * kHd + sub * 8);
      }
    }
    const float2 a0 = unpack_bf16x2(qu.x), a1 = unpack_bf16x2(qu.y), a2 = unpack_bf16x2(qu.z), a3 = unpack_bf16x2(qu.w);
    qv[0] = a0.x * c; qv[1] = a0.y * c; qv[2] = a1.x * c; qv[3] = a1.y * c;
    qv[4] = a2.x * c; qv[5] = a2.y * c; qv[6] = a3.x * c; qv[7] = a3.y * c;
  }
  // rows of position j: cached pages for j < p0 (and, in split mode, the row this CTA just appended); this step's
  // qkv rows otherwise (other CTAs append those concurrently).  Branch-free on purpose: positions past the end are
  // clamped to the last key (their values are masked by the callers), the page ids of a sweep are fetched first and
  // the row loads follow back to back -- with per-load branches the compiler put every row's unpacking between the
  // loads and the sweep ran one load at a time.
  const long long qkv_seq = (long long)b * n_q * 3 * d + h * kHd + sub * 8;
  auto load_rows = [&](uint4 (&u)[kSelfUnroll], int j0, int which) {
    int jc[kSelfUnroll], pg[kSelfUnroll];
#pragma unroll
    for (int i = 0; i < kSelfUnroll; ++i) {
      jc[i] = min(j0 + kg + i * (kSelfWarps * 4), n_keys - 1);
      pg[i] = bt[jc[i] >> pshift];
    }
    __nv_bfloat16* pages = which == 1 ? k_pages : v_pages;
#pragma unroll
    for (int i = 0; i < kSelfUnroll; ++i) {
      const bool fresh = jc[i] >= p0 && !split_mode;
      const __nv_bfloat16* base = fresh ? qkv : pages;
      const long long off = fresh ? qkv_seq + (long long)(jc[i] - p0) * 3 * d + (long long)which * d
                                  : ((long long)pg[i] * page_size + (jc[i] & (page_size - 1))) * d + h * kHd + sub * 8;
      u[i] = *reinterpret_cast<const uint4*>(base + off);
    }
  };

  // ---- scores ----
  float mx = -INFINITY;
  auto score_rows = [&](const uint4 (&u)[kSelfUnroll], int j0) {
#pragma unroll
    for (int i = 0; i < kSelfUnroll; ++i) {
      const int j = j0 + kg + i * (kSelfWarps * 4);
      const float2 a0 = unpack_bf16x2(u[i].x), a1 = unpack_bf16x2(u[i].y), a2 = unpack_bf16x2(u[i].z), a3 = unpack_bf16x2(u[i].w);
      float sc = a0.x * qv[0];
      sc = fmaf(a0.y, qv[1], sc);
      sc = fmaf(a1.x, qv[2], sc);
      sc = fmaf(a1.y, qv[3], sc);
      sc = fmaf(a2.x, qv[4], sc);
      sc = fmaf(a2.y, qv[5], sc);
      sc = fmaf(a3.x, qv[6], sc);
      sc = fmaf(a3.y, qv[7], sc);
      sc += __shfl_xor_sync(0xffffffffu, sc, 1);
      sc += __shfl_xor_sync(0xffffffffu, sc, 2);
      sc += __shfl_xor_sync(0xffffffffu, sc, 4);
      if (j < n_keys) {
        if (sub == 0) s_p[j] = sc;
        mx = fmaxf(mx, sc);
      }
    }
  };
  for (int j0 = warp * 4; j0 < n_keys; j0 += kSelfSweep) {  // warp-uniform trip count: the shuffles need every lane
    uint4 u[kSelfUnroll];
    load_rows(u, j0, 1);
    score_rows(u, j0);
  }
  // the first sweep of V does not depend on the probabilities: put it in flight across the softmax barriers
  uint4 v0[kSelfUnroll];
  load_rows(v0, warp * 4, 2);
  mx = warp_max(mx);
  if (lane == 0) s_red[warp] = mx;
  __syncthreads();
  mx = s_red[0];
#pragma unroll
  for (int i = 1; i < kSelfWarps; ++i) mx = fmaxf(mx, s_red[i]);
  __syncthreads();
  float sum = 0.0f;
  for (int j = tid; j < n_keys; j += kSelfThreads) {
    const float p = fast_exp2(s_p[j] - mx);
    sum += p;
    s_p[j] = __bfloat162float(__float2bfloat16(p));  // bf16 probabilities, as in the tensor-core path
  }
  sum = warp_sum(sum);
  if (lane == 0) s_red[warp] = sum;
  __syncthreads();
  sum = 0.0f;
#pragma unroll
  for (int i = 0; i < kSelfWarps; ++i) sum += s_red[i];

  // ---- output ----
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.0f;
  auto accumulate = [&](const uint4 (&u)[kSelfUnroll], int j0) {
#pragma unroll
    for (int i = 0; i < kSelfUnroll; ++i) {
      const int j = j0 + kg + i * (kSelfWarps * 4);
      const float p = (j < n_keys) ? s_p[j] : 0.0f;
      const float2 a0 = unpack_bf16x2(u[i].x), a1 = unpack_bf16x2(u[i].y), a2 = unpack_bf16x2(u[i].z), a3 = unpack_bf16x2(u[i].w);
      acc[0] = fmaf(p, a0.x, acc[0]);
      acc[1] = fmaf(p, a0.y, acc[1]);
      acc[2] = fmaf(p, a1.x, acc[2]);
      acc[3] = fmaf(p, a1.y, acc[3]);
      acc[4] = fmaf(p, a2.x, acc[4]);
      acc[5] = fmaf(p, a2.y, acc[5]);
      acc[6] = fmaf(p, a3.x, acc[6]);
      acc[7] = fmaf(p, a3.y, acc[7]);
    }
  };
  accumulate(v0, warp * 4);
  for (int j0 = warp * 4 + kSelfSweep; j0 < n_keys; j0 += kSelfSweep) {
    uint4 u[kSelfUnroll];
    load_rows(u, j0, 2);
    accumulate(u, j0);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 8);
    acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 16);
  }
  if (kg == 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) s_part[warp][sub * 8 + i] = acc[i];
  }
  __syncthreads();
  if (tid < kHd) {
    float v = 0.0f;
#pragma unroll
    for (int w = 0; w < kSelfWarps; ++w) v += s_part[w][tid];
    out[((long long)b * n_q + qi) * d + h * kHd + tid] = __float2bfloat16(v / sum);
  }
}

// K7r: the decode-step form (n_q == 1, q | k | v as split-K slabs) with the cached rows streamed through a thread-private
// cp.async ring like K8r: one page-table entry (staged in shared memory) per request, K rows then V rows in one
// continuous stream that starts while warp 0 still reduces this token's q | k | v.  K7 issues a sweep of 8 loads per
// thread, waits, computes: at 200 cached keys its marginal bandwidth was 3.1 TB/s.
#ifndef B200W_SELF_RING
#define B200W_SELF_RING 8
#endif
#ifndef B200W_SELF_RING_CTAS
#define B200W_SELF_RING_CTAS 8
#endif
constexpr int kSelfRing = B200W_SELF_RING;  // a power of two
constexpr int kSelfMaxPages = 64;

__global__ void __launch_bounds__(kSelfThreads, B200W_SELF_RING_CTAS)
decoder_self_attention_ring_kernel(int n_head, const int* __restrict__ pos, __nv_bfloat16* k_pages, __nv_bfloat16* v_pages,
                                   const int* __restrict__ block_table, int max_pages, int page_size, __nv_bfloat16* __restrict__ out,
                                   const float* __restrict__ part, int n_split, long long split_stride, const float* __restrict__ bias,
                                   const int* __restrict__ finished) {
  __shared__ __align__(16) uint4 s_ring[kSelfRing][kSelfThreads];
  __shared__ float s_p[kMaxSelfKeys + 16];
  __shared__ float s_red[kSelfWarps];
  __shared__ float s_part[kSelfWarps][kHd];
  __shared__ uint4 s_qkv[3][8];  // this token's q, k, v of the head (bf16), 8 dims per entry
  __shared__ int s_bt[kSelfMaxPages];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int h = blockIdx.x % n_head, b = blockIdx.x / n_head;
  pdl_wait();
  pdl_launch_dependents();
  if (finished != nullptr && finished[b]) return;
  const int sub = lane & 7, kg = lane >> 3;
  const int d = n_head * kHd;
  const int p0 = pos[b], n_keys = p0 + 1;
  const int pshift = __ffs(page_size) - 1;
  for (int i = tid; i < max_pages; i += kSelfThreads) s_bt[i] = block_table[b * max_pages + i];
  __syncthreads();
  constexpr int kStep = kSelfWarps * 4;
  const int n_it = (n_keys + kStep - 1) / kStep, total = 2 * n_it;
  const int jt = warp * 4 + kg;
  const long long col = (long long)h * kHd + sub * 8;
  uint4* ring = &s_ring[0][tid];
  auto request = [&](int s) {
    if (s < total) {
      const bool is_v = s >= n_it;
      const int j = min(jt + (is_v ? s - n_it : s) * kStep, n_keys - 1);
      const long long row = (long long)s_bt[j >> pshift] * page_size + (j & (page_size - 1));
      cross_cp16(ring + (s & (kSelfRing - 1)) * kSelfThreads, (is_v ? v_pages : k_pages) + row * d + col);
    }
    cross_commit();
  };
#pragma unroll
  for (int s = 0; s < kSelfRing - 1; ++s) request(s);
  // this token's q | k | v: warp 0, lanes 0-7 reduce q, 8-15 k, 16-23 v; k / v are appended to the paged cache (the
  // stream's own read of that row is ignored: the row is taken from shared memory)
  if (warp == 0 && kg < 3) {
    float v8[8];
    const long long c3 = (long long)kg * d + col;
#pragma unroll
    for (int i = 0; i < 8; ++i) v8[i] = bias[c3 + i];
    for (int sidx = 0; sidx < n_split; ++sidx) {
      const float* pp = part + sidx * split_stride + (long long)b * 3 * d + c3;
      const float4 a0 = *reinterpret_cast<const float4*>(pp), a1 = *reinterpret_cast<const float4*>(pp + 4);
      v8[0] += a0.x; v8[1] += a0.y; v8[2] += a0.z; v8[3] += a0.w;
      v8[4] += a1.x; v8[5] += a1.y; v8[6] += a1.z; v8[7] += a1.w;
    }
    const uint4 packed = make_uint4(pack_bf16x2(v8[0], v8[1]), pack_bf16x2(v8[2], v8[3]), pack_bf16x2(v8[4], v8[5]),
                                    pack_bf16x2(v8[6], v8[7]));
    s_qkv[kg][sub] = packed;
    if (kg > 0) {
      const long long row = (long long)s_bt[p0 >> pshift] * page_size + (p0 & (page_size - 1));
      *reinterpret_cast<uint4*>((kg == 1 ? k_pages : v_pages) + row * d + col) = packed;
    }
  }
  __syncthreads();
  float qv[8];
  {
    float f[8];
    cross_unpack8(s_qkv[0][sub], f);
    const float c = 0.125f * kLog2e;
#pragma unroll
    for (int i = 0; i < 8; ++i) qv[i] = f[i] * c;
  }
  auto take = [&](int s, int j, int which) -> uint4 {
    cross_wait<kSelfRing - 2>();
    uint4 u = ring[(s & (kSelfRing - 1)) * kSelfThreads];
    request(s + kSelfRing - 1);
    if (j >= p0) u = s_qkv[which][sub];  // this token's row (and the clamped padding rows after it)
    return u;
  };
  float mx = -INFINITY;
  for (int i = 0; i < n_it; ++i) {
    const int j = jt + i * kStep;
    const uint4 u = take(i, j, 1);
    float f[8];
    cross_unpack8(u, f);
    float sc = f[0] * qv[0];
#pragma unroll
    for (int e = 1; e < 8; ++e) sc = fmaf(f[e], qv[e], sc);
    sc += __shfl_xor_sync(0xffffffffu, sc, 1);
    sc += __shfl_xor_sync(0xffffffffu, sc, 2);
    sc += __shfl_xor_sync(0xffffffffu, sc, 4);
    if (sub == 0) s_p[j] = sc;  // (padding rows repeat the last key: harmless for the maximum, zeroed below)
    mx = fmaxf(mx, sc);
  }
  mx = warp_max(mx);
  if (lane == 0) s_red[warp] = mx;
  __syncthreads();
  mx = s_red[0];
#pragma unroll
  for (int i = 1; i < kSelfWarps; ++i) mx = fmaxf(mx, s_red[i]);
  __syncthreads();
  float sum = 0.0f;
  for (int j = tid; j < n_keys; j += kSelfThreads) {
    const float p = fast_exp2(s_p[j] - mx);
    sum += p;
    s_p[j] = __bfloat162float(__float2bfloat16(p));  // bf16 probabilities, as in the tensor-core path
  }
  if (tid < n_it * kStep - n_keys) s_p[n_keys + tid] = 0.0f;
  sum = warp_sum(sum);
  if (lane == 0) s_red[warp] = sum;
  __syncthreads();
  sum = 0.0f;
#pragma unroll
  for (int i = 0; i < kSelfWarps; ++i) sum += s_red[i];
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.0f;
  for (int i = 0; i < n_it; ++i) {
    const int j = jt + i * kStep;
    const uint4 u = take(n_it + i, j, 2);
    const float p = s_p[j];
    float f[8];
    cross_unpack8(u, f);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = fmaf(p, f[e], acc[e]);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 8);
    acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 16);
  }
  if (kg == 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) s_part[warp][sub * 8 + i] = acc[i];
  }
  __syncthreads();
  if (tid < kHd) {
    float v = 0.0f;
#pragma unroll
    for (int w = 0; w < kSelfWarps; ++w) v += s_part[w][tid];
    out[(long long)b * d + h * kHd + tid] = __float2bfloat16(v / sum);
  }
}

// B200W_SELF_STREAM=0 keeps K7 for every shape (A/B; read on every call, see cross_ring_enabled)
static bool self_ring_enabled() {
  const char* e = getenv("B200W_SELF_STREAM");
  return !(e != nullptr && e[0] == '0');
}

int launch_decoder_self_attention(const __nv_bfloat16* qkv, int n_seq, int n_q, int n_head, const int* pos,
                                  __nv_bfloat16* k_pages, __nv_bfloat16* v_pages, const int* block_table,
                                  int max_pages_per_seq, int page_size, __nv_bfloat16* out, cudaStream_t stream,
                                  const float* part, int n_split, long long split_stride, const float* bias,
                                  const int* finished) {
  B200W_CHECK_ARG(n_split == 0 || (n_q == 1 && part && bias), "self_attention: split-K input needs n_q == 1");
  B200W_CHECK_ARG(n_split > 0 || qkv, "self_attention: null qkv");
  B200W_CHECK_ARG(n_seq > 0 && n_q > 0 && (long long)n_seq * n_q * n_head < (1ll << 30), "self_attention: bad sizes");
  B200W_CHECK_ARG(max_pages_per_seq * page_size <= kMaxSelfKeys, "self_attention: context above %d", kMaxSelfKeys);
  B200W_CHECK_ARG(page_size > 0 && (page_size & (page_size - 1)) == 0, "self_attention: page_size must be a power of two");
  const int units = n_seq * n_head * n_q;
  ProfScope prof_("decoder_self_attention", stream);
  if (n_split > 0 && n_q == 1 && max_pages_per_seq <= kSelfMaxPages && self_ring_enabled()) {
    B200W_CUDA_OK(launch_k(decoder_self_attention_ring_kernel, dim3(units), dim3(kSelfThreads), 0, stream, n_head, pos, k_pages, v_pages,
                           block_table, max_pages_per_seq, page_size, out, part, n_split, split_stride, bias, finished));
    count_launch();
    return kOk;
  }
  B200W_CUDA_OK(launch_k(decoder_self_attention_kernel, dim3(units), dim3(kSelfThreads), 0, stream,
                         qkv, n_seq, n_q, n_head, pos, k_pages, v_pages, block_table, max_pages_per_seq, page_size, out,
                         part, n_split, split_stride, bias, finished));
  count_launch();
  return kOk;
}

// =============================================================================================== K8
constexpr int kCrossThreads = 256;
constexpr int kCrossWarps = kCrossThreads / 32;
constexpr int kMaxCrossKeys = 1536;
#ifndef B200W_CROSS_UNROLL
#define B200W_CROSS_UNROLL 8
#endif
// 16-byte loads in flight per thread.  The kernel is bound by bytes in flight, not by DRAM: measured on B200 at the
// bench shape (CUDA events, incl. ~5 us event overhead) 4 -> 186 us, 6 -> 184, 8 -> 159, 12 -> 177, 16 -> 204
// (registers 48 / 56 / 64 / 96 / 128 cut the resident CTAs); prefetching V across the softmax barriers: 176.
constexpr int kCrossUnroll = B200W_CROSS_UNROLL;

// kProbs: also store the attention probabilities (f32, [b][qi][h][T]) -- the cross-attention weights the word-level
// alignment reads (UPSTREAM timing.py: softmax of the captured QK).  A separate instantiation: the decode-step
// kernel is sensitive to every extra instruction in its prologue (section on kCrossUnroll above).
// kSplit (small batches, e.g. the batch-1 steps of the exact sequential mode): n_seq x n_head CTAs cannot fill the GPU,
// so the keys of one (sequence, head) are cut into kv_splits chunks handled by different CTAs; each writes
// (max, sum, unnormalised output) and the last one to arrive merges the chunks (self-resetting arrival counter).
constexpr int kCrossPartFloats = 2 + kHd;

template <bool kProbs, bool kSplit>
__global__ void __launch_bounds__(kCrossThreads)
decoder_cross_attention_kernel(const __nv_bfloat16* __restrict__ q, int n_q, int n_head,
                               const __nv_bfloat16* __restrict__ cross_kv, long long seq_stride, int T_all,
                               const int* __restrict__ slot, __nv_bfloat16* __restrict__ out,
                               const float* __restrict__ part, int n_split, long long split_stride,
                               const float* __restrict__ bias, const int* __restrict__ finished,
                               float* __restrict__ probs_out, int kv_splits, float* __restrict__ kv_part,
                               int* __restrict__ kv_cnt) {
  __shared__ float s_p[kMaxCrossKeys];
  __shared__ float s_red[kCrossWarps];
  __shared__ float s_part[kCrossWarps][kHd];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int sub = lane & 7;   // which 8-wide slice of the head dim
  const int kg = lane >> 3;   // which of the warp's 4 concurrent keys
  const int d = n_head * kHd;
  const long long ld = 2ll * d;  // K | V interleaved per row
  // queries of one (sequence, head) are adjacent CTAs: they stream the same K/V rows and share them through L2
  const int qi = blockIdx.x, b = blockIdx.z;
  const int h = kSplit ? (int)blockIdx.y / kv_splits : (int)blockIdx.y;
  const int chunk = kSplit ? (int)blockIdx.y - h * kv_splits : 0;
  pdl_wait();
  pdl_launch_dependents();  // after the wait: at most one dependent grid is resident ahead of the running one
  if (finished != nullptr && finished[b]) return;  // no K/V streaming for sequences that have emitted EOT
  // key range of this CTA: everything, or chunk `chunk` of kv_splits (multiples of 32 keys)
  int k0 = 0, T = T_all;
  if constexpr (kSplit) {
    const int per = ((T_all + kv_splits - 1) / kv_splits + 31) & ~31;
    k0 = min(chunk * per, T_all);
    T = min(per, T_all - k0);
  }
  const __nv_bfloat16* kbase = cross_kv + (long long)slot[b] * seq_stride + (long long)k0 * ld + h * kHd + sub * 8;
  const __nv_bfloat16* vbase = kbase + d;

  float qv[8];
  if (n_split > 0) {
    // q arrives as split-K fp32 partial slabs of the query projection (+ bias): reduce, round to bf16
    const long long col = ((long long)b * n_q + qi) * d + h * kHd + sub * 8;
    const float c = 0.125f * kLog2e;
#pragma unroll
    for (int i = 0; i < 8; ++i) qv[i] = bias[h * kHd + sub * 8 + i];
    for (int sidx = 0; sidx < n_split; ++sidx) {
      const float4 a = *reinterpret_cast<const float4*>(part + sidx * split_stride + col);
      const float4 e = *reinterpret_cast<const float4*>(part + sidx * split_stride + col + 4);
      qv[0] += a.x; qv[1] += a.y; qv[2] += a.z; qv[3] += a.w;
      qv[4] += e.x; qv[5] += e.y; qv[6] += e.z; qv[7] += e.w;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) qv[i] = __bfloat162float(__float2bfloat16(qv[i])) * c;
  } else {
    const uint4 u = *reinterpret_cast<const uint4*>(q + ((long long)b * n_q + qi) * d + h * kHd + sub * 8);
    const float2 a0 = unpack_bf16x2(u.x), a1 = unpack_bf16x2(u.y), a2 = unpack_bf16x2(u.z), a3 = unpack_bf16x2(u.w);
    const float c = 0.125f * kLog2e;
    qv[0] = a0.x * c; qv[1] = a0.y * c; qv[2] = a1.x * c; qv[3] = a1.y * c;
    qv[4] = a2.x * c; qv[5] = a2.y * c; qv[6] = a3.x * c; qv[7] = a3.y * c;
  }

  // ---- scores: 4 keys per warp per load instruction, kCrossUnroll loads in flight per thread ----
  constexpr int kStep = kCrossWarps * 4;  // keys per CTA sweep
  float mx = -INFINITY;
  for (int j0 = warp * 4; j0 < T; j0 += kCrossUnroll * kStep) {  // warp-uniform trip count: the shuffles need every lane
    uint4 u[kCrossUnroll];
#pragma unroll
    for (int i = 0; i < kCrossUnroll; ++i) {
      const int j = j0 + kg + i * kStep;
      u[i] = (j < T) ? __ldg(reinterpret_cast<const uint4*>(kbase + j * ld)) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int i = 0; i < kCrossUnroll; ++i) {
      const int j = j0 + kg + i * kStep;
      const float2 a0 = unpack_bf16x2(u[i].x), a1 = unpack_bf16x2(u[i].y), a2 = unpack_bf16x2(u[i].z),
                   a3 = unpack_bf16x2(u[i].w);
      float s = a0.x * qv[0];
      s = fmaf(a0.y, qv[1], s);
      s = fmaf(a1.x, qv[2], s);
      s = fmaf(a1.y, qv[3], s);
      s = fmaf(a2.x, qv[4], s);
      s = fmaf(a2.y, qv[5], s);
      s = fmaf(a3.x, qv[6], s);
      s = fmaf(a3.y, qv[7], s);
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      s += __shfl_xor_sync(0xffffffffu, s, 4);
      if (j < T) {
        if (sub == 0) s_p[j] = s;
        mx = fmaxf(mx, s);
      }
    }
  }
  mx = warp_max(mx);
  if (lane == 0) s_red[warp] = mx;
  __syncthreads();
  mx = s_red[0];
#pragma unroll
  for (int i = 1; i < kCrossWarps; ++i) mx = fmaxf(mx, s_red[i]);
  __syncthreads();
  float sum = 0.0f;
  for (int j = tid; j < T; j += kCrossThreads) {
    const float p = fast_exp2(s_p[j] - mx);
    sum += p;
    s_p[j] = __bfloat162float(__float2bfloat16(p));  // bf16 probabilities, as in the tensor-core path
  }
  sum = warp_sum(sum);
  if (lane == 0) s_red[warp] = sum;
  __syncthreads();
  sum = 0.0f;
#pragma unroll
  for (int i = 0; i < kCrossWarps; ++i) sum += s_red[i];
  if constexpr (kProbs) {
    // each thread re-reads the entries it wrote itself (same j = tid + k * kCrossThreads mapping): no barrier needed
    float* po = probs_out + (((long long)b * n_q + qi) * n_head + h) * T;
    const float inv = 1.0f / sum;
    for (int j = tid; j < T; j += kCrossThreads) po[j] = s_p[j] * inv;
  }

  // ---- output ----
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.0f;
  for (int j0 = warp * 4; j0 < T; j0 += kCrossUnroll * kStep) {
    uint4 u[kCrossUnroll];
    float p[kCrossUnroll];
#pragma unroll
    for (int i = 0; i < kCrossUnroll; ++i) {
      const int j = j0 + kg + i * kStep;
      u[i] = (j < T) ? __ldg(reinterpret_cast<const uint4*>(vbase + j * ld)) : make_uint4(0, 0, 0, 0);
      p[i] = (j < T) ? s_p[j] : 0.0f;
    }
#pragma unroll
    for (int i = 0; i < kCrossUnroll; ++i) {
      const float2 a0 = unpack_bf16x2(u[i].x), a1 = unpack_bf16x2(u[i].y), a2 = unpack_bf16x2(u[i].z),
                   a3 = unpack_bf16x2(u[i].w);
      acc[0] = fmaf(p[i], a0.x, acc[0]);
      acc[1] = fmaf(p[i], a0.y, acc[1]);
      acc[2] = fmaf(p[i], a1.x, acc[2]);
      acc[3] = fmaf(p[i], a1.y, acc[3]);
      acc[4] = fmaf(p[i], a2.x, acc[4]);
      acc[5] = fmaf(p[i], a2.y, acc[5]);
      acc[6] = fmaf(p[i], a3.x, acc[6]);
      acc[7] = fmaf(p[i], a3.y, acc[7]);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 8);
    acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 16);
  }
  if (kg == 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) s_part[warp][sub * 8 + i] = acc[i];
  }
  __syncthreads();
  if constexpr (!kSplit) {
    if (tid < kHd) {
      float v = 0.0f;
#pragma unroll
      for (int w = 0; w < kCrossWarps; ++w) v += s_part[w][tid];
      out[((long long)b * n_q + qi) * d + h * kHd + tid] = __float2bfloat16(v / sum);
    }
  } else {
    __shared__ int s_last;
    const int unit = (b * n_q + qi) * n_head + h;
    float* mine = kv_part + ((long long)unit * kv_splits + chunk) * kCrossPartFloats;
    if (tid < kHd) {
      float v = 0.0f;
#pragma unroll
      for (int w = 0; w < kCrossWarps; ++w) v += s_part[w][tid];
      mine[2 + tid] = v;
    }
    if (tid == 0) {
      mine[0] = (T > 0) ? mx : -INFINITY;  // an empty chunk contributes nothing
      mine[1] = (T > 0) ? sum : 0.0f;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(kv_cnt + unit, 1) == kv_splits - 1) ? 1 : 0;
    __syncthreads();
    if (s_last) {  // every chunk of this (sequence, head, query) is in memory: merge
      __threadfence();
      if (tid < kHd) {
        const float* all = kv_part + (long long)unit * kv_splits * kCrossPartFloats;
        float M = -INFINITY;
        for (int c2 = 0; c2 < kv_splits; ++c2) M = fmaxf(M, __ldcg(all + c2 * kCrossPartFloats));
        float L = 0.0f, o = 0.0f;
        for (int c2 = 0; c2 < kv_splits; ++c2) {
          const float w2 = fast_exp2(__ldcg(all + c2 * kCrossPartFloats) - M);
          L = fmaf(__ldcg(all + c2 * kCrossPartFloats + 1), w2, L);
          o = fmaf(__ldcg(all + c2 * kCrossPartFloats + 2 + tid), w2, o);
        }
        out[((long long)b * n_q + qi) * d + h * kHd + tid] = __float2bfloat16(o / L);
      }
      if (tid == 0) kv_cnt[unit] = 0;  // ready for the next launch
    }
  }
}

// K8r: the decode-step form for full batches (n_q == 1, one CTA per (sequence, head), no key split, no probability
// output).  Same arithmetic as K8 above, but the K / V rows do not pass through registers on their way in: every
// thread owns kRing 16-byte slots of shared memory and keeps them filled with cp.async (LDGSTS), consuming slot i while
// slots i + 1 .. i + kRing - 1 are in flight.  K8 issues 8 loads, waits for all of them, computes, and only then issues
// the next 8 -- its bytes in flight swing between 32 KB and 0 per CTA and it stops at 0.95 of the COPY bandwidth
// (6.3 TB/s), while a read-only stream with >= 64 KB per SM continuously in flight reaches 7.3 TB/s on this part
// (tools/probes/probe_read.cu).  Here the stream is continuous from the first instruction of the CTA (the loads are
// in flight while the query is reduced from its split-K slabs) across the K -> V transition (V rows are requested
// before the softmax barriers) to the end.  A thread reads back only what it copied itself: cp.async.wait_group is
// all the synchronisation the ring needs.
#ifndef B200W_CROSS_RING
#define B200W_CROSS_RING 8
#endif
constexpr int kRing = B200W_CROSS_RING;  // a power of two: slot = position & (kRing - 1)
static_assert((kRing & (kRing - 1)) == 0 && kRing >= 4, "ring size");
constexpr int kCrossRingSmem = kRing * kCrossThreads * 16;
#ifndef B200W_CROSS_RING_CTAS
#define B200W_CROSS_RING_CTAS 4
#endif

// The kernel issues 58 -> 40 warp-instructions per 512 bytes in this form (ncu r02: the first ring version ran at
// 78 % issue-active -- co-limited by instruction issue, not only by HBM): the source pointer advances by a constant,
// the K -> V switch and the clamp of the last (partial) iteration live in peeled iterations instead of selects in the
// loop, ring slots are position & (kRing - 1), scores of the padding keys (duplicates of key T - 1, so the maximum is
// unaffected) are stored unconditionally and their probabilities zeroed once, bf16 pairs unpack with a shift and a mask.
__global__ void __launch_bounds__(kCrossThreads, B200W_CROSS_RING_CTAS)
decoder_cross_attention_ring_kernel(int n_head, const __nv_bfloat16* __restrict__ cross_kv, long long seq_stride, int T,
                                    const int* __restrict__ slot, __nv_bfloat16* __restrict__ out,
                                    const __nv_bfloat16* __restrict__ q, const float* __restrict__ part, int n_split,
                                    long long split_stride, const float* __restrict__ bias, const int* __restrict__ finished) {
  extern __shared__ __align__(16) unsigned char cross_ring_raw[];
  uint4* ring = reinterpret_cast<uint4*>(cross_ring_raw) + threadIdx.x;  // [kRing][kCrossThreads], this thread's column
  __shared__ float s_p[kMaxCrossKeys];
  __shared__ float s_red[kCrossWarps];
  __shared__ float s_part[kCrossWarps][kHd];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int sub = lane & 7, kg = lane >> 3;
  const int d = n_head * kHd;
#ifdef B200W_CROSS_HEADMAJOR_PROBE  // timing probe only: (slot, head, K | V, key, 64) layout -- a unit's rows are contiguous
  const long long ld = kHd;
#else
  const long long ld = 2ll * d;
#endif
  const int h = blockIdx.y, b = blockIdx.z;
  pdl_wait();
  pdl_launch_dependents();
  if (finished != nullptr && finished[b]) return;
  constexpr int kStep = kCrossWarps * 4;     // keys per CTA per iteration
  const int n_it = (T + kStep - 1) / kStep;  // iterations of the K stream (>= kRing: checked by the launcher); V follows
  const int jt = warp * 4 + kg;              // this thread's key within an iteration
  // byte pointers: this thread's 16 bytes of row jt (K half); rows of later iterations are `stepb` apart, V is `k2v` on
#ifdef B200W_CROSS_HEADMAJOR_PROBE
  const unsigned char* kp = reinterpret_cast<const unsigned char*>(cross_kv + (long long)slot[b] * seq_stride + (long long)h * 2 * T * kHd +
                                                                   (long long)jt * ld + sub * 8);
  const long long stepb = (long long)kStep * ld * 2, k2v = (long long)T * kHd * 2;
#else
  const unsigned char* kp = reinterpret_cast<const unsigned char*>(cross_kv + (long long)slot[b] * seq_stride + (long long)jt * ld +
                                                                   h * kHd + sub * 8);
  const long long stepb = (long long)kStep * ld * 2, k2v = (long long)d * 2;
#endif
  const int j_last = jt + (n_it - 1) * kStep;  // row of the last iteration: past the end for some warps when T % 32 != 0
  const long long fix = j_last >= T ? (long long)(T - 1 - j_last) * ld * 2 : 0;  // ... those re-read row T - 1 instead
  auto request = [&](const unsigned char* src, int pos) {
    cross_cp16(ring + (pos & (kRing - 1)) * kCrossThreads, src);
    cross_commit();
  };
  auto take = [&](int pos) -> uint4 {  // (before this iteration's request: kRing - 1 + pos groups are committed)
    cross_wait<kRing - 2>();
    return ring[(pos & (kRing - 1)) * kCrossThreads];
  };
  const unsigned char* rq = kp;
#pragma unroll
  for (int s = 0; s < kRing - 1; ++s) {
    request(rq, s);
    rq += stepb;
  }

  float qv[8];
  if (n_split > 0) {
    // q arrives as split-K fp32 partial slabs of the query projection (+ bias): reduce, round to bf16
    const long long col = (long long)b * d + h * kHd + sub * 8;
    const float c = 0.125f * kLog2e;
#pragma unroll
    for (int i = 0; i < 8; ++i) qv[i] = bias[h * kHd + sub * 8 + i];
    for (int sidx = 0; sidx < n_split; ++sidx) {
      const float4 a = *reinterpret_cast<const float4*>(part + sidx * split_stride + col);
      const float4 e = *reinterpret_cast<const float4*>(part + sidx * split_stride + col + 4);
      qv[0] += a.x; qv[1] += a.y; qv[2] += a.z; qv[3] += a.w;
      qv[4] += e.x; qv[5] += e.y; qv[6] += e.z; qv[7] += e.w;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) qv[i] = __bfloat162float(__float2bfloat16(qv[i])) * c;
  } else {
    float f[8];
    cross_unpack8(*reinterpret_cast<const uint4*>(q + (long long)b * d + h * kHd + sub * 8), f);
    const float c = 0.125f * kLog2e;
#pragma unroll
    for (int i = 0; i < 8; ++i) qv[i] = f[i] * c;
  }

  // ---- scores ----
  float mx = -INFINITY;
  float* sp = s_p + jt;  // (sub == 0 lanes write)
  const bool writer = sub == 0;
  auto score = [&](const uint4& u, int i) {
    float f[8];
    cross_unpack8(u, f);
    float s = f[0] * qv[0];
#pragma unroll
    for (int e = 1; e < 8; ++e) s = fmaf(f[e], qv[e], s);
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    if (writer) sp[i * kStep] = s;  // rows past T are copies of row T - 1: harmless for the maximum, zeroed below
    mx = fmaxf(mx, s);
  };
  int i = 0;
  for (; i < n_it - kRing; ++i) {  // requests: K rows up to the second to last
    const uint4 u = take(i);
    request(rq, i + kRing - 1);
    rq += stepb;
    score(u, i);
  }
  {  // request: the last K row (clamped); from here on V
    const uint4 u = take(i);
    request(rq + fix, i + kRing - 1);
    rq = kp + k2v;
    score(u, i);
    ++i;
  }
  for (; i < n_it; ++i) {  // requests: the first kRing - 1 V rows, in flight across the softmax barriers
    const uint4 u = take(i);
    request(rq, i + kRing - 1);
    rq += stepb;
    score(u, i);
  }
  mx = warp_max(mx);
  if (lane == 0) s_red[warp] = mx;
  __syncthreads();
  mx = s_red[0];
#pragma unroll
  for (int w = 1; w < kCrossWarps; ++w) mx = fmaxf(mx, s_red[w]);
  __syncthreads();
  float sum = 0.0f;
  for (int j = tid; j < T; j += kCrossThreads) {
    const float p = fast_exp2(s_p[j] - mx);
    sum += p;
    s_p[j] = __bfloat162float(__float2bfloat16(p));  // bf16 probabilities, as in the tensor-core path
  }
  if (tid < n_it * kStep - T) s_p[T + tid] = 0.0f;  // padding keys carry no weight
  sum = warp_sum(sum);
  if (lane == 0) s_red[warp] = sum;
  __syncthreads();
  sum = 0.0f;
#pragma unroll
  for (int w = 0; w < kCrossWarps; ++w) sum += s_red[w];

  // ---- output ----
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.0f;
  auto accumulate = [&](const uint4& u, int k) {
    const float p = sp[k * kStep];
    float f[8];
    cross_unpack8(u, f);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = fmaf(p, f[e], acc[e]);
  };
  int k = 0;
  for (; k < n_it - kRing; ++k) {
    const uint4 u = take(n_it + k);
    request(rq, n_it + k + kRing - 1);
    rq += stepb;
    accumulate(u, k);
  }
  {
    const uint4 u = take(n_it + k);
    request(rq + fix, n_it + k + kRing - 1);  // the last V row
    accumulate(u, k);
    ++k;
  }
  for (; k < n_it; ++k) {
    const uint4 u = take(n_it + k);
    cross_commit();  // (an empty group keeps the wait count uniform)
    accumulate(u, k);
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 8);
    acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 16);
  }
  if (kg == 0) {
#pragma unroll
    for (int e = 0; e < 8; ++e) s_part[warp][sub * 8 + e] = acc[e];
  }
  __syncthreads();
  if (tid < kHd) {
    float v = 0.0f;
#pragma unroll
    for (int w = 0; w < kCrossWarps; ++w) v += s_part[w][tid];
    out[(long long)b * d + h * kHd + tid] = __float2bfloat16(v / sum);
  }
}

// K8p: the form for two half-batches decoded on two streams (B200W_DECODE_STREAMS=2): ONE persistent CTA per SM that
// leaves room for a co-resident decode-chain CTA of the other half-batch (K11: 43 K registers).  512 threads at <= 40
// registers, a ring of kRingP x 16 B per thread (64 KB in flight per SM), (sequence, head) units dealt round-robin and
// walked back to back: the request stream runs ahead across the K -> V switch AND across unit boundaries, so the SM's
// bytes in flight never drain.  Queries of all of a CTA's units are reduced once, at the start of the launch.
constexpr int kCrossPThreads = 512;
constexpr int kCrossPWarps = kCrossPThreads / 32;
constexpr int kRingP = 8;
constexpr int kCrossPSmem = kRingP * kCrossPThreads * 16;
constexpr int kCrossPMaxUnits = 24;  // units per CTA (round-robin): n_seq * n_head <= 24 * grid

__global__ void __launch_bounds__(kCrossPThreads, 3)
decoder_cross_attention_persistent_kernel(int n_head, int n_units, const __nv_bfloat16* __restrict__ cross_kv, long long seq_stride,
                                          int T, const int* __restrict__ slot, __nv_bfloat16* __restrict__ out,
                                          const __nv_bfloat16* __restrict__ q, const float* __restrict__ part, int n_split,
                                          long long split_stride, const float* __restrict__ bias, const int* __restrict__ finished) {
  extern __shared__ __align__(16) unsigned char cross_ring_raw[];
  uint4* ring = reinterpret_cast<uint4*>(cross_ring_raw) + threadIdx.x;
  __shared__ float s_p[kMaxCrossKeys];
  __shared__ float s_red[kCrossPWarps];
  __shared__ float s_part[kCrossPWarps][kHd];
  __shared__ __align__(16) float s_q[kCrossPMaxUnits][kHd];  // scaled queries of this CTA's units
  __shared__ long long s_base[kCrossPMaxUnits + 1];          // byte offset of a unit's first K row (this CTA's list)
  __shared__ int s_unit[kCrossPMaxUnits];
  __shared__ int s_n;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int sub = lane & 7, kg = lane >> 3;
  const int d = n_head * kHd;
  const long long ld = 2ll * d;
  pdl_wait();
  pdl_launch_dependents();
  // this CTA's units (sequences that have emitted EOT are skipped)
  if (tid == 0) {
    int n = 0;
    for (int u = blockIdx.x; u < n_units && n < kCrossPMaxUnits; u += gridDim.x) {
      const int b = u / n_head, h = u - b * n_head;
      if (finished != nullptr && finished[b]) continue;
      s_unit[n] = u;
      s_base[n] = ((long long)slot[b] * seq_stride + h * kHd) * 2;
      ++n;
    }
    s_base[n] = n > 0 ? s_base[n - 1] : 0;  // (the request stream may look one unit past the end; never dereferenced)
    s_n = n;
  }
  __syncthreads();
  const int n_mine = s_n;
  if (n_mine == 0) return;
  constexpr int kStep = kCrossPWarps * 4;
  const int n_it = (T + kStep - 1) / kStep;  // >= kRingP (checked by the launcher)
  const int jt = warp * 4 + kg;
  const unsigned char* kv0 = reinterpret_cast<const unsigned char*>(cross_kv) + ((long long)jt * ld + sub * 8) * 2;
  const long long stepb = (long long)kStep * ld * 2, k2v = (long long)d * 2;
  const int j_last = jt + (n_it - 1) * kStep;
  const long long fix = j_last >= T ? (long long)(T - 1 - j_last) * ld * 2 : 0;
  // request stream: unit ru, pass (0 = K, 1 = V), iteration ri
  int ru = 0, ri = 0, rpass = 0;
  const unsigned char* rq = kv0 + s_base[0];
  int rpos = 0;  // ring position of the next request (== number of requests issued)
  auto request = [&]() {
    if (ru < n_mine) {
      cross_cp16(ring + (rpos & (kRingP - 1)) * kCrossPThreads, ri == n_it - 1 ? rq + fix : rq);
      rq += stepb;
      if (++ri == n_it) {
        ri = 0;
        if (rpass == 0) {
          rpass = 1;
          rq = kv0 + s_base[ru] + k2v;
        } else {
          rpass = 0;
          ++ru;
          rq = kv0 + s_base[ru];
        }
      }
    }
    cross_commit();
    ++rpos;
  };
#pragma unroll
  for (int s = 0; s < kRingP - 1; ++s) request();
  // queries of all units of this CTA: one (unit, dim) per thread and round
  for (int i = tid; i < n_mine * kHd; i += kCrossPThreads) {
    const int k = i >> 6, e = i & 63, u = s_unit[k];
    const int b = u / n_head, h = u - b * n_head;
    float v;
    if (n_split > 0) {
      v = bias[h * kHd + e];
      for (int sidx = 0; sidx < n_split; ++sidx) v += part[sidx * split_stride + (long long)b * d + h * kHd + e];
      v = __bfloat162float(__float2bfloat16(v));
    } else {
      v = __bfloat162float(q[(long long)b * d + h * kHd + e]);
    }
    s_q[k][e] = v * (0.125f * kLog2e);
  }
  __syncthreads();

  int cpos = 0;  // ring position of the next row to consume
  auto take = [&]() -> uint4 {
    cross_wait<kRingP - 2>();
    const uint4 u = ring[(cpos & (kRingP - 1)) * kCrossPThreads];
    ++cpos;
    return u;
  };
  float* sp = s_p + jt;
  const bool writer = sub == 0;
  for (int k = 0; k < n_mine; ++k) {
    float qv[8];
    {
      const float4 a = *reinterpret_cast<const float4*>(&s_q[k][sub * 8]), e = *reinterpret_cast<const float4*>(&s_q[k][sub * 8 + 4]);
      qv[0] = a.x; qv[1] = a.y; qv[2] = a.z; qv[3] = a.w; qv[4] = e.x; qv[5] = e.y; qv[6] = e.z; qv[7] = e.w;
    }
    float mx = -INFINITY;
    for (int i = 0; i < n_it; ++i) {
      const uint4 u = take();
      request();
      float f[8];
      cross_unpack8(u, f);
      float s = f[0] * qv[0];
#pragma unroll
      for (int e = 1; e < 8; ++e) s = fmaf(f[e], qv[e], s);
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      s += __shfl_xor_sync(0xffffffffu, s, 4);
      if (writer) sp[i * kStep] = s;
      mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    if (lane == 0) s_red[warp] = mx;
    __syncthreads();
    mx = s_red[0];
#pragma unroll
    for (int w = 1; w < kCrossPWarps; ++w) mx = fmaxf(mx, s_red[w]);
    __syncthreads();
    float sum = 0.0f;
    for (int j = tid; j < T; j += kCrossPThreads) {
      const float p = fast_exp2(s_p[j] - mx);
      sum += p;
      s_p[j] = __bfloat162float(__float2bfloat16(p));
    }
    if (tid < n_it * kStep - T) s_p[T + tid] = 0.0f;
    sum = warp_sum(sum);
    if (lane == 0) s_red[warp] = sum;
    __syncthreads();
    sum = 0.0f;
#pragma unroll
    for (int w = 0; w < kCrossPWarps; ++w) sum += s_red[w];
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.0f;
    for (int i = 0; i < n_it; ++i) {
      const uint4 u = take();
      request();
      const float p = sp[i * kStep];
      float f[8];
      cross_unpack8(u, f);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = fmaf(p, f[e], acc[e]);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 8);
      acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 16);
    }
    if (kg == 0) {
#pragma unroll
      for (int e = 0; e < 8; ++e) s_part[warp][sub * 8 + e] = acc[e];
    }
    __syncthreads();  // every warp has left the V pass: s_p may be rewritten by the next unit's scores
    if (tid < kHd) {
      float v = 0.0f;
#pragma unroll
      for (int w = 0; w < kCrossPWarps; ++w) v += s_part[w][tid];
      const int u = s_unit[k];
      out[(long long)(u / n_head) * d + (u % n_head) * kHd + tid] = __float2bfloat16(v / sum);
    }
  }
}

// B200W_CROSS_PERSIST=1: the persistent one-CTA-per-SM form (K8p) for full batches (read on every call, see cross_ring_enabled)
static bool cross_persist_enabled() {
  const char* e = getenv("B200W_CROSS_PERSIST");
  if (!(e != nullptr && e[0] == '1')) return false;
  // the same shared-memory carve-out as the chain kernel it is meant to share SMs with (an SM is drained before its
  // carve-out changes): this kernel, and the self-attention kernels that run beside it on the other stream
  static const bool ready = [] {
    if (cudaFuncSetAttribute(decoder_cross_attention_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kCrossPSmem) != cudaSuccess)
      return false;
    cudaFuncSetAttribute(decoder_cross_attention_persistent_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(decoder_self_attention_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(decoder_self_attention_ring_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    return true;
  }();
  return ready;
}

// B200W_CROSS_STREAM=0 keeps the register-staged K8 for every shape (A/B; read on every call: the parity tests switch
// between the two forms inside one process)
static bool cross_ring_enabled() {
  static const bool ready =
      cudaFuncSetAttribute(decoder_cross_attention_ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kCrossRingSmem) == cudaSuccess;
  const char* e = getenv("B200W_CROSS_STREAM");
  return ready && !(e != nullptr && e[0] == '0');
}

int cross_attention_kv_splits(int n_seq, int n_q, int n_head) {
  const long long units = (long long)n_seq * n_q * n_head;
  const int s = (int)(device_sm_count() / (units > 0 ? units : 1));
  return s >= 2 ? (s > 8 ? 8 : s) : 1;  // split only when the (sequence, head) CTAs leave most SMs idle
}

int launch_decoder_cross_attention(const __nv_bfloat16* q, int n_seq, int n_q, int n_head,
                                   const __nv_bfloat16* cross_kv, long long seq_stride, int T, const int* slot,
                                   __nv_bfloat16* out, cudaStream_t stream, const float* part, int n_split,
                                   long long split_stride, const float* bias, const int* finished, float* probs_out,
                                   float* kv_part, int* kv_cnt) {
  B200W_CHECK_ARG(n_seq > 0 && n_seq <= 65535 && n_q > 0 && n_q <= 65535, "cross_attention: bad sizes");
  B200W_CHECK_ARG(n_split > 0 ? (part && bias) : (q != nullptr), "cross_attention: missing query input");
  B200W_CHECK_ARG(T > 0 && T <= kMaxCrossKeys, "cross_attention: T above %d", kMaxCrossKeys);
  const int kv_splits = (kv_part != nullptr && kv_cnt != nullptr && probs_out == nullptr) ? cross_attention_kv_splits(n_seq, n_q, n_head) : 1;
  ProfScope prof_("decoder_cross_attention", stream);
  if (kv_splits == 1 && probs_out == nullptr && n_q == 1 && T >= kRingP * kCrossPWarps * 4 && cross_persist_enabled() &&
      n_seq * n_head <= kCrossPMaxUnits * device_sm_count()) {
    B200W_CUDA_OK(launch_k(decoder_cross_attention_persistent_kernel, dim3(std::min(device_sm_count(), n_seq * n_head)),
                           dim3(kCrossPThreads), (size_t)kCrossPSmem, stream, n_head, n_seq * n_head, cross_kv, seq_stride, T, slot, out,
                           q, part, n_split, split_stride, bias, finished));
    count_launch();
    return kOk;
  }
  if (kv_splits == 1 && probs_out == nullptr && n_q == 1 && T >= kRing * kCrossWarps * 4 && cross_ring_enabled()) {
    B200W_CUDA_OK(launch_k(decoder_cross_attention_ring_kernel, dim3(1, n_head, n_seq), dim3(kCrossThreads), (size_t)kCrossRingSmem,
                           stream, n_head, cross_kv, seq_stride, T, slot, out, q, part, n_split, split_stride, bias, finished));
    count_launch();
    return kOk;
  }
  if (kv_splits > 1) {
    dim3 grid(n_q, n_head * kv_splits, n_seq);
    B200W_CUDA_OK(launch_k(decoder_cross_attention_kernel<false, true>, grid, dim3(kCrossThreads), 0, stream, q, n_q, n_head,
                           cross_kv, seq_stride, T, slot, out, part, n_split, split_stride, bias, finished,
                           static_cast<float*>(nullptr), kv_splits, kv_part, kv_cnt));
  } else {
    dim3 grid(n_q, n_head, n_seq);
    if (probs_out != nullptr)
      B200W_CUDA_OK(launch_k(decoder_cross_attention_kernel<true, false>, grid, dim3(kCrossThreads), 0, stream, q, n_q, n_head,
                             cross_kv, seq_stride, T, slot, out, part, n_split, split_stride, bias, finished, probs_out, 1,
                             static_cast<float*>(nullptr), static_cast<int*>(nullptr)));
    else
      B200W_CUDA_OK(launch_k(decoder_cross_attention_kernel<false, false>, grid, dim3(kCrossThreads), 0, stream, q, n_q, n_head,
                             cross_kv, seq_stride, T, slot, out, part, n_split, split_stride, bias, finished,
                             static_cast<float*>(nullptr), 1, static_cast<float*>(nullptr), static_cast<int*>(nullptr)));
  }
  count_launch();
  return kOk;
}

}  // namespace b200w
