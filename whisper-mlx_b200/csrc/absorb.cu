// K14: decoder cross-attention of single-token steps in ABSORBED form, on tcgen05.
//
// mlx_whisper's MultiHeadAttention (whisper.py, reached from /root/reference/run:3-6) caches
// K = xa Wk^T and V = xa Wv^T + bv per decoder layer and every decode step reads both: 2 x 1500 x d bf16 per window and
// layer (245.8 MB per window and step for large-v3) -- the HBM floor of batched decoding (K8 runs at 0.95 of the HBM peak
// and still takes 64 % of the step).  The encoder states xa are the SAME for all layers and serve both sides:
//     scores  s[h][t] = q_h . (Wk_h xa_t)            = (Wk_h^T q_h) . xa_t         = qa_h . xa_t
//     output  o_h     = sum_t p[h][t] (Wv_h xa_t + bv_h) = Wv_h (sum_t p[h][t] xa_t) + bv_h   (sum_t p = 1)
// so a step only has to stream xa (1500 x d bf16 per window and layer: HALF the bytes) if it spends 2 x 20 x more
// multiply-adds -- which the tensor cores have idle during decoding.  Three kernels per layer:
//   K14a absorb_q_kernel   qa[b][h][:] = Wk_h^T q[b][h]          (q from the split-K slabs of the query projection)
//   K14b absorb_attn_kernel  flash-style over 64-key tiles of xa: S^T = xa_tile qa^T (tcgen05, M = 64 keys, N = 24 heads),
//        online softmax per head across the key lanes, O'^T += xa_tile^T P^T (M = 128 features, the SAME shared-memory
//        tile read MN-major), partial (max, sum, O') per contiguous tile range, merged by the last CTA of a window
//   K14c absorb_v_kernel   att[b][h*64 + j] = Wv_h[j] . O'[b][h] + bv
// The xa tiles are loaded once by TMA (64 keys x d features = 160 KB for d = 1280) and used by both MMAs.
#include "common.cuh"
#include "kernels.h"

namespace b200w {

constexpr float kAbLog2e = 1.4426950408889634f;
constexpr int kAbKeys = 64;                       // keys per tile
constexpr int kAbQRows = 24;                      // rows of qa per sequence (heads, zero padded): N of the score MMA
constexpr int kAbPRows = 32;                      // rows of P^T: N of the output MMA (N % 16 == 0 at M = 128)
constexpr int kAbMaxHeads = 20;
constexpr int kAbBoxBytes = kAbKeys * 128;        // 64 keys x 64 features
constexpr int kAbSlotBytes = 2 * kAbBoxBytes;     // 64 keys x 128 features (one M tile of the output MMA)
constexpr int kAbXSlots = 12;
constexpr int kAbQBoxBytes = kAbQRows * 128;
constexpr int kAbQSlotBytes = 2 * kAbQBoxBytes;
constexpr int kAbQSlots = 3;
constexpr int kAbPBytes = kAbPRows * 128;
constexpr int kAbMiscBytes = 8192;
constexpr int kAbThreads = 256;                   // warp 0 TMA, warp 1 MMA, warp 2 TMEM, warps 4-7 softmax / epilogue
constexpr int kAbMaxDblk = 10;                    // d <= 1280
constexpr int kAbTmemCols = 512;
constexpr int kAbSCol = 0;                        // S^T: 24 columns
constexpr int kAbOCol = 32;                       // O'^T tiles: d / 128 x 32 columns
constexpr int kAbSmemBytes = kAbXSlots * kAbSlotBytes + kAbQSlots * kAbQSlotBytes + kAbPBytes + kAbMiscBytes + 1024;
constexpr int kAbWinPerCta = 4;                   // partial slots per CTA (windows a CTA's tile range can touch)
constexpr int kAbMaxContrib = 32;
constexpr int kAbMaxSeq = 256;

__device__ __forceinline__ uint64_t ab_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void ab_mma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  if (elect_one()) umma_f16(d_tmem, a_desc, b_desc, idesc, accumulate ? 1u : 0u);
}
__device__ __forceinline__ void ab_commit(uint64_t* bar) {
  if (elect_one()) umma_commit(bar);
}
// barrier over the 4 softmax warps that also ORs a predicate across them
__device__ __forceinline__ bool ab_group_or(bool v) {
  uint32_t r;
  asm volatile(
      "{\n\t"
      ".reg .pred p, q;\n\t"
      "setp.ne.b32 p, %1, 0;\n\t"
      "bar.red.or.pred q, 1, 128, p;\n\t"
      "selp.b32 %0, 1, 0, q;\n\t"
      "}\n"
      : "=r"(r)
      : "r"((uint32_t)v)
      : "memory");
  return r != 0;
}
__device__ __forceinline__ float ab_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ int ab_ord(float f) {
  const int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ab_unord(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }
__device__ __forceinline__ void ab_group_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }  // the 4 softmax warps

// The accumulator of an M = 64 tcgen05.mma (cta_group::1) occupies 16 lanes of each 32-lane TMEM quadrant:
// row r lives in lane 32 * (r / 16) + r % 16 (measured with b200w_debug_absorb_probe).
__device__ __forceinline__ int ab_key_of_lane(int quadrant, int lane) { return quadrant * 16 + lane; }
__device__ __forceinline__ bool ab_lane_valid(int lane) { return lane < 16; }

struct AbsorbAttnArgs {
  int n_seq, n_head, d, T;
  const int* finished;   // or null
  const int* slot;       // sequence -> row of xa
  float* part;           // (grid * kAbWinPerCta, n_head, d) f32 unnormalised partial outputs
  float* part_ml;        // (grid * kAbWinPerCta, 2, kAbMaxHeads) running maximum (exp2 domain) and sum
  int* cnt;              // (n_seq) arrival counters, zero on entry, left zero
  __nv_bfloat16* out;    // (n_seq, n_head, d) normalised sum_t p xa_t
  float scale;           // hd^-0.5 * log2(e)
  long long* timeline;   // development aid: clock64 stamps of CTA 0 (8 per tile), or null
};

__global__ void __launch_bounds__(kAbThreads, 1)
absorb_attn_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_q,
                   const __grid_constant__ AbsorbAttnArgs a) {
  extern __shared__ unsigned char ab_smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(ab_smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  unsigned char* sX = smem;
  unsigned char* sQ = sX + kAbXSlots * kAbSlotBytes;
  unsigned char* sP = sQ + kAbQSlots * kAbQSlotBytes;
  unsigned char* misc = sP + kAbPBytes;
  uint64_t* xfull = reinterpret_cast<uint64_t*>(misc);
  uint64_t* xempty = xfull + kAbXSlots;
  uint64_t* qfull = xempty + kAbXSlots;
  uint64_t* qempty = qfull + kAbQSlots;
  uint64_t* s_full = qempty + kAbQSlots;
  uint64_t* p_full = s_full + 1;
  uint64_t* o_done = p_full + 1;
  uint64_t* o_free = o_done + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + 1);
  int* s_nact = reinterpret_cast<int*>(tmem_slot + 1);
  int* s_last = s_nact + 1;
  int* s_ncontrib = s_last + 1;
  int* s_wcnt = s_ncontrib + 1;                                              // [8] per-warp active counts
  int* s_sid = s_wcnt + 8;                                                   // [kAbMaxContrib] partial slots of a window
  float* s_wmax = reinterpret_cast<float*>(misc + 512);                      // [kAbMaxHeads][4]
  float* s_lsum = s_wmax + kAbMaxHeads * 4;                                  // [4][kAbMaxHeads]
  float* s_w = s_lsum + 4 * kAbMaxHeads;                                     // [kAbMaxContrib][kAbMaxHeads] merge weights
  float* s_linv = s_w + kAbMaxContrib * kAbMaxHeads;                         // [kAbMaxHeads]
  short* s_act = reinterpret_cast<short*>(misc + 512 + 4 * (8 * kAbMaxHeads + kAbMaxContrib * kAbMaxHeads + kAbMaxHeads));

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // provably warp-uniform: role code stays on the uniform datapath
  const int H = a.n_head, d = a.d, T = a.T;
  const int n_dblk = d >> 7;
  const int tpw = (T + kAbKeys - 1) / kAbKeys;

  if (tid == 0) {
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_q);
    for (int i = 0; i < kAbXSlots; ++i) {
      mbar_init(&xfull[i], 1);
      mbar_init(&xempty[i], 1);
    }
    for (int i = 0; i < kAbQSlots; ++i) {
      mbar_init(&qfull[i], 1);
      mbar_init(&qempty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(o_done, 1);
    mbar_init(o_free, 128);
    fence_barrier_init();
  }
  // rows >= n_head of P^T stay zero for the whole kernel
  for (int i = tid; i < kAbPBytes / 16; i += kAbThreads) reinterpret_cast<uint4*>(sP)[i] = make_uint4(0, 0, 0, 0);
  // ordered list of the sequences that still decode
  {
    const int i = tid;  // n_seq <= 256 == blockDim
    const bool on = i < a.n_seq && (a.finished == nullptr || a.finished[i] == 0);
    const unsigned int m = __ballot_sync(0xffffffffu, on);
    if (lane == 0) s_wcnt[warp] = __popc(m);
    __syncthreads();
    int base = 0;
    for (int w2 = 0; w2 < warp; ++w2) base += s_wcnt[w2];
    if (on) s_act[base + __popc(m & ((1u << lane) - 1))] = (short)i;
    if (tid == 0) {
      int n = 0;
      for (int w2 = 0; w2 < kAbThreads / 32; ++w2) n += s_wcnt[w2];
      *s_nact = n;
    }
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, kAbTmemCols);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n_act = *s_nact;
  const long long total = (long long)n_act * tpw;
  const int G = gridDim.x, cta = blockIdx.x;
  const int lo = (int)((cta * total) / G), hi = (int)(((cta + 1) * total) / G);
  const int w_first = lo / tpw;

  if (warp == 0) {
    // ------------------------------------------------------------------------------------------ TMA producer: xa tiles
    // (its own thread: a slot frees when the output MMA has read it, independent of the qa ring)
    if (lane == 0) {
      int xs = 0;
      uint32_t xph = 0;
      for (int t = lo; t < hi; ++t) {
        const int w = t / tpw, kt = t - w * tpw, b = s_act[w], row = a.slot[b];
        for (int j = 0; j < n_dblk; ++j) {
          mbar_wait(&xempty[xs], xph ^ 1);
          mbar_expect_tx(&xfull[xs], kAbSlotBytes);
          tma_load_3d(sX + xs * kAbSlotBytes, &tm_x, &xfull[xs], j * 128, kt * kAbKeys, row);
          tma_load_3d(sX + xs * kAbSlotBytes + kAbBoxBytes, &tm_x, &xfull[xs], j * 128 + 64, kt * kAbKeys, row);
          if (++xs == kAbXSlots) {
            xs = 0;
            xph ^= 1;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 3) {
    // ------------------------------------------------------------------------------------------ TMA producer: qa (from L2)
    if (lane == 0) {
      int qs = 0;
      uint32_t qph = 0;
      for (int t = lo; t < hi; ++t) {
        const int b = s_act[t / tpw];
        for (int j = 0; j < n_dblk; ++j) {
          mbar_wait(&qempty[qs], qph ^ 1);
          mbar_expect_tx(&qfull[qs], kAbQSlotBytes);
          tma_load_3d(sQ + qs * kAbQSlotBytes, &tm_q, &qfull[qs], j * 128, 0, b);
          tma_load_3d(sQ + qs * kAbQSlotBytes + kAbQBoxBytes, &tm_q, &qfull[qs], j * 128 + 64, 0, b);
          if (++qs == kAbQSlots) {
            qs = 0;
            qph ^= 1;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------------------------------ MMA issuer
    // The whole warp walks the loop (uniform control flow, descriptors in uniform registers) and one elected lane issues:
    // issued from a divergent single-thread branch every tcgen05.mma costs a 54-cycle ELECT / R2UR loop, this way the
    // score MMA (M 64, N 24) takes 28 cycles and the output MMA (M 128, N 32) 40 -- their shared-memory operand reads.
    {
      constexpr uint32_t idesc_s = make_idesc_bf16(64, kAbQRows, 0, 0);
      constexpr uint32_t idesc_o = make_idesc_bf16(128, kAbPRows, 1, 0);  // A = xa tile read MN-major (features x keys)
      int xs = 0, qs = 0;
      uint32_t xph = 0, qph = 0, pph = 0, ofph = 0;
      bool first_seg = true;
      const uint64_t pd = make_sw128_desc(smem_u32(sP));
      const uint32_t x_base = smem_u32(sX), q_base = smem_u32(sQ);
      for (int t = lo; t < hi; ++t) {
        const int w = t / tpw, kt = t - w * tpw;
        const bool seg_first = (t == lo) || (kt == 0);
        const bool seg_last = (t + 1 == hi) || (kt + 1 == tpw);
        const int xs0 = xs;
        for (int j = 0; j < n_dblk; ++j) {  // S^T (64 keys x 24) = xa_tile (64 x d) qa^T
          mbar_wait(&qfull[qs], qph);
          mbar_wait(&xfull[xs], xph);
          tcgen05_fence_after();
          const uint32_t a0 = x_base + xs * kAbSlotBytes, b0 = q_base + qs * kAbQSlotBytes;
#pragma unroll
          for (int bx = 0; bx < 2; ++bx) {
            const uint64_t ad = make_sw128_desc(a0 + bx * kAbBoxBytes), bd = make_sw128_desc(b0 + bx * kAbQBoxBytes);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) ab_mma(tmem_base + kAbSCol, ad + 2 * kk, bd + 2 * kk, idesc_s, (j | bx | kk) != 0);
          }
          ab_commit(&qempty[qs]);
          if (++qs == kAbQSlots) {
            qs = 0;
            qph ^= 1;
          }
          if (++xs == kAbXSlots) {
            xs = 0;
            xph ^= 1;
          }
        }
        ab_commit(s_full);
        if (a.timeline && cta == 0 && lane == 0 && t - lo < 64) a.timeline[(t - lo) * 8 + 0] = clock64();
        if (seg_first && !first_seg) {  // the previous window's O' has been read out of tensor memory
          mbar_wait(o_free, ofph);
          ofph ^= 1;
        }
        first_seg = false;
        mbar_wait(p_full, pph);
        pph ^= 1;
        tcgen05_fence_after();
        if (a.timeline && cta == 0 && lane == 0 && t - lo < 64) a.timeline[(t - lo) * 8 + 1] = clock64();
        int sl = xs0;
        for (int j = 0; j < n_dblk; ++j) {  // O'^T tile j (128 features x 32) += xa_tile^T (128 x 64 keys) P^T
          const uint32_t a0 = x_base + sl * kAbSlotBytes;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            ab_mma(tmem_base + kAbOCol + j * 32, ab_desc(a0 + kk * 2048, kAbBoxBytes, 1024), pd + 2 * kk, idesc_o,
                   !seg_first || kk != 0);
          ab_commit(&xempty[sl]);
          if (++sl == kAbXSlots) sl = 0;
        }
        if (seg_last) ab_commit(o_done);
        if (a.timeline && cta == 0 && lane == 0 && t - lo < 64) a.timeline[(t - lo) * 8 + 2] = clock64();
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ------------------------------------------------------------------------------------------ softmax / epilogue
    const int q = warp & 3, gt = tid - 128;
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    const int key_l = ab_key_of_lane(q, lane);
    const bool lane_ok = ab_lane_valid(lane);
    // byte offset of this key's element in row h of P^T: h * 128 + (((key >> 3) ^ (h & 7)) << 4) + (key & 7) * 2
    unsigned char* p_col = sP + (key_l & 7) * 2;
    const int key_piece = key_l >> 3;
    float m_run[kAbMaxHeads], l_part[kAbMaxHeads];
    uint32_t sph = 0, odph = 0;
    for (int t = lo; t < hi; ++t) {
      const int w = t / tpw, kt = t - w * tpw;
      const bool seg_first = (t == lo) || (kt == 0);
      const bool seg_last = (t + 1 == hi) || (kt + 1 == tpw);
      if (seg_first) {
#pragma unroll
        for (int h = 0; h < kAbMaxHeads; ++h) {
          m_run[h] = -INFINITY;
          l_part[h] = 0.0f;
        }
      }
      const bool tl = a.timeline && cta == 0 && gt == 0 && t - lo < 64;
      mbar_wait(s_full, sph);
      sph ^= 1;
      tcgen05_fence_after();
      if (tl) a.timeline[(t - lo) * 8 + 3] = clock64();
      uint32_t r[32];
      tmem_ld_32x32(t_lane + kAbSCol, r);
      tmem_wait_ld();
      const bool valid = lane_ok && (kt * kAbKeys + key_l < T);
      float s[kAbMaxHeads];
      float over = -INFINITY;  // how far this key's scores exceed the running maxima
#pragma unroll
      for (int h = 0; h < kAbMaxHeads; ++h) {
        s[h] = valid ? __uint_as_float(r[h]) * a.scale : -INFINITY;
        if (h < H) over = fmaxf(over, s[h] - m_run[h]);
      }
      // Fast path (every tile but a window's first few): no score exceeds its head's running maximum by more than 2^8,
      // the probabilities are taken against the (lagging) running maxima -- one barrier with an OR instead of 20 maxima.
      if (ab_group_or(over > 8.0f)) {
#pragma unroll
        for (int h = 0; h < kAbMaxHeads; ++h) {
          const int wm = __reduce_max_sync(0xffffffffu, ab_ord(s[h]));
          if (lane == h) s_wmax[h * 4 + q] = ab_unord(wm);
        }
        ab_group_sync();
        float alpha[kAbMaxHeads];
#pragma unroll
        for (int h = 0; h < kAbMaxHeads; ++h) {
          const float4 v = *reinterpret_cast<const float4*>(s_wmax + h * 4);
          const float m_new = fmaxf(m_run[h], fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
          alpha[h] = ab_exp2(m_run[h] - m_new);  // 0 on a window's first tile (m_run = -inf)
          m_run[h] = m_new;
          l_part[h] *= alpha[h];
        }
        if (!seg_first) {  // rescale O' (tensor memory); the previous tile's output MMAs completed before S arrived
          for (int j = 0; j < n_dblk; ++j) {
            uint32_t o[32];
            tmem_ld_32x32(t_lane + kAbOCol + j * 32, o);
            tmem_wait_ld();
#pragma unroll
            for (int h = 0; h < kAbMaxHeads; ++h) o[h] = __float_as_uint(__uint_as_float(o[h]) * alpha[h]);
            tmem_st_32x32(t_lane + kAbOCol + j * 32, o);
          }
          tmem_wait_st();
        }
        ab_group_sync();  // s_wmax may be rewritten by the next slow tile
      }
      if (tl) a.timeline[(t - lo) * 8 + 4] = clock64();
      // probabilities (allowed to reach 2^8: harmless in bf16 / fp32) -> P^T rows (heads), K-major over this tile's keys
#pragma unroll
      for (int h = 0; h < kAbMaxHeads; ++h) {
        const float p = ab_exp2(s[h] - m_run[h]);  // 0 for masked keys
        l_part[h] += p;
        if (lane_ok && h < H)
          *reinterpret_cast<__nv_bfloat16*>(p_col + h * 128 + ((key_piece ^ (h & 7)) << 4)) = __float2bfloat16(p);
      }
      fence_proxy_async_smem();
      tcgen05_fence_before();
      mbar_arrive(p_full);
      if (tl) a.timeline[(t - lo) * 8 + 5] = clock64();

      if (seg_last) {
        // ---- flush this CTA's part of window w: (max, sum, O') -> partial slot; the last CTA of the window merges ----
        const int b = s_act[w];
        const int slot_id = cta * kAbWinPerCta + (w - w_first);
#pragma unroll
        for (int h = 0; h < kAbMaxHeads; ++h) {
          const float v = warp_sum(l_part[h]);
          if (lane == 0) s_lsum[q * kAbMaxHeads + h] = v;
        }
        // contributors of window w: the CTAs whose (non-empty) tile range meets [w * tpw, (w + 1) * tpw); their partial
        // slots go to s_sid (warp 4 tests one candidate CTA per lane)
        const long long wlo = (long long)w * tpw, whi = wlo + tpw;
        if (q == 0) {
          const int c_first = (int)(((wlo + 1) * G - 1) / total), c_last = (int)((whi * G - 1) / total);
          const int c2 = c_first + lane;
          const long long l2 = ((long long)c2 * total) / G, h2 = ((long long)(c2 + 1) * total) / G;
          const bool on = c2 <= c_last && h2 > l2 && l2 < whi && h2 > wlo;
          const unsigned int mk = __ballot_sync(0xffffffffu, on);
          if (on) s_sid[__popc(mk & ((1u << lane) - 1))] = c2 * kAbWinPerCta + (w - (int)(l2 / tpw));
          if (lane == 0) *s_ncontrib = __popc(mk);
        }
        mbar_wait(o_done, odph);
        odph ^= 1;
        tcgen05_fence_after();
        float* dst = a.part + (long long)slot_id * H * d;
        for (int j = 0; j < n_dblk; ++j) {
          uint32_t o[32];
          tmem_ld_32x32(t_lane + kAbOCol + j * 32, o);
          tmem_wait_ld();
          const int dim = j * 128 + q * 32 + lane;
#pragma unroll
          for (int h = 0; h < kAbMaxHeads; ++h)
            if (h < H) dst[(long long)h * d + dim] = __uint_as_float(o[h]);
        }
        tcgen05_fence_before();
        mbar_arrive(o_free);
        ab_group_sync();  // s_lsum / s_sid complete
        if (gt < H) {
          float* ml = a.part_ml + (long long)slot_id * 2 * kAbMaxHeads;
          float mv = 0.0f;
#pragma unroll
          for (int h = 0; h < kAbMaxHeads; ++h)
            if (h == gt) mv = m_run[h];
          ml[gt] = mv;
          ml[kAbMaxHeads + gt] = (s_lsum[gt] + s_lsum[kAbMaxHeads + gt]) + (s_lsum[2 * kAbMaxHeads + gt] + s_lsum[3 * kAbMaxHeads + gt]);
        }
        const int n_contrib = *s_ncontrib;
        __threadfence();
        ab_group_sync();
        if (gt == 0) *s_last = (atomicAdd(a.cnt + b, 1) == n_contrib - 1) ? 1 : 0;
        ab_group_sync();
        if (*s_last) {
          __threadfence();
          // weights of the contributors: w_c[h] = 2^(m_c[h] - max_c m_c[h]) / sum_c l_c[h] 2^(m_c[h] - max)
          if (gt < H) {
            float M = -INFINITY;
            for (int k2 = 0; k2 < n_contrib; ++k2) M = fmaxf(M, __ldcg(a.part_ml + (long long)s_sid[k2] * 2 * kAbMaxHeads + gt));
            float L = 0.0f;
            for (int k2 = 0; k2 < n_contrib; ++k2) {
              const float* ml = a.part_ml + (long long)s_sid[k2] * 2 * kAbMaxHeads;
              const float wc = ab_exp2(__ldcg(ml + gt) - M);
              L = fmaf(__ldcg(ml + kAbMaxHeads + gt), wc, L);
              s_w[k2 * kAbMaxHeads + gt] = wc;
            }
            const float linv = 1.0f / L;
            for (int k2 = 0; k2 < n_contrib; ++k2) s_w[k2 * kAbMaxHeads + gt] *= linv;
          }
          ab_group_sync();
          __nv_bfloat16* o = a.out + (long long)b * H * d;
          const int n4 = H * d / 4;  // float4 columns of the (n_head, d) tile; thread gt takes gt, gt + 128, ...
          if (n_contrib <= 3) {      // the common case: up to three contributors, all loads of four columns in flight
            const float* p0 = a.part + (long long)s_sid[0] * H * d;
            const float* p1 = a.part + (long long)s_sid[n_contrib > 1 ? 1 : 0] * H * d;
            const float* p2 = a.part + (long long)s_sid[n_contrib > 2 ? 2 : 0] * H * d;
            for (int i4 = gt; i4 < n4; i4 += 128 * 4) {
              float4 v0[4], v1[4], v2[4];
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const int i = min(i4 + u * 128, n4 - 1);
                v0[u] = __ldcg(reinterpret_cast<const float4*>(p0) + i);
                v1[u] = __ldcg(reinterpret_cast<const float4*>(p1) + i);
                v2[u] = __ldcg(reinterpret_cast<const float4*>(p2) + i);
              }
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const int i = i4 + u * 128;
                if (i < n4) {
                  const int h = (i * 4) / d;
                  const float w0 = s_w[h], w1 = n_contrib > 1 ? s_w[kAbMaxHeads + h] : 0.0f,
                              w2 = n_contrib > 2 ? s_w[2 * kAbMaxHeads + h] : 0.0f;
                  const float ox = fmaf(v2[u].x, w2, fmaf(v1[u].x, w1, v0[u].x * w0));
                  const float oy = fmaf(v2[u].y, w2, fmaf(v1[u].y, w1, v0[u].y * w0));
                  const float oz = fmaf(v2[u].z, w2, fmaf(v1[u].z, w1, v0[u].z * w0));
                  const float ow = fmaf(v2[u].w, w2, fmaf(v1[u].w, w1, v0[u].w * w0));
                  reinterpret_cast<uint2*>(o)[i] = make_uint2(pack_bf16x2(ox, oy), pack_bf16x2(oz, ow));
                }
              }
            }
          } else {
            for (int i = gt; i < n4; i += 128) {
              const int h = (i * 4) / d;
              float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
              for (int k2 = 0; k2 < n_contrib; ++k2) {
                const float4 v = __ldcg(reinterpret_cast<const float4*>(a.part + (long long)s_sid[k2] * H * d) + i);
                const float wc = s_w[k2 * kAbMaxHeads + h];
                acc.x = fmaf(v.x, wc, acc.x);
                acc.y = fmaf(v.y, wc, acc.y);
                acc.z = fmaf(v.z, wc, acc.z);
                acc.w = fmaf(v.w, wc, acc.w);
              }
              reinterpret_cast<uint2*>(o)[i] = make_uint2(pack_bf16x2(acc.x, acc.y), pack_bf16x2(acc.z, acc.w));
            }
          }
          if (gt == 0) a.cnt[b] = 0;  // ready for the next launch
        }
        ab_group_sync();  // s_last / s_w / s_sid / s_lsum are reused by the next window
        if (tl) a.timeline[(t - lo) * 8 + 6] = clock64();
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, kAbTmemCols);
  }
}

// =============================================================================================== K14a
// qa[b][h][i] = sum_j q[b][h*64 + j] Wk[h*64 + j][i],  q = bf16(sum of the split-K slabs + bias) (or a bf16 q).
// CTA = (head, group of kQaIb 64-wide feature blocks, 128-row tile): the A operand (128 rows x 64) is built in shared
// memory by the 4 epilogue warps, the B operand is a (64 j x 64 i) box of Wk consumed MN-major (K6 reads V the same way).
constexpr int kQaIb = 4;
constexpr int kQaThreads = 160;  // warps 0-3: A staging + epilogue, warp 4: TMA + MMA
constexpr int kQaSmem = 16384 + kQaIb * 8192 + 256 + 1024;

struct AbsorbQArgs {
  int rows, n_head, d;
  const float* part;      // split-K slabs of the query projection, or null
  int n_split;
  long long split_stride;
  const float* bias;
  const __nv_bfloat16* q;  // bf16 queries (rows, d) when part == null
  __nv_bfloat16* qa;       // (rows, kAbQRows, d)
};

__global__ void __launch_bounds__(kQaThreads)
absorb_q_kernel(const __grid_constant__ CUtensorMap tm_wk, const __grid_constant__ AbsorbQArgs a) {
  extern __shared__ unsigned char qa_smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(qa_smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  unsigned char* sA = smem;
  unsigned char* sW = smem + 16384;
  uint64_t* wfull = reinterpret_cast<uint64_t*>(sW + kQaIb * 8192);
  uint64_t* dfull = wfull + kQaIb;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dfull + kQaIb);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int h = blockIdx.x, ib0 = blockIdx.y * kQaIb, row0 = blockIdx.z * 128;
  const int d = a.d;
  const int n_ib = min(kQaIb, d / 64 - ib0);

  if (tid == 128) {
    tma_prefetch_desc(&tm_wk);
    for (int i = 0; i < kQaIb; ++i) {
      mbar_init(&wfull[i], 1);
      mbar_init(&dfull[i], 1);
    }
    fence_barrier_init();
    for (int i = 0; i < n_ib; ++i) {
      mbar_expect_tx(&wfull[i], 8192);
      tma_load_3d(sW + i * 8192, &tm_wk, &wfull[i], (ib0 + i) * 64, h * 64, 0);
    }
  }
  if (warp == 4) {
    __syncwarp();
    tmem_alloc(tmem_slot, kQaIb * 64);
    tmem_relinquish();
  } else {
    // A tile: thread covers 4 consecutive columns of rows (tid / 16) + 8 * it; 16-byte pieces XOR-swizzled with row % 8
    const int c4 = (tid & 15) * 4, r0 = tid >> 4;
    const float4 bias = a.part ? *reinterpret_cast<const float4*>(a.bias + h * 64 + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int it = 0; it < 16; ++it) {
      const int r = r0 + 8 * it, row = row0 + r;
      uint2 packed = make_uint2(0, 0);
      if (row < a.rows) {
        if (a.part) {
          float4 v = bias;
          float4 pp[8];
#pragma unroll
          for (int s2 = 0; s2 < 8; ++s2)
            pp[s2] = s2 < a.n_split ? *reinterpret_cast<const float4*>(a.part + s2 * a.split_stride + (long long)row * d + h * 64 + c4)
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int s2 = 0; s2 < 8; ++s2) {
            v.x += pp[s2].x; v.y += pp[s2].y; v.z += pp[s2].z; v.w += pp[s2].w;
          }
          packed = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
        } else {
          packed = *reinterpret_cast<const uint2*>(a.q + (long long)row * d + h * 64 + c4);
        }
      }
      const int piece = (c4 >> 3) ^ (r & 7);
      *reinterpret_cast<uint2*>(sA + r * 128 + piece * 16 + (c4 & 7) * 2) = packed;
    }
    fence_proxy_async_smem();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int uwarp = __shfl_sync(0xffffffffu, warp, 0);

  if (uwarp == 4) {
    constexpr uint32_t idesc = make_idesc_bf16(128, 64, 0, 1);  // B = Wk box: N (features i) contiguous
    const uint64_t ad = make_sw128_desc(smem_u32(sA));
    const uint32_t w_base = smem_u32(sW);
    for (int i = 0; i < n_ib; ++i) {
      mbar_wait(&wfull[i], 0);
      tcgen05_fence_after();
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
        ab_mma(tmem_base + i * 64, ad + 2 * kk, ab_desc(w_base + i * 8192 + kk * 2048, 8192, 1024), idesc, kk != 0);
      ab_commit(&dfull[i]);
    }
    __syncwarp();
  } else {
    const int row = row0 + warp * 32 + lane;
    for (int i = 0; i < n_ib; ++i) {
      mbar_wait(&dfull[i], 0);
      tcgen05_fence_after();
      uint32_t r[64];
      const uint32_t ta = tmem_base + ((uint32_t)(warp * 32) << 16) + i * 64;
      tmem_ld_32x32(ta, reinterpret_cast<uint32_t(&)[32]>(r[0]));
      tmem_ld_32x32(ta + 32, reinterpret_cast<uint32_t(&)[32]>(r[32]));
      tmem_wait_ld();
      if (row < a.rows) {
        uint4* dst = reinterpret_cast<uint4*>(a.qa + ((long long)row * kAbQRows + h) * d + (ib0 + i) * 64);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          dst[k] = make_uint4(pack_bf16x2(__uint_as_float(r[8 * k]), __uint_as_float(r[8 * k + 1])),
                              pack_bf16x2(__uint_as_float(r[8 * k + 2]), __uint_as_float(r[8 * k + 3])),
                              pack_bf16x2(__uint_as_float(r[8 * k + 4]), __uint_as_float(r[8 * k + 5])),
                              pack_bf16x2(__uint_as_float(r[8 * k + 6]), __uint_as_float(r[8 * k + 7])));
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 4) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, kQaIb * 64);
  }
}

// =============================================================================================== K14c
// att[b][h*64 + j] = Wv[h*64 + j] . O'[b][h] + bv[h*64 + j]: per head a (rows x d) x (d x 64) GEMM whose A operand is
// that head's slice of the merged attention output.  CTA = (32-column tile, 128-row tile), K = d in 64-wide blocks.
constexpr int kVoThreads = 192;  // warp 0 TMA, warp 1 MMA + TMEM, warps 2-5 epilogue
constexpr int kVoStages = 4;
constexpr int kVoABytes = 128 * 128, kVoBBytes = 32 * 128;
constexpr int kVoStageBytes = kVoABytes + kVoBBytes;
constexpr int kVoSmem = kVoStages * kVoStageBytes + 256 + 1024;

__global__ void __launch_bounds__(kVoThreads)
absorb_v_kernel(const __grid_constant__ CUtensorMap tm_o, const __grid_constant__ CUtensorMap tm_wv, int rows, int d,
                const float* __restrict__ bias, __nv_bfloat16* __restrict__ att) {
  extern __shared__ unsigned char vo_smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(vo_smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kVoStages * kVoStageBytes);
  uint64_t* empty = full + kVoStages;
  uint64_t* dfull = empty + kVoStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dfull + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nt = blockIdx.x, row0 = blockIdx.y * 128;
  const int head = nt >> 1, num_kb = d / 64;
  if (tid == 0) {
    tma_prefetch_desc(&tm_o);
    tma_prefetch_desc(&tm_wv);
    for (int i = 0; i < kVoStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(dfull, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 32);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int uwarp = __shfl_sync(0xffffffffu, warp, 0);
  if (uwarp == 0) {
    if (lane == 0) {
      int st = 0;
      uint32_t ph = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty[st], ph ^ 1);
        mbar_expect_tx(&full[st], kVoStageBytes);
        tma_load_3d(smem + st * kVoStageBytes, &tm_o, &full[st], head * d + kb * 64, row0, 0);
        tma_load_3d(smem + st * kVoStageBytes + kVoABytes, &tm_wv, &full[st], kb * 64, nt * 32, 0);
        if (++st == kVoStages) {
          st = 0;
          ph ^= 1;
        }
      }
    }
    __syncwarp();
  } else if (uwarp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16(128, 32, 0, 0);
    int st = 0;
    uint32_t ph = 0;
    const uint32_t s_base = smem_u32(smem);
    for (int kb = 0; kb < num_kb; ++kb) {
      mbar_wait(&full[st], ph);
      tcgen05_fence_after();
      const uint32_t sa = s_base + st * kVoStageBytes;
      const uint64_t ad = make_sw128_desc(sa), bd = make_sw128_desc(sa + kVoABytes);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) ab_mma(tmem_base, ad + 2 * kk, bd + 2 * kk, idesc, (kb | kk) != 0);
      ab_commit(&empty[st]);
      if (++st == kVoStages) {
        st = 0;
        ph ^= 1;
      }
    }
    ab_commit(dfull);
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int row = row0 + q * 32 + lane;
    mbar_wait(dfull, 0);
    tcgen05_fence_after();
    uint32_t r[32];
    tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16), r);
    tmem_wait_ld();
    if (row < rows) {
      const float* bp = bias + nt * 32;
      uint4* dst = reinterpret_cast<uint4*>(att + (long long)row * d + nt * 32);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(r[8 * k + e]) + __ldg(bp + 8 * k + e);
        dst[k] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 32);
  }
}

// =============================================================================================== layout probe
// Development aid (tools/probe_absorb.py): one score MMA chain (M = 64, N = 24, K = 128) and one output MMA chain
// (M = 128 features read MN-major with the given LBO / SBO, N = 32, K = 64 keys) on a single 64 x 128 tile; both
// accumulators are dumped lane by lane so that the host can identify the tensor-memory layouts.
__global__ void __launch_bounds__(128)
absorb_probe_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_q,
                    const __grid_constant__ CUtensorMap tm_p, uint32_t lbo, uint32_t sbo, float* dump_s, float* dump_o) {
  extern __shared__ unsigned char pr_smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(pr_smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  unsigned char* sX = smem;
  unsigned char* sQ = sX + kAbSlotBytes;
  unsigned char* sP = sQ + kAbQSlotBytes + 2048;  // keep 1024-byte alignment: 6144 + 2048 = 8192
  uint64_t* bar = reinterpret_cast<uint64_t*>(sP + kAbPBytes);
  uint64_t* done = bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_init(done, 1);
    fence_barrier_init();
    mbar_expect_tx(bar, kAbSlotBytes + kAbQSlotBytes + kAbPBytes);
    tma_load_3d(sX, &tm_x, bar, 0, 0, 0);
    tma_load_3d(sX + kAbBoxBytes, &tm_x, bar, 64, 0, 0);
    tma_load_3d(sQ, &tm_q, bar, 0, 0, 0);
    tma_load_3d(sQ + kAbQBoxBytes, &tm_q, bar, 64, 0, 0);
    tma_load_3d(sP, &tm_p, bar, 0, 0, 0);
  }
  if (warp == 0) {
    __syncwarp();
    tmem_alloc(tmem_slot, 64);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  {  // clear both accumulators so that untouched lanes read as zero
    uint32_t z[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) z[i] = 0;
    tmem_st_32x32(tmem_base + ((uint32_t)(warp * 32) << 16), z);
    tmem_st_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + 32, z);
    tmem_wait_st();
  }
  tcgen05_fence_before();
  __syncthreads();
  if (tid == 0) {
    tcgen05_fence_after();
    mbar_wait(bar, 0);
    tcgen05_fence_after();
    constexpr uint32_t idesc_s = make_idesc_bf16(64, kAbQRows, 0, 0);
    constexpr uint32_t idesc_o = make_idesc_bf16(128, kAbPRows, 1, 0);
    for (int bx = 0; bx < 2; ++bx) {
      const uint64_t ad = make_sw128_desc(smem_u32(sX) + bx * kAbBoxBytes), bd = make_sw128_desc(smem_u32(sQ) + bx * kAbQBoxBytes);
      for (int kk = 0; kk < 4; ++kk) umma_f16(tmem_base, ad + 2 * kk, bd + 2 * kk, idesc_s, (bx | kk) != 0 ? 1u : 0u);
    }
    const uint64_t pd = make_sw128_desc(smem_u32(sP));
    for (int kk = 0; kk < 4; ++kk)
      umma_f16(tmem_base + 32, ab_desc(smem_u32(sX) + kk * 2048, lbo, sbo), pd + 2 * kk, idesc_o, kk != 0 ? 1u : 0u);
    umma_commit(done);
  }
  __syncwarp();
  mbar_wait(done, 0);
  tcgen05_fence_after();
  uint32_t r[32];
  tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16), r);
  tmem_wait_ld();
  for (int i = 0; i < 32; ++i) dump_s[tid * 32 + i] = __uint_as_float(r[i]);
  tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + 32, r);
  tmem_wait_ld();
  for (int i = 0; i < 32; ++i) dump_o[tid * 32 + i] = __uint_as_float(r[i]);
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 64);
  }
}

// Development aid (tools/probe_absorb.py --mma): cycles for `reps` back-to-back tcgen05.mma of one shape (operands are
// whatever the shared memory holds), one commit at the end.  a_mn: A read MN-major with LBO 8192; ts: A from tensor memory.
__global__ void __launch_bounds__(128)
absorb_mma_bench_kernel(int m, int n, int a_mn, int ts, int reps, long long* cycles) {
  extern __shared__ unsigned char mb_smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(mb_smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* done = reinterpret_cast<uint64_t*>(smem + 98304);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 98304 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int uwarp = __shfl_sync(0xffffffffu, warp, 0);  // provably warp-uniform
  if (uwarp == 1) {
    const uint32_t idesc = make_idesc_bf16(m, n, a_mn, 0);
    const uint32_t sa = smem_u32(smem), sb = smem_u32(smem + 65536);
    uint64_t ad[4], bd[4];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      ad[kk] = a_mn ? ab_desc(sa + kk * 2048, 8192, 1024) : make_sw128_desc(sa) + 2 * kk;
      bd[kk] = make_sw128_desc(sb) + 2 * kk;
    }
    const long long t0 = clock64();
    if (ts) {
      for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          if (elect_one())
            asm volatile("tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, 1;" ::"r"(tmem_base),
                         "r"(tmem_base + 256 + kk * 8), "l"(bd[kk]), "r"(idesc)
                         : "memory");
      }
    } else {
      for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          if (elect_one())
            asm volatile("tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, 1;" ::"r"(tmem_base), "l"(ad[kk]), "l"(bd[kk]),
                         "r"(idesc)
                         : "memory");
      }
    }
    const long long t1 = clock64();
    if (elect_one()) umma_commit(done);
    __syncwarp();
    mbar_wait(done, 0);
    const long long t2 = clock64();
    if (elect_one()) {
      cycles[0] = t1 - t0;
      cycles[1] = t2 - t0;
    }
  }
  __syncthreads();
  if (warp == 0) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------- host
static long long* g_absorb_timeline = nullptr;
void set_absorb_timeline(long long* dev) { g_absorb_timeline = dev; }

int init_absorb() {
  static bool done = false;
  if (done) return kOk;
  B200W_CUDA_OK(cudaFuncSetAttribute(absorb_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAbSmemBytes));
  B200W_CUDA_OK(cudaFuncSetAttribute(absorb_q_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kQaSmem));
  B200W_CUDA_OK(cudaFuncSetAttribute(absorb_v_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kVoSmem));
  B200W_CUDA_OK(cudaFuncSetAttribute(absorb_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  done = true;
  return kOk;
}

bool absorb_applicable(int n_seq, int n_head, int d, int T) {
  return n_seq >= 1 && n_seq <= kAbMaxSeq && n_head <= kAbMaxHeads && n_head * 64 == d && d % 128 == 0 &&
         d / 128 <= kAbMaxDblk && T >= 1 && (T + kAbKeys - 1) / kAbKeys <= 24;
}

size_t absorb_workspace_bytes(int n_seq, int n_head, int d) {
  const size_t rows = ((size_t)n_seq + 127) / 128 * 128;
  const size_t G = (size_t)device_sm_count();
  size_t b = 0;
  auto add = [&](size_t x) { b += (x + 1023) / 1024 * 1024; };
  add(rows * kAbQRows * d * 2);                        // qa
  add(rows * n_head * d * 2);                          // merged O'
  add(G * kAbWinPerCta * n_head * d * 4);              // partial O'
  add(G * kAbWinPerCta * 2 * kAbMaxHeads * 4);         // partial (max, sum)
  add(kAbMaxSeq * 4);                                  // arrival counters
  return b;
}

// Workspace layout (absorb_workspace_bytes): qa | merged | part | part_ml | cnt.  `cnt` must be zero before the first use
// (the kernels leave it zero); the padding rows of qa must be zero (absorb_prepare does both).
struct AbsorbWs {
  __nv_bfloat16 *qa, *merged;
  float *part, *part_ml;
  int* cnt;
};
static AbsorbWs carve_absorb(void* ws, int n_seq, int n_head, int d) {
  const size_t rows = ((size_t)n_seq + 127) / 128 * 128;
  const size_t G = (size_t)device_sm_count();
  unsigned char* p = static_cast<unsigned char*>(ws);
  auto take = [&](size_t x) {
    void* r = p;
    p += (x + 1023) / 1024 * 1024;
    return r;
  };
  AbsorbWs o;
  o.qa = static_cast<__nv_bfloat16*>(take(rows * kAbQRows * d * 2));
  o.merged = static_cast<__nv_bfloat16*>(take(rows * n_head * d * 2));
  o.part = static_cast<float*>(take(G * kAbWinPerCta * n_head * d * 4));
  o.part_ml = static_cast<float*>(take(G * kAbWinPerCta * 2 * kAbMaxHeads * 4));
  o.cnt = static_cast<int*>(take(kAbMaxSeq * 4));
  return o;
}

int absorb_prepare(void* ws, int n_seq, int n_head, int d, cudaStream_t stream) {
  AbsorbWs w = carve_absorb(ws, n_seq, n_head, d);
  const size_t rows = ((size_t)n_seq + 127) / 128 * 128;
  B200W_CUDA_OK(cudaMemsetAsync(w.qa, 0, rows * kAbQRows * d * 2, stream));
  B200W_CUDA_OK(cudaMemsetAsync(w.merged, 0, rows * n_head * d * 2, stream));
  B200W_CUDA_OK(cudaMemsetAsync(w.cnt, 0, kAbMaxSeq * 4, stream));
  return kOk;
}

int launch_absorbed_cross_attention(const float* q_part, int n_split, long long split_stride, const float* bias_q,
                                    const __nv_bfloat16* q_bf16, int n_seq, int n_head, const void* w_ckv,
                                    const float* b_ckv, const __nv_bfloat16* xa, int n_slots, int T, const int* slot,
                                    const int* finished, void* ws, __nv_bfloat16* att, cudaStream_t stream) {
  const int d = n_head * 64;
  B200W_CHECK_ARG(absorb_applicable(n_seq, n_head, d, T), "absorbed cross-attention: unsupported shape");
  B200W_CHECK_ARG((q_part && bias_q && n_split >= 1 && n_split <= 8) || q_bf16, "absorbed cross-attention: missing query input");
  B200W_TRY(init_absorb());
  AbsorbWs w = carve_absorb(ws, n_seq, n_head, d);
  const int tiles_m = ceil_div(n_seq, 128);
  {  // K14a
    CUtensorMap twk;
    uint64_t dims[3] = {(uint64_t)d, (uint64_t)d, 1};
    uint64_t strides[2] = {(uint64_t)d * 2, (uint64_t)d * d * 2};
    uint32_t box[3] = {64, 64, 1};
    B200W_TRY(encode_tmap_bf16(&twk, w_ckv, 3, dims, strides, box));
    AbsorbQArgs qa{};
    qa.rows = n_seq;
    qa.n_head = n_head;
    qa.d = d;
    qa.part = q_part;
    qa.n_split = n_split;
    qa.split_stride = split_stride;
    qa.bias = bias_q;
    qa.q = q_bf16;
    qa.qa = w.qa;
    ProfScope prof_("absorb_q", stream);
    B200W_CUDA_OK(launch_k(absorb_q_kernel, dim3(n_head, ceil_div(d / 64, kQaIb), tiles_m), dim3(kQaThreads), kQaSmem, stream,
                           twk, qa));
    count_launch();
  }
  {  // K14b
    CUtensorMap tx, tq;
    uint64_t xd[3] = {(uint64_t)d, (uint64_t)T, (uint64_t)n_slots};
    uint64_t xs[2] = {(uint64_t)d * 2, (uint64_t)T * d * 2};
    uint32_t xb[3] = {64, kAbKeys, 1};
    B200W_TRY(encode_tmap_bf16(&tx, xa, 3, xd, xs, xb));
    uint64_t qd[3] = {(uint64_t)d, (uint64_t)kAbQRows, (uint64_t)n_seq};
    uint64_t qs[2] = {(uint64_t)d * 2, (uint64_t)kAbQRows * d * 2};
    uint32_t qb[3] = {64, kAbQRows, 1};
    B200W_TRY(encode_tmap_bf16(&tq, w.qa, 3, qd, qs, qb));
    AbsorbAttnArgs aa{};
    aa.n_seq = n_seq;
    aa.n_head = n_head;
    aa.d = d;
    aa.T = T;
    aa.finished = finished;
    aa.slot = slot;
    aa.part = w.part;
    aa.part_ml = w.part_ml;
    aa.cnt = w.cnt;
    aa.out = w.merged;
    aa.scale = 0.125f * kAbLog2e;
    aa.timeline = g_absorb_timeline;
    ProfScope prof_("absorb_attn", stream);
    B200W_CUDA_OK(launch_k(absorb_attn_kernel, dim3(device_sm_count()), dim3(kAbThreads), kAbSmemBytes, stream, tx, tq, aa));
    count_launch();
  }
  {  // K14c
    CUtensorMap to, twv;
    B200W_TRY(make_tmap_a(&to, w.merged, 1, n_seq, n_head * d, (long long)n_head * d, (long long)n_seq * n_head * d));
    B200W_TRY(make_tmap_w(&twv, static_cast<const __nv_bfloat16*>(w_ckv) + (size_t)d * d, d, d, 32));
    ProfScope prof_("absorb_v", stream);
    B200W_CUDA_OK(launch_k(absorb_v_kernel, dim3(d / 32, tiles_m), dim3(kVoThreads), kVoSmem, stream, to, twv, n_seq, d,
                           b_ckv + d, att));
    count_launch();
  }
  return kOk;
}

int launch_absorb_mma_bench(int m, int n, int a_mn, int ts, int reps, long long* cycles, cudaStream_t stream) {
  B200W_CUDA_OK(cudaFuncSetAttribute(absorb_mma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100352));
  B200W_CUDA_OK(launch_k(absorb_mma_bench_kernel, dim3(1), dim3(128), 100352, stream, m, n, a_mn, ts, reps, cycles));
  return kOk;
}

int launch_absorb_probe(const __nv_bfloat16* x, const __nv_bfloat16* q, const __nv_bfloat16* p, unsigned int lbo,
                        unsigned int sbo, float* dump_s, float* dump_o, cudaStream_t stream) {
  B200W_TRY(init_absorb());
  CUtensorMap tx, tq, tp;
  uint64_t xd[3] = {128, 64, 1}, xs[2] = {256, 256 * 64};
  uint32_t xb[3] = {64, 64, 1};
  B200W_TRY(encode_tmap_bf16(&tx, x, 3, xd, xs, xb));
  uint64_t qd[3] = {128, kAbQRows, 1}, qs[2] = {256, 256 * kAbQRows};
  uint32_t qb[3] = {64, kAbQRows, 1};
  B200W_TRY(encode_tmap_bf16(&tq, q, 3, qd, qs, qb));
  uint64_t pd[3] = {64, kAbPRows, 1}, ps[2] = {128, 128 * kAbPRows};
  uint32_t pb[3] = {64, kAbPRows, 1};
  B200W_TRY(encode_tmap_bf16(&tp, p, 3, pd, ps, pb));
  B200W_CUDA_OK(launch_k(absorb_probe_kernel, dim3(1), dim3(128), 65536, stream, tx, tq, tp, lbo, sbo, dump_s, dump_o));
  count_launch();
  return kOk;
}

}  // namespace b200w
