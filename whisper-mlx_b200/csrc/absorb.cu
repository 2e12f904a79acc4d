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
constexpr int kAbXSlots = 10;
constexpr int kAbQBoxBytes = kAbQRows * 128;
constexpr int kAbQSlotBytes = 2 * kAbQBoxBytes;
constexpr int kAbQSlots = 6;
constexpr int kAbPBytes = kAbPRows * 128;
constexpr int kAbMiscBytes = 4096;
constexpr int kAbExchBytes = 2 * kAbKeys * kAbMaxHeads * 4;  // the peer CTA's partial scores of two tiles
constexpr int kAbThreads = 256;                   // warp 0 TMA, warp 1 MMA, warp 2 TMEM, warps 4-7 softmax / epilogue
constexpr int kAbMaxDblk = 5;                     // 128-feature blocks per CTA: d / 2 <= 640
constexpr int kAbTmemCols = 256;
constexpr int kAbSCol = 0;                        // S^T of two tiles: 2 x 32 columns (24 used)
constexpr int kAbOCol = 64;                       // O'^T tiles: d / 256 x 32 columns
constexpr int kAbSmemBytes = kAbXSlots * kAbSlotBytes + kAbQSlots * kAbQSlotBytes + 2 * kAbPBytes + kAbExchBytes + kAbMiscBytes + 1024;
static_assert(kAbSmemBytes <= 232448, "shared memory budget of one CTA per SM");
constexpr int kAbMaxUnits = 512;                  // partial slots: n_seq * nsplit stays below this when nsplit > 1
constexpr int kAbMaxContrib = 24;
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
  float* part;           // (live sequences * nsplit, n_head, d) f32 unnormalised partial outputs (nsplit > 1)
  float* part_ml;        // (live sequences * nsplit, 2, kAbMaxHeads) running maximum (exp2 domain) and sum
  int* cnt;              // (n_seq) arrival counters, zero on entry, left zero
  __nv_bfloat16* out;    // (n_seq, n_head, d) normalised sum_t p xa_t
  float scale;           // hd^-0.5 * log2(e)
  int nsplit;            // chunks per window (a function of n_seq only: results do not depend on the batch position)
  long long* timeline;   // development aid: clock64 stamps of CTA 0 (8 per tile), or null
};

__device__ __forceinline__ uint32_t ab_cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void ab_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t ab_mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void ab_st_remote_v4(uint32_t cluster_addr, float4 v) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(cluster_addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void ab_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// wait with cluster-scope acquire: the peer's st.shared::cluster data is visible afterwards
__device__ __forceinline__ void ab_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) break;
    if (++spins > (1u << 26)) __trap();
  }
}

// A CTA PAIR (cluster of 2) works on one unit: CTA r holds the features [r d/2, (r+1) d/2) of every key tile (80 KB per
// tile for d = 1280, so that two tiles and more are in flight / resident instead of one), computes the partial scores
// over its features, the two exchange them through distributed shared memory, both run the (identical) softmax and each
// accumulates its half of O'.  The score MMAs of tile t+1 are issued before the output MMAs of tile t.
__global__ void __launch_bounds__(kAbThreads, 1)
absorb_attn_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_q,
                   const __grid_constant__ AbsorbAttnArgs a) {
  extern __shared__ unsigned char ab_smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(ab_smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  unsigned char* sX = smem;
  unsigned char* sQ = sX + kAbXSlots * kAbSlotBytes;
  unsigned char* sP = sQ + kAbQSlots * kAbQSlotBytes;
  float* sE = reinterpret_cast<float*>(sP + 2 * kAbPBytes);      // [2][64 keys][kAbMaxHeads]: the peer's partial scores
  unsigned char* misc = reinterpret_cast<unsigned char*>(sE) + kAbExchBytes;
  uint64_t* xfull = reinterpret_cast<uint64_t*>(misc);
  uint64_t* xempty = xfull + kAbXSlots;
  uint64_t* qfull = xempty + kAbXSlots;
  uint64_t* qempty = qfull + kAbQSlots;
  uint64_t* s_full = qempty + kAbQSlots;   // [2]
  uint64_t* e_full = s_full + 2;           // [2]
  uint64_t* p_full = e_full + 2;
  uint64_t* o_done = p_full + 1;
  uint64_t* o_free = o_done + 1;
  uint64_t* o_sync = o_free + 1;           // completes once per tile, when its output MMAs have finished
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_sync + 1);
  int* s_nact = reinterpret_cast<int*>(tmem_slot + 1);
  int* s_last = s_nact + 1;
  int* s_wcnt = s_last + 1;                                                  // [8] per-warp active counts
  float* s_wmax = reinterpret_cast<float*>(misc + 512);                      // [kAbMaxHeads][4]
  float* s_lsum = s_wmax + kAbMaxHeads * 4;                                  // [4][kAbMaxHeads]
  float* s_w = s_lsum + 4 * kAbMaxHeads;                                     // [kAbMaxContrib][kAbMaxHeads] merge weights
  short* s_act = reinterpret_cast<short*>(misc + 512 + 4 * (8 * kAbMaxHeads + kAbMaxContrib * kAbMaxHeads));

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // provably warp-uniform: role code stays on the uniform datapath
  const int H = a.n_head, d = a.d, T = a.T;
  const int rank = (int)ab_cluster_rank(), peer = rank ^ 1;
  const int dh = d >> 1;             // features of this CTA
  const int n_dblk = dh >> 7;        // 128-feature blocks of this CTA
  const int f0 = rank * dh;
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
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&e_full[i], 1);    // one arrive from the peer per tile (after its 128 softmax threads have stored)
    }
    mbar_init(p_full, 128);
    mbar_init(o_done, 1);
    mbar_init(o_free, 128);
    mbar_init(o_sync, 1);
    fence_barrier_init();
  }
  // rows >= n_head of P^T stay zero for the whole kernel
  for (int i = tid; i < 2 * kAbPBytes / 16; i += kAbThreads) reinterpret_cast<uint4*>(sP)[i] = make_uint4(0, 0, 0, 0);
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
  ab_cluster_sync();  // the peer's barriers exist before anything is sent to them
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n_act = *s_nact;
  // Work units: (live sequence, chunk of its key tiles).  Every window is cut at the SAME tile boundaries whatever the
  // batch holds (nsplit comes from the host, a function of n_seq only), so a window's result does not depend on its
  // position in the batch or on which other sequences have finished.  nsplit == 1: the unit owns the whole window and
  // writes the normalised output itself; else partial (max, sum, O') per chunk, merged by the last chunk to finish.
  const int tpc = (tpw + a.nsplit - 1) / a.nsplit;   // tiles per chunk
  const int nsplit = (tpw + tpc - 1) / tpc;           // chunks that actually hold tiles
  const int n_units = n_act * nsplit;
  const int G = gridDim.x >> 1, cl = blockIdx.x >> 1;  // clusters; both CTAs of a pair walk the same units

  if (warp == 0) {
    // ------------------------------------------------------------------------------------------ TMA producer: xa tiles
    if (lane == 0) {
      int xs = 0;
      uint32_t xph = 0;
      for (int u = cl; u < n_units; u += G) {
        const int w = u / nsplit, chunk = u - w * nsplit, row = a.slot[s_act[w]];
        const int kt1 = min(tpw, (chunk + 1) * tpc);
        for (int kt = chunk * tpc; kt < kt1; ++kt) {
          for (int j = 0; j < n_dblk; ++j) {
            mbar_wait(&xempty[xs], xph ^ 1);
            mbar_expect_tx(&xfull[xs], kAbSlotBytes);
            tma_load_3d(sX + xs * kAbSlotBytes, &tm_x, &xfull[xs], f0 + j * 128, kt * kAbKeys, row);
            tma_load_3d(sX + xs * kAbSlotBytes + kAbBoxBytes, &tm_x, &xfull[xs], f0 + j * 128 + 64, kt * kAbKeys, row);
            if (++xs == kAbXSlots) {
              xs = 0;
              xph ^= 1;
            }
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
      for (int u = cl; u < n_units; u += G) {
        const int w = u / nsplit, chunk = u - w * nsplit, b = s_act[w];
        const int kt1 = min(tpw, (chunk + 1) * tpc);
        for (int kt = chunk * tpc; kt < kt1; ++kt) {
          for (int j = 0; j < n_dblk; ++j) {
            mbar_wait(&qempty[qs], qph ^ 1);
            mbar_expect_tx(&qfull[qs], kAbQSlotBytes);
            tma_load_3d(sQ + qs * kAbQSlotBytes, &tm_q, &qfull[qs], f0 + j * 128, 0, b);
            tma_load_3d(sQ + qs * kAbQSlotBytes + kAbQBoxBytes, &tm_q, &qfull[qs], f0 + j * 128 + 64, 0, b);
            if (++qs == kAbQSlots) {
              qs = 0;
              qph ^= 1;
            }
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
    // Order: scores(t + 1) are issued before output(t), so they run while the softmax warps work on tile t.
    {
      constexpr uint32_t idesc_s = make_idesc_bf16(64, kAbQRows, 0, 0);
      constexpr uint32_t idesc_o = make_idesc_bf16(128, kAbPRows, 1, 0);  // A = xa tile read MN-major (features x keys)
      int xs = 0, qs = 0;
      uint32_t xph = 0, qph = 0, pph = 0, ofph = 0;
      const uint32_t x_base = smem_u32(sX), q_base = smem_u32(sQ), p_base = smem_u32(sP);
      auto scores = [&](int sbuf) -> int {  // S^T (64 keys x 24) = xa_tile (64 x d/2) qa^T; returns the tile's first slot
        const int first = xs;
        for (int j = 0; j < n_dblk; ++j) {
          mbar_wait(&qfull[qs], qph);
          mbar_wait(&xfull[xs], xph);
          tcgen05_fence_after();
          const uint32_t a0 = x_base + xs * kAbSlotBytes, b0 = q_base + qs * kAbQSlotBytes;
#pragma unroll
          for (int bx = 0; bx < 2; ++bx) {
            const uint64_t ad = make_sw128_desc(a0 + bx * kAbBoxBytes), bd = make_sw128_desc(b0 + bx * kAbQBoxBytes);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              ab_mma(tmem_base + kAbSCol + sbuf * 32, ad + 2 * kk, bd + 2 * kk, idesc_s, (j | bx | kk) != 0);
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
        ab_commit(&s_full[sbuf]);
        return first;
      };
      // the CTA's tiles as one sequence: (unit, tile) pairs; `nu`, `nkt` are the tile after `u`, `kt`
      int u = cl, kt = 0, kt1 = 0, ti = 0;
      if (u < n_units) {
        const int chunk = u % nsplit;
        kt = chunk * tpc;
        kt1 = min(tpw, kt + tpc);
      }
      int slot_cur = 0, slot_nxt = 0;
      bool have_cur = false, have_nxt = false;  // scores of the current / next tile already issued
      bool first_unit = true;
      while (u < n_units) {
        const bool seg_first = kt == (u % nsplit) * tpc, seg_last = kt + 1 == kt1;
        int nu = u, nkt = kt + 1, nkt1 = kt1;
        if (nkt == kt1) {
          nu = u + G;
          if (nu < n_units) {
            const int chunk = nu % nsplit;
            nkt = chunk * tpc;
            nkt1 = min(tpw, nkt + tpc);
          }
        }
        if (!have_cur) {
          slot_cur = scores(ti & 1);
          have_cur = true;
        }
        // Look ahead only when it costs nothing: the next tile's features have landed already.  (Issued unconditionally,
        // the next tile's scores would make this tile's output MMAs -- which free the slots the tile after needs -- wait
        // for data that is still in flight.)
        if (nu < n_units && !have_nxt) {
          int last = xs + n_dblk - 1;
          uint32_t lph = xph;
          if (last >= kAbXSlots) {
            last -= kAbXSlots;
            lph ^= 1;
          }
          int qlast = qs + n_dblk - 1;
          uint32_t qlph = qph;
          if (qlast >= kAbQSlots) {
            qlast -= kAbQSlots;
            qlph ^= 1;
          }
          if (mbar_try_wait(&xfull[last], lph) && mbar_try_wait(&qfull[qlast], qlph)) {
            slot_nxt = scores((ti + 1) & 1);
            have_nxt = true;
          }
        }
        if (a.timeline && blockIdx.x == 0 && lane == 0 && ti < 64) a.timeline[ti * 8 + 0] = clock64();
        if (seg_first && !first_unit) {  // the previous unit's O' has been read out of tensor memory
          mbar_wait(o_free, ofph);
          ofph ^= 1;
        }
        mbar_wait(p_full, pph);
        pph ^= 1;
        tcgen05_fence_after();
        if (a.timeline && blockIdx.x == 0 && lane == 0 && ti < 64) a.timeline[ti * 8 + 1] = clock64();
        int sl = slot_cur;
        const uint64_t pd = make_sw128_desc(p_base + (ti & 1) * kAbPBytes);
        for (int j = 0; j < n_dblk; ++j) {  // O'^T tile j (128 features x 32) += xa_tile^T (128 x 64 keys) P^T
          const uint32_t a0 = x_base + sl * kAbSlotBytes;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            ab_mma(tmem_base + kAbOCol + j * 32, ab_desc(a0 + kk * 2048, kAbBoxBytes, 1024), pd + 2 * kk, idesc_o,
                   !seg_first || kk != 0);
          ab_commit(&xempty[sl]);
          if (++sl == kAbXSlots) sl = 0;
        }
        ab_commit(o_sync);
        if (seg_last) {
          ab_commit(o_done);
          first_unit = false;
        }
        if (a.timeline && blockIdx.x == 0 && lane == 0 && ti < 64) a.timeline[ti * 8 + 2] = clock64();
        u = nu;
        kt = nkt;
        kt1 = nkt1;
        slot_cur = slot_nxt;
        have_cur = have_nxt;
        have_nxt = false;
        ++ti;
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
    unsigned char* p_col0 = sP + (key_l & 7) * 2;
    const int key_piece = key_l >> 3;
    const uint32_t peer_e = ab_mapa(smem_u32(sE), peer), peer_ebar = ab_mapa(smem_u32(e_full), peer);
    float m_run[kAbMaxHeads], l_part[kAbMaxHeads];
    uint32_t odph = 0;
    int ti = 0;
    for (int u = cl; u < n_units; u += G) {
      const int w = u / nsplit, chunk = u - w * nsplit;
      const int kt0 = chunk * tpc, kt1 = min(tpw, kt0 + tpc);
      for (int kt = kt0; kt < kt1; ++kt, ++ti) {
        const bool seg_first = kt == kt0, seg_last = kt + 1 == kt1;
        const int sb = ti & 1;
        const uint32_t par = (ti >> 1) & 1;
        if (seg_first) {
#pragma unroll
          for (int h = 0; h < kAbMaxHeads; ++h) {
            m_run[h] = -INFINITY;
            l_part[h] = 0.0f;
          }
        }
        const bool tl = a.timeline && blockIdx.x == 0 && gt == 0 && ti < 64;
        mbar_wait(&s_full[sb], par);
        tcgen05_fence_after();
        if (tl) a.timeline[ti * 8 + 3] = clock64();
        uint32_t r[32];
        tmem_ld_32x32(t_lane + kAbSCol + sb * 32, r);
        tmem_wait_ld();
        // this CTA's partial scores -> the peer's exchange buffer; the peer's arrive in ours
        if (lane_ok) {
          const uint32_t dst = peer_e + (uint32_t)((sb * kAbKeys + key_l) * kAbMaxHeads) * 4u;
#pragma unroll
          for (int h4 = 0; h4 < kAbMaxHeads / 4; ++h4)
            ab_st_remote_v4(dst + h4 * 16, make_float4(__uint_as_float(r[4 * h4]), __uint_as_float(r[4 * h4 + 1]),
                                                       __uint_as_float(r[4 * h4 + 2]), __uint_as_float(r[4 * h4 + 3])));
        }
        ab_group_sync();  // all partial scores of this CTA are stored: one release-arrive on the peer's barrier
        if (gt == 0) ab_arrive_remote(peer_ebar + sb * 8);
        if (lane == 0) ab_wait_cluster(&e_full[sb], par);
        __syncwarp();
        const bool valid = lane_ok && (kt * kAbKeys + key_l < T);
        float s[kAbMaxHeads];
        float over = -INFINITY;  // how far this key's scores exceed the running maxima
        {
          const float4* pe = reinterpret_cast<const float4*>(sE + (sb * kAbKeys + (lane_ok ? key_l : 0)) * kAbMaxHeads);
#pragma unroll
          for (int h4 = 0; h4 < kAbMaxHeads / 4; ++h4) {
            const float4 v = pe[h4];
            s[4 * h4 + 0] = __uint_as_float(r[4 * h4 + 0]) + v.x;
            s[4 * h4 + 1] = __uint_as_float(r[4 * h4 + 1]) + v.y;
            s[4 * h4 + 2] = __uint_as_float(r[4 * h4 + 2]) + v.z;
            s[4 * h4 + 3] = __uint_as_float(r[4 * h4 + 3]) + v.w;
          }
        }
#pragma unroll
        for (int h = 0; h < kAbMaxHeads; ++h) {
          s[h] = valid ? s[h] * a.scale : -INFINITY;
          if (h < H) over = fmaxf(over, s[h] - m_run[h]);
        }
        // Fast path (every tile but a unit's first few): no score exceeds its head's running maximum by more than 2^8,
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
            alpha[h] = ab_exp2(m_run[h] - m_new);  // 0 on a unit's first tile (m_run = -inf)
            m_run[h] = m_new;
            l_part[h] *= alpha[h];
          }
          if (!seg_first) {
            // rescale O' (tensor memory).  The scores of this tile were issued BEFORE the output MMAs of the previous
            // one, so their arrival says nothing about those: wait for the previous tile's o_sync phase (the output MMAs
            // of this tile cannot start before the p_full arrive below, so that phase is the latest one).
            mbar_wait(o_sync, (uint32_t)((ti - 1) & 1));
            tcgen05_fence_after();
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
        if (tl) a.timeline[ti * 8 + 4] = clock64();
        // probabilities (allowed to reach 2^8: harmless in bf16 / fp32) -> P^T rows (heads), K-major over this tile's keys
        // (two P buffers: the output MMAs of the previous tile may still be reading theirs)
        unsigned char* p_col = p_col0 + sb * kAbPBytes;
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
        if (tl) a.timeline[ti * 8 + 5] = clock64();

        if (seg_last) {
          const int b = s_act[w];
#pragma unroll
          for (int h = 0; h < kAbMaxHeads; ++h) {
            const float v = warp_sum(l_part[h]);
            if (lane == 0) s_lsum[q * kAbMaxHeads + h] = v;
          }
          ab_group_sync();
          float l_tot[kAbMaxHeads];
#pragma unroll
          for (int h = 0; h < kAbMaxHeads; ++h)
            l_tot[h] = (s_lsum[h] + s_lsum[kAbMaxHeads + h]) + (s_lsum[2 * kAbMaxHeads + h] + s_lsum[3 * kAbMaxHeads + h]);
          mbar_wait(o_done, odph);
          odph ^= 1;
          tcgen05_fence_after();
          if (nsplit == 1) {
            // ---- the unit is the whole window: normalise and store this CTA's features of sum_t p xa_t ----
            __nv_bfloat16* o = a.out + (long long)b * H * d + f0;
#pragma unroll
            for (int h = 0; h < kAbMaxHeads; ++h) l_tot[h] = 1.0f / l_tot[h];
            for (int j = 0; j < n_dblk; ++j) {
              uint32_t v[32];
              tmem_ld_32x32(t_lane + kAbOCol + j * 32, v);
              tmem_wait_ld();
              const int dim = j * 128 + q * 32 + lane;
#pragma unroll
              for (int h = 0; h < kAbMaxHeads; ++h)
                if (h < H) o[(long long)h * d + dim] = __float2bfloat16(__uint_as_float(v[h]) * l_tot[h]);
            }
            tcgen05_fence_before();
            mbar_arrive(o_free);
          } else {
            // ---- (max, sum, O') of this chunk -> partial slot (sequence, chunk), [feature][head]; absorb_merge_kernel
            //      combines the chunks of a window ----
            const int slot_id = b * nsplit + chunk;
            float* dst = a.part + ((long long)slot_id * d + f0) * kAbMaxHeads;
            for (int j = 0; j < n_dblk; ++j) {
              uint32_t v[32];
              tmem_ld_32x32(t_lane + kAbOCol + j * 32, v);
              tmem_wait_ld();
              float4* o4 = reinterpret_cast<float4*>(dst + (long long)(j * 128 + q * 32 + lane) * kAbMaxHeads);
#pragma unroll
              for (int h4 = 0; h4 < kAbMaxHeads / 4; ++h4)
                o4[h4] = make_float4(__uint_as_float(v[4 * h4]), __uint_as_float(v[4 * h4 + 1]), __uint_as_float(v[4 * h4 + 2]),
                                     __uint_as_float(v[4 * h4 + 3]));
            }
            tcgen05_fence_before();
            mbar_arrive(o_free);
            if (rank == 0 && gt < kAbMaxHeads) {  // (both CTAs of the pair hold the same values)
              float* ml = a.part_ml + (long long)slot_id * 2 * kAbMaxHeads;
              float mv = 0.0f, lv = 0.0f;
#pragma unroll
              for (int h = 0; h < kAbMaxHeads; ++h)
                if (h == gt) {
                  mv = m_run[h];
                  lv = l_tot[h];
                }
              ml[gt] = mv;
              ml[kAbMaxHeads + gt] = lv;
            }
          }
          ab_group_sync();  // s_lsum is reused by the next unit
          if (tl) a.timeline[ti * 8 + 6] = clock64();
        }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  ab_cluster_sync();  // nobody exits while the peer may still write its partial scores here or signal these barriers
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, kAbTmemCols);
  }
}

// Combines the chunks of every live window (nsplit > 1): out[b][h][:] = sum_c w_c[h] O'_c[:][h], w_c[h] =
// 2^(m_c[h] - max_c m_c[h]) / sum_c l_c[h] 2^(m_c[h] - max).  CTA = (sequence, 64-feature block): the [feature][head]
// partial rows are read as they lie and transposed to [head][feature] through shared memory.
constexpr int kMgThreads = 256;
constexpr int kMgDims = 64;
__global__ void __launch_bounds__(kMgThreads)
absorb_merge_kernel(const float* __restrict__ part, const float* __restrict__ part_ml, const int* __restrict__ finished,
                    int n_head, int d, int nsplit, __nv_bfloat16* __restrict__ out) {
  __shared__ float s_w[kAbMaxContrib][kAbMaxHeads];
  __shared__ float s_o[kMgDims][kAbMaxHeads + 1];
  const int b = blockIdx.y, d0 = blockIdx.x * kMgDims, tid = threadIdx.x;
  if (finished != nullptr && finished[b]) return;
  if (tid < kAbMaxHeads) {
    const float* ml = part_ml + (long long)b * nsplit * 2 * kAbMaxHeads;
    float M = -INFINITY;
    for (int c = 0; c < nsplit; ++c) M = fmaxf(M, ml[c * 2 * kAbMaxHeads + tid]);
    float L = 0.0f;
    for (int c = 0; c < nsplit; ++c) {
      const float wc = ab_exp2(ml[c * 2 * kAbMaxHeads + tid] - M);
      L = fmaf(ml[c * 2 * kAbMaxHeads + kAbMaxHeads + tid], wc, L);
      s_w[c][tid] = wc;
    }
    const float linv = 1.0f / L;
    for (int c = 0; c < nsplit; ++c) s_w[c][tid] *= linv;
  }
  __syncthreads();
  // 64 features x 20 heads = 320 float4 per chunk: thread -> (feature, head quad)
  constexpr int kQuads = kAbMaxHeads / 4;
  for (int i = tid; i < kMgDims * kQuads; i += kMgThreads) {
    const int f = i / kQuads, hq = i - f * kQuads;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int c = 0; c < nsplit; ++c) {
      const float4 v = *reinterpret_cast<const float4*>(part + (((long long)b * nsplit + c) * d + d0 + f) * kAbMaxHeads + hq * 4);
      acc.x = fmaf(v.x, s_w[c][hq * 4 + 0], acc.x);
      acc.y = fmaf(v.y, s_w[c][hq * 4 + 1], acc.y);
      acc.z = fmaf(v.z, s_w[c][hq * 4 + 2], acc.z);
      acc.w = fmaf(v.w, s_w[c][hq * 4 + 3], acc.w);
    }
    s_o[f][hq * 4 + 0] = acc.x;
    s_o[f][hq * 4 + 1] = acc.y;
    s_o[f][hq * 4 + 2] = acc.z;
    s_o[f][hq * 4 + 3] = acc.w;
  }
  __syncthreads();
  for (int i = tid; i < n_head * (kMgDims / 2); i += kMgThreads) {
    const int h = i / (kMgDims / 2), f2 = (i - h * (kMgDims / 2)) * 2;
    *reinterpret_cast<uint32_t*>(out + ((long long)b * n_head + h) * d + d0 + f2) = pack_bf16x2(s_o[f2][h], s_o[f2 + 1][h]);
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
    // A tile: thread covers 4 consecutive columns of rows (tid / 16) + 8 * it; 16-byte pieces XOR-swizzled with row % 8.
    // The slab loads of four row groups (up to 32 x 16 bytes per thread) are issued before any of them is used.
    const int c4 = (tid & 15) * 4, r0 = tid >> 4;
    auto put = [&](int r, uint2 packed) {
      const int piece = (c4 >> 3) ^ (r & 7);
      *reinterpret_cast<uint2*>(sA + r * 128 + piece * 16 + (c4 & 7) * 2) = packed;
    };
    if (a.part) {
      const float4 bias = *reinterpret_cast<const float4*>(a.bias + h * 64 + c4);
      const float* base = a.part + (long long)row0 * d + h * 64 + c4;
      const int ns = a.n_split;
#pragma unroll 1
      for (int it0 = 0; it0 < 16; it0 += 4) {
        float4 pp[4][8];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int r = r0 + 8 * (it0 + u);
          const bool on = row0 + r < a.rows;
#pragma unroll
          for (int s2 = 0; s2 < 8; ++s2)
            pp[u][s2] = (on && s2 < ns) ? __ldcg(reinterpret_cast<const float4*>(base + s2 * a.split_stride + (long long)r * d))
                                        : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int r = r0 + 8 * (it0 + u);
          float4 v = bias;
#pragma unroll
          for (int s2 = 0; s2 < 8; ++s2) {
            v.x += pp[u][s2].x; v.y += pp[u][s2].y; v.z += pp[u][s2].z; v.w += pp[u][s2].w;
          }
          put(r, row0 + r < a.rows ? make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w)) : make_uint2(0, 0));
        }
      }
    } else {
#pragma unroll 4
      for (int it = 0; it < 16; ++it) {
        const int r = r0 + 8 * it;
        put(r, row0 + r < a.rows ? *reinterpret_cast<const uint2*>(a.q + (long long)(row0 + r) * d + h * 64 + c4) : make_uint2(0, 0));
      }
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
// that head's slice of the merged attention output.  CTA = (16-column tile, 128-row tile), K = d in 64-wide blocks
// through an 8-stage TMA ring (80 CTAs for d = 1280: the kernel is a latency chain, not work).
constexpr int kVoThreads = 192;  // warp 0 TMA, warp 1 MMA + TMEM, warps 2-5 epilogue
constexpr int kVoStages = 8;
constexpr int kVoBN = 16;
constexpr int kVoABytes = 128 * 128, kVoBBytes = kVoBN * 128;
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
  const int head = (nt * kVoBN) >> 6, num_kb = d / 64;
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
        tma_load_3d(smem + st * kVoStageBytes + kVoABytes, &tm_wv, &full[st], kb * 64, nt * kVoBN, 0);
        if (++st == kVoStages) {
          st = 0;
          ph ^= 1;
        }
      }
    }
    __syncwarp();
  } else if (uwarp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16(128, kVoBN, 0, 0);
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
    if (row < rows) {  // (the accumulator is 16 columns wide; the upper half of the 32-column load is ignored)
      const float* bp = bias + nt * kVoBN;
      uint4* dst = reinterpret_cast<uint4*>(att + (long long)row * d + nt * kVoBN);
#pragma unroll
      for (int k = 0; k < kVoBN / 8; ++k) {
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
  return n_seq >= 1 && n_seq <= kAbMaxSeq && n_head <= kAbMaxHeads && n_head * 64 == d && d % 256 == 0 &&
         d / 256 <= kAbMaxDblk && T >= 1 && (T + kAbKeys - 1) / kAbKeys <= 2 * kAbMaxContrib && (device_sm_count() & 1) == 0;
}

// Chunks per window: the smallest split whose units fill the CTA-pair rounds nearly as well as the best split does
// (a window alone is one unit; 120 windows on 74 pairs would leave the second round 38 % empty).  A function of
// (n_seq, T) only, so that a window's result does not depend on what else is in the batch.
int absorb_nsplit(int n_seq, int T) {
  const int tpw = (T + kAbKeys - 1) / kAbKeys, G = device_sm_count() / 2;
  double fill[kAbMaxContrib + 1];
  double best = 0.0;
  int last = 0;
  for (int ns = 1; ns <= kAbMaxContrib; ++ns) {
    const int tpc = (tpw + ns - 1) / ns, eff = (tpw + tpc - 1) / tpc;
    fill[ns] = 0.0;
    if (eff != ns || (ns > 1 && (tpc < 2 || n_seq * ns > kAbMaxUnits))) continue;  // (not a distinct / allowed split)
    const int units = n_seq * ns, rounds = (units + G - 1) / G;
    fill[ns] = (double)units / ((double)rounds * G);
    if (fill[ns] > best) best = fill[ns];
    last = ns;
  }
  for (int ns = 1; ns <= last; ++ns)
    if (fill[ns] >= 0.93 * best) return ns;
  return 1;
}

size_t absorb_workspace_bytes(int n_seq, int n_head, int d) {
  const size_t rows = ((size_t)n_seq + 127) / 128 * 128;
  size_t b = 0;
  auto add = [&](size_t x) { b += (x + 1023) / 1024 * 1024; };
  add(rows * kAbQRows * d * 2);                        // qa
  add(rows * n_head * d * 2);                          // merged O'
  add((size_t)kAbMaxUnits * d * kAbMaxHeads * 4);      // partial O' ([feature][kAbMaxHeads] per slot)
  add((size_t)kAbMaxUnits * 2 * kAbMaxHeads * 4);      // partial (max, sum)
  add(2 * kAbMaxSeq * 4);                              // arrival counters (per feature half)
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
  unsigned char* p = static_cast<unsigned char*>(ws);
  auto take = [&](size_t x) {
    void* r = p;
    p += (x + 1023) / 1024 * 1024;
    return r;
  };
  AbsorbWs o;
  o.qa = static_cast<__nv_bfloat16*>(take(rows * kAbQRows * d * 2));
  o.merged = static_cast<__nv_bfloat16*>(take(rows * n_head * d * 2));
  o.part = static_cast<float*>(take((size_t)kAbMaxUnits * d * kAbMaxHeads * 4));
  o.part_ml = static_cast<float*>(take((size_t)kAbMaxUnits * 2 * kAbMaxHeads * 4));
  o.cnt = static_cast<int*>(take(2 * kAbMaxSeq * 4));
  return o;
}

int absorb_prepare(void* ws, int n_seq, int n_head, int d, cudaStream_t stream) {
  AbsorbWs w = carve_absorb(ws, n_seq, n_head, d);
  const size_t rows = ((size_t)n_seq + 127) / 128 * 128;
  B200W_CUDA_OK(cudaMemsetAsync(w.qa, 0, rows * kAbQRows * d * 2, stream));
  B200W_CUDA_OK(cudaMemsetAsync(w.merged, 0, rows * n_head * d * 2, stream));
  B200W_CUDA_OK(cudaMemsetAsync(w.cnt, 0, 2 * kAbMaxSeq * 4, stream));
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
    aa.nsplit = absorb_nsplit(n_seq, T);
    aa.timeline = g_absorb_timeline;
    ProfScope prof_("absorb_attn", stream);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(device_sm_count() & ~1);
    cfg.blockDim = dim3(kAbThreads);
    cfg.dynamicSmemBytes = kAbSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;  // CTA pairs: each CTA holds half of the features of a key tile
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    B200W_CUDA_OK(cudaLaunchKernelEx(&cfg, absorb_attn_kernel, tx, tq, aa));
    count_launch();
    if (aa.nsplit > 1) {
      ProfScope prof2_("absorb_merge", stream);
      B200W_CUDA_OK(launch_k(absorb_merge_kernel, dim3(d / kMgDims, n_seq), dim3(kMgThreads), 0, stream,
                             static_cast<const float*>(w.part), static_cast<const float*>(w.part_ml), finished, n_head, d,
                             aa.nsplit, w.merged));
      count_launch();
    }
  }
  {  // K14c
    CUtensorMap to, twv;
    B200W_TRY(make_tmap_a(&to, w.merged, 1, n_seq, n_head * d, (long long)n_head * d, (long long)n_seq * n_head * d));
    B200W_TRY(make_tmap_w(&twv, static_cast<const __nv_bfloat16*>(w_ckv) + (size_t)d * d, d, d, kVoBN));
    ProfScope prof_("absorb_v", stream);
    B200W_CUDA_OK(launch_k(absorb_v_kernel, dim3(d / kVoBN, tiles_m), dim3(kVoThreads), kVoSmem, stream, to, twv, n_seq, d,
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
