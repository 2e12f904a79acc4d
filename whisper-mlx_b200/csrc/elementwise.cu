// K4 LayerNorm, K10 embedding gather, K9 fused logit filter + log-softmax + greedy / sampled token
// selection, plus the tiny no-speech / language-id reductions.  All HBM/L2-bound byte movers.
//
// K9 restates mlx_whisper/decoding.py::{SuppressBlank, SuppressTokens, ApplyTimestampRules,
// GreedyDecoder.update} (reached from /root/reference/run:3-6) as ONE pass structure on
// the device: the reference builds numpy masks on the host each step and synchronises; here the rules
// are evaluated per vocabulary entry from the token history already in HBM.
#include "common.cuh"
#include "kernels.h"

namespace b200w {

// =============================================================================================== K4
constexpr int kLnMaxVec = 10;  // d <= 1280, d % 128 == 0

__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                 int rows, int d, __nv_bfloat16* __restrict__ out_bf16, float* __restrict__ out_f32) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  pdl_wait();
  pdl_launch_dependents();  // after the wait: at most one dependent grid is resident ahead of the running one
  if (row >= rows) return;
  const int nvec = d >> 7;  // float4 per lane
  const float4* xr = reinterpret_cast<const float4*>(x + (long long)row * d);
  float4 v[kLnMaxVec];
  float sum = 0.0f;
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    if (i < nvec) {
      v[i] = xr[i * 32 + lane];
      sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  }
  const float mean = warp_sum(sum) / (float)d;
  float var = 0.0f;
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    if (i < nvec) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, e = v[i].w - mean;
      var += (a * a + b * b) + (c * c + e * e);
    }
  }
  const float rstd = rsqrtf(warp_sum(var) / (float)d + 1e-5f);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
  for (int i = 0; i < kLnMaxVec; ++i) {
    if (i < nvec) {
      const float4 g = __ldg(g4 + i * 32 + lane), bb = __ldg(b4 + i * 32 + lane);
      float4 y;
      y.x = (v[i].x - mean) * rstd * g.x + bb.x;
      y.y = (v[i].y - mean) * rstd * g.y + bb.y;
      y.z = (v[i].z - mean) * rstd * g.z + bb.z;
      y.w = (v[i].w - mean) * rstd * g.w + bb.w;
      const long long off = (long long)row * d + (i * 32 + lane) * 4;
      if (out_bf16 != nullptr)
        *reinterpret_cast<uint2*>(out_bf16 + off) = make_uint2(pack_bf16x2(y.x, y.y), pack_bf16x2(y.z, y.w));
      if (out_f32 != nullptr) *reinterpret_cast<float4*>(out_f32 + off) = y;
    }
  }
}

int launch_layernorm(const float* x, const float* gamma, const float* beta, int rows, int d, __nv_bfloat16* out_bf16,
                     float* out_f32, cudaStream_t stream) {
  B200W_CHECK_ARG(rows > 0 && d > 0 && d % 128 == 0 && d <= 128 * kLnMaxVec, "layernorm: unsupported d=%d", d);
  ProfScope prof_("layernorm", stream);
  B200W_CUDA_OK(launch_k(layernorm_kernel, dim3(ceil_div(rows, 8)), dim3(256), 0, stream, x, gamma, beta, rows, d, out_bf16,
                         out_f32));
  count_launch();
  return kOk;
}

// K4b: residual update + LayerNorm for the decode step (few rows): one CTA per row, one float4 per thread.
//   x <- x + bias + sum_s part[s]   (the split-K partial slabs of the preceding residual GEMM)
//   h <- LayerNorm(x) as bf16
// Fuses the split-K reduction, the bias, the residual add and the next sub-layer's LayerNorm in one launch.
__global__ void __launch_bounds__(320)
resid_ln_small_kernel(float* __restrict__ x, const float* __restrict__ part, int n_split, long long split_stride,
                      const float* __restrict__ bias, const float* __restrict__ gamma, const float* __restrict__ beta,
                      int d, __nv_bfloat16* __restrict__ out_bf16) {
  __shared__ float s_red[2][10];
  const int row = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
  const long long off = (long long)row * d + tid * 4;
  pdl_wait();
  pdl_launch_dependents();  // after the wait: at most one dependent grid is resident ahead of the running one
  float4 v = *reinterpret_cast<const float4*>(x + off);
  if (n_split > 0) {
    const float4 b = __ldg(reinterpret_cast<const float4*>(bias) + tid);
    v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
    // all slabs in flight at once (n_split <= 8): one L2 round trip instead of one per slab
    float4 p[8];
#pragma unroll
    for (int s = 0; s < 8; ++s)
      p[s] = (s < n_split) ? *reinterpret_cast<const float4*>(part + s * split_stride + off) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      v.x += p[s].x; v.y += p[s].y; v.z += p[s].z; v.w += p[s].w;
    }
    *reinterpret_cast<float4*>(x + off) = v;
  }
  float sum = warp_sum((v.x + v.y) + (v.z + v.w));
  if (lane == 0) s_red[0][warp] = sum;
  __syncthreads();
  sum = 0.0f;
  for (int i = 0; i < nwarp; ++i) sum += s_red[0][i];
  const float mean = sum / (float)d;
  const float a = v.x - mean, b2 = v.y - mean, c = v.z - mean, e = v.w - mean;
  float var = warp_sum((a * a + b2 * b2) + (c * c + e * e));
  if (lane == 0) s_red[1][warp] = var;
  __syncthreads();
  var = 0.0f;
  for (int i = 0; i < nwarp; ++i) var += s_red[1][i];
  const float rstd = rsqrtf(var / (float)d + 1e-5f);
  const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + tid), bb = __ldg(reinterpret_cast<const float4*>(beta) + tid);
  *reinterpret_cast<uint2*>(out_bf16 + off) =
      make_uint2(pack_bf16x2(a * rstd * g.x + bb.x, b2 * rstd * g.y + bb.y), pack_bf16x2(c * rstd * g.z + bb.z, e * rstd * g.w + bb.w));
}

int launch_resid_ln_small(float* x, const float* part, int n_split, long long split_stride, const float* bias,
                          const float* gamma, const float* beta, int rows, int d, __nv_bfloat16* out_bf16,
                          cudaStream_t stream) {
  B200W_CHECK_ARG(rows > 0 && d % 128 == 0 && d <= 1280, "resid_ln: unsupported d=%d", d);
  B200W_CHECK_ARG(n_split == 0 || (part != nullptr && bias != nullptr), "resid_ln: partials without bias");
  B200W_CHECK_ARG(n_split >= 0 && n_split <= 8, "resid_ln: at most 8 partial slabs (got %d)", n_split);
  ProfScope prof_("resid_ln", stream);
  B200W_CUDA_OK(launch_k(resid_ln_small_kernel, dim3(rows), dim3(d / 4), 0, stream, x, part, n_split, split_stride, bias,
                         gamma, beta, d, out_bf16));
  count_launch();
  return kOk;
}

// =============================================================================================== K10
__global__ void embed_kernel(const int* __restrict__ tokens, int tokens_ld, const int* __restrict__ pos, int n_q,
                             const __nv_bfloat16* __restrict__ tok_emb, const __nv_bfloat16* __restrict__ pos_emb,
                             int d, int n_ctx, float* __restrict__ x) {
  const int r = blockIdx.x;  // b * n_q + qi
  const int b = r / n_q, qi = r - b * n_q;
  pdl_wait();
  pdl_launch_dependents();  // after the wait: at most one dependent grid is resident ahead of the running one
  int p = pos[b] + qi;
  const int tok = tokens[(long long)b * tokens_ld + p];
  p = min(p, n_ctx - 1);
  const uint32_t* te = reinterpret_cast<const uint32_t*>(tok_emb + (long long)tok * d);
  const uint32_t* pe = reinterpret_cast<const uint32_t*>(pos_emb + (long long)p * d);
  float2* xr = reinterpret_cast<float2*>(x + (long long)r * d);
  for (int i = threadIdx.x; i < d / 2; i += blockDim.x) {
    const float2 a = unpack_bf16x2(__ldg(te + i)), c = unpack_bf16x2(__ldg(pe + i));
    xr[i] = make_float2(a.x + c.x, a.y + c.y);
  }
}

int launch_embed(const int* tokens, int tokens_ld, const int* pos, int n_seq, int n_q, const __nv_bfloat16* tok_emb,
                 const __nv_bfloat16* pos_emb, int d, int n_ctx, float* x, cudaStream_t stream) {
  B200W_CHECK_ARG(n_seq > 0 && n_q > 0 && d % 2 == 0, "embed: bad sizes");
  ProfScope prof_("embed", stream);
  B200W_CUDA_OK(launch_k(embed_kernel, dim3(n_seq * n_q), dim3(128), 0, stream, tokens, tokens_ld, pos, n_q, tok_emb, pos_emb,
                         d, n_ctx, x));
  count_launch();
  return kOk;
}

// =============================================================================================== K9
constexpr int kFaThreads = 1024;

struct ArgMax {
  float v;
  int i;
};
__device__ __forceinline__ ArgMax better(ArgMax a, ArgMax b) {
  // larger value wins; ties go to the lower index (numpy / torch argmax semantics)
  if (b.v > a.v || (b.v == a.v && b.i < a.i)) return b;
  return a;
}
__device__ __forceinline__ ArgMax warp_argmax(ArgMax a) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ArgMax b;
    b.v = __shfl_xor_sync(0xffffffffu, a.v, o);
    b.i = __shfl_xor_sync(0xffffffffu, a.i, o);
    a = better(a, b);
  }
  return a;
}

__device__ __forceinline__ float uniform_hash(unsigned long long seed, unsigned int seq, unsigned int step,
                                              unsigned int v) {
  // counter-based generator: splitmix64 finaliser over (seed, sequence, step, vocab index)
  unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (((unsigned long long)seq << 40) ^
                                                          ((unsigned long long)step << 20) ^ (unsigned long long)v);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return ((float)(z >> 40) + 0.5f) * (1.0f / 16777216.0f);  // (0, 1)
}

// kVec: rows are 16-byte aligned (logits_ld % 4 == 0): a thread loads seven float4 per batch before it looks at any of
// them -- the scalar form's loads sat behind the filter's branches, one round trip to L2 per element and thread
// (62 us for 120 sequences of 51866 logits).
constexpr int kFaBatch = 7;
constexpr int kFaMaskWords = 2048 + 1;  // vocabularies of up to 65536 ids
template <bool kVec>
__global__ void __launch_bounds__(kFaThreads)
filter_argmax_kernel(const float* __restrict__ logits, const uint32_t* __restrict__ suppress_bits,
                     int* __restrict__ tokens, int* __restrict__ n_tokens, int* __restrict__ pos,
                     float* __restrict__ sum_logprob, int* __restrict__ finished, const FilterParams fp) {
  __shared__ ArgMax s_am[3][kFaThreads / 32];
  __shared__ float s_sum[2][kFaThreads / 32];
  __shared__ int s_last_ts_idx;

  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  pdl_wait();
  pdl_launch_dependents();  // after the wait: at most one dependent grid is resident ahead of the running one
  const int n = n_tokens[b];
  int* tok = tokens + (long long)b * fp.tokens_ld;
  const float* lg = logits + (long long)b * fp.logits_ld;
  const int tb = fp.timestamp_begin;

  // ---- token-history state ----
  if (tid == 0) s_last_ts_idx = -1;
  __syncthreads();
  if (fp.apply_timestamp_rules) {
    int best = -1;
    for (int i = fp.sample_begin + tid; i < n; i += kFaThreads)
      if (tok[i] >= tb) best = i;
    if (best >= 0) atomicMax(&s_last_ts_idx, best);
  }
  __syncthreads();
  const int n_sampled = n - fp.sample_begin;
  const int last = tok[n - 1];
  const bool last_ts = n_sampled >= 1 && last >= tb;
  const bool penult_ts = n_sampled < 2 || tok[n - 2] >= tb;
  const bool has_ts = s_last_ts_idx >= 0;
  int ts_floor = tb;
  if (has_ts) {
    const int ts_last = tok[s_last_ts_idx];
    ts_floor = (last_ts && !penult_ts) ? ts_last : ts_last + 1;
  }
  const bool at_begin = (n == fp.sample_begin);

  auto allowed = [&](int v) -> bool {
    if ((__ldg(suppress_bits + (v >> 5)) >> (v & 31)) & 1u) return false;
    if (fp.suppress_blank && at_begin && (v == fp.blank || v == fp.eot)) return false;
    if (fp.apply_timestamp_rules) {
      if (v == fp.no_timestamps) return false;
      if (last_ts) {
        if (penult_ts) {
          if (v >= tb) return false;
        } else {
          if (v < fp.eot) return false;
        }
      }
      if (has_ts && v >= tb && v < ts_floor) return false;
      if (at_begin) {
        if (v < tb) return false;
        if (fp.max_initial_timestamp_index >= 0 && v > tb + fp.max_initial_timestamp_index) return false;
      }
    }
    return true;
  };
  // kVec: the rules above as data -- every rule allows or forbids index RANGES (text ids [t_lo, t_hi), timestamps
  // [s_lo, s_hi]) except three single ids, which are OR-ed into a shared-memory copy of the suppress mask together with
  // the ids past the vocabulary.  Per logit the filter is then a bit test and two unsigned range compares instead of
  // the ~50 compare / select instructions the rule chain compiled to (ncu r02: 64 warp-instructions per logit, the
  // kernel compute-bound at 58 us for 120 sequences).
  __shared__ uint32_t s_mask[kVec ? kFaMaskWords : 1];
  int t_lo = 0, t_w = tb, s_lo = tb, s_w = fp.n_vocab - tb;  // range starts and widths (0: nothing allowed)
  if constexpr (kVec) {
    const int n_words = (fp.n_vocab + 31) / 32, n_words4 = (4 * ((fp.n_vocab + 3) / 4) + 31) / 32;
    for (int i = tid; i < n_words4; i += kFaThreads) {
      uint32_t w = i < n_words ? __ldg(suppress_bits + i) : 0xffffffffu;
      if (i == n_words - 1 && (fp.n_vocab & 31)) w |= 0xffffffffu << (fp.n_vocab & 31);  // ids past the vocabulary
      s_mask[i] = w;
    }
    __syncthreads();
    if (tid == 0) {
      if (fp.suppress_blank && at_begin) {
        s_mask[fp.blank >> 5] |= 1u << (fp.blank & 31);
        s_mask[fp.eot >> 5] |= 1u << (fp.eot & 31);
      }
      if (fp.apply_timestamp_rules) s_mask[fp.no_timestamps >> 5] |= 1u << (fp.no_timestamps & 31);
    }
    __syncthreads();
    if (fp.apply_timestamp_rules) {
      int t_hi = tb, s_hi = fp.n_vocab - 1;
      if (last_ts) {
        if (penult_ts) s_hi = tb - 1;  // no timestamp after a pair
        else t_lo = fp.eot;            // a lone timestamp is followed by a timestamp or EOT
      }
      if (has_ts) s_lo = max(s_lo, ts_floor);
      if (at_begin) {
        t_hi = 0;
        if (fp.max_initial_timestamp_index >= 0) s_hi = min(s_hi, tb + fp.max_initial_timestamp_index);
      }
      t_w = max(t_hi - t_lo, 0);
      s_w = max(s_hi + 1 - s_lo, 0);
    }
  }
  auto allowed4 = [&](int v0, bool (&ok)[4]) {  // four consecutive ids (v0 % 4 == 0) share one word of the mask
    const uint32_t word = s_mask[v0 >> 5] >> (v0 & 31);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int v = v0 + e;
      ok[e] = !((word >> e) & 1u) && ((unsigned)(v - t_lo) < (unsigned)t_w || (unsigned)(v - s_lo) < (unsigned)s_w);
    }
  };
  const float4* lg4 = reinterpret_cast<const float4*>(lg);
  const int n4 = (fp.n_vocab + 3) / 4;

  // ---- pass 1: maxima (text / timestamp) and, when sampling, Gumbel-perturbed maxima ----
  const bool sampling = fp.temperature > 0.0f;
  const float inv_t = sampling ? 1.0f / fp.temperature : 1.0f;
  ArgMax mt{-INFINITY, 0x7fffffff}, ms{-INFINITY, 0x7fffffff};    // plain maxima: text, timestamps
  ArgMax gt{-INFINITY, 0x7fffffff}, gs{-INFINITY, 0x7fffffff};    // perturbed maxima
  float st = 0.0f, ss = 0.0f, rm = -INFINITY;  // kVec: partition sums relative to the thread's running maximum rm
  auto visit1 = [&](int v, float x) {
    ArgMax cur{x, v};
    if (v < tb) mt = better(mt, cur); else ms = better(ms, cur);
    if (sampling) {
      const float u = uniform_hash(fp.seed, b, n, v);
      ArgMax g{x * inv_t - __logf(-__logf(u)), v};
      if (v < tb) gt = better(gt, g); else gs = better(gs, g);
    }
  };
  if constexpr (kVec) {
    for (int c0 = tid; c0 < n4; c0 += kFaThreads * kFaBatch) {
      float4 x[kFaBatch];
#pragma unroll
      for (int k = 0; k < kFaBatch; ++k) {
        const int c = c0 + k * kFaThreads;
        x[k] = c < n4 ? __ldcg(lg4 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int k = 0; k < kFaBatch; ++k) {
        const int c = c0 + k * kFaThreads;
        if (c >= n4) break;
        bool ok[4];
        allowed4(4 * c, ok);
        // forbidden ids take part as -inf: no branch per logit (exp(-inf) = 0, -inf never beats a maximum)
        const float xv[4] = {ok[0] ? x[k].x : -INFINITY, ok[1] ? x[k].y : -INFINITY, ok[2] ? x[k].z : -INFINITY,
                             ok[3] ? x[k].w : -INFINITY};
        const float m4 = fmaxf(fmaxf(xv[0], xv[1]), fmaxf(xv[2], xv[3]));
        if (m4 == -INFINITY) continue;  // nothing allowed in this chunk
        const int v0 = 4 * c;
        const bool all_text = v0 + 3 < tb, all_ts = v0 >= tb;  // (practically warp-uniform: tb sits near the end)
        // (ids ascend within a thread: a strict > keeps the lowest index among equal values)
        if (all_text) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const bool gtr = xv[e] > mt.v;
            mt.v = gtr ? xv[e] : mt.v;
            mt.i = gtr ? v0 + e : mt.i;
          }
        } else if (all_ts) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const bool gtr = xv[e] > ms.v;
            ms.v = gtr ? xv[e] : ms.v;
            ms.i = gtr ? v0 + e : ms.i;
          }
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            if (v0 + e < tb) {
              if (xv[e] > mt.v) mt = ArgMax{xv[e], v0 + e};
            } else {
              if (xv[e] > ms.v) ms = ArgMax{xv[e], v0 + e};
            }
          }
        }
        if (sampling) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            if (!ok[e]) continue;
            const int v = v0 + e;
            const float u = uniform_hash(fp.seed, b, n, v);
            ArgMax g{xv[e] * inv_t - __logf(-__logf(u)), v};
            if (v < tb) gt = better(gt, g); else gs = better(gs, g);
          }
        }
        {  // online partition sums: one rescale per chunk
          const float m_new = fmaxf(rm, m4), sc = __expf(rm - m_new);
          st *= sc;
          ss *= sc;
          const float e0 = __expf(xv[0] - m_new), e1 = __expf(xv[1] - m_new), e2 = __expf(xv[2] - m_new), e3 = __expf(xv[3] - m_new);
          if (all_text) {
            st += (e0 + e1) + (e2 + e3);
          } else if (all_ts) {
            ss += (e0 + e1) + (e2 + e3);
          } else {
            const float ev[4] = {e0, e1, e2, e3};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              if (v0 + e < tb) st += ev[e]; else ss += ev[e];
            }
          }
          rm = m_new;
        }
      }
    }
  } else {
    for (int v = tid; v < fp.n_vocab; v += kFaThreads) {
      if (!allowed(v)) continue;
      visit1(v, lg[v]);
    }
  }
  mt = warp_argmax(mt);
  ms = warp_argmax(ms);
  if (lane == 0) {
    s_am[0][warp] = mt;
    s_am[1][warp] = ms;
  }
  __syncthreads();
  mt = s_am[0][0];
  ms = s_am[1][0];
  for (int i = 1; i < kFaThreads / 32; ++i) {
    mt = better(mt, s_am[0][i]);
    ms = better(ms, s_am[1][i]);
  }
  const float M = fmaxf(mt.v, ms.v);

  // ---- partition sums: a second pass (scalar form) or the per-thread online sums brought to the common maximum ----
  if constexpr (kVec) {
    const float sc = rm > -INFINITY ? __expf(rm - M) : 0.0f;
    st *= sc;
    ss *= sc;
  } else {
    for (int v = tid; v < fp.n_vocab; v += kFaThreads) {
      if (!allowed(v)) continue;
      const float e = __expf(lg[v] - M);
      if (v < tb) st += e; else ss += e;
    }
  }
  st = warp_sum(st);
  ss = warp_sum(ss);
  if (lane == 0) {
    s_sum[0][warp] = st;
    s_sum[1][warp] = ss;
  }
  __syncthreads();
  st = 0.0f;
  ss = 0.0f;
  for (int i = 0; i < kFaThreads / 32; ++i) {
    st += s_sum[0][i];
    ss += s_sum[1][i];
  }

  // "if the probability mass on timestamps exceeds every single text token, emit a timestamp"
  bool text_off = false;
  if (fp.apply_timestamp_rules) text_off = (ss > 0.0f) && (__logf(ss) + M > mt.v);
  if (mt.v == -INFINITY) text_off = true;

  ArgMax pick;
  if (!sampling) {
    pick = text_off ? ms : better(mt, ms);
  } else {
    __syncthreads();
    gt = warp_argmax(gt);
    gs = warp_argmax(gs);
    if (lane == 0) {
      s_am[0][warp] = gt;
      s_am[1][warp] = gs;
    }
    __syncthreads();
    gt = s_am[0][0];
    gs = s_am[1][0];
    for (int i = 1; i < kFaThreads / 32; ++i) {
      gt = better(gt, s_am[0][i]);
      gs = better(gs, s_am[1][i]);
    }
    pick = text_off ? gs : better(gt, gs);
    pick.v = lg[pick.i];
  }

  if (tid == 0) {
    const float Z = text_off ? ss : (st + ss);
    const float logprob = pick.v - M - __logf(Z);
    const bool alive = last != fp.eot;
    int next = pick.i;
    if (alive) sum_logprob[b] += logprob; else next = fp.eot;
    tok[n] = next;
    n_tokens[b] = n + 1;
    pos[b] = n;
    finished[b] = (next == fp.eot) ? 1 : 0;
  }
}

int launch_filter_argmax(const float* logits, const uint32_t* suppress_bits, int* tokens, int* n_tokens, int* pos,
                         float* sum_logprob, int* finished, int n_seq, const FilterParams& fp, cudaStream_t stream) {
  B200W_CHECK_ARG(n_seq > 0 && fp.n_vocab > 0 && fp.logits_ld >= fp.n_vocab, "filter_argmax: bad sizes");
  ProfScope prof_("filter_argmax", stream);
  const bool vec = fp.n_vocab <= 65536 && fp.logits_ld % 4 == 0 && (reinterpret_cast<uintptr_t>(logits) & 15) == 0 && fp.logits_ld >= ((fp.n_vocab + 3) / 4) * 4;
  if (vec)
    B200W_CUDA_OK(launch_k(filter_argmax_kernel<true>, dim3(n_seq), dim3(kFaThreads), 0, stream, logits, suppress_bits, tokens,
                           n_tokens, pos, sum_logprob, finished, fp));
  else
    B200W_CUDA_OK(launch_k(filter_argmax_kernel<false>, dim3(n_seq), dim3(kFaThreads), 0, stream, logits, suppress_bits, tokens,
                           n_tokens, pos, sum_logprob, finished, fp));
  count_launch();
  return kOk;
}

__global__ void __launch_bounds__(1024)
no_speech_kernel(const float* __restrict__ logits, int logits_ld, int n_vocab, int no_speech, float* __restrict__ out) {
  __shared__ float s_red[32];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  pdl_wait();
  pdl_launch_dependents();  // after the wait: at most one dependent grid is resident ahead of the running one
  const float* lg = logits + (long long)b * logits_ld;
  float mx = -INFINITY;
  for (int v = tid; v < n_vocab; v += 1024) mx = fmaxf(mx, lg[v]);
  mx = warp_max(mx);
  if (lane == 0) s_red[warp] = mx;
  __syncthreads();
  mx = s_red[0];
  for (int i = 1; i < 32; ++i) mx = fmaxf(mx, s_red[i]);
  __syncthreads();
  float s = 0.0f;
  for (int v = tid; v < n_vocab; v += 1024) s += __expf(lg[v] - mx);
  s = warp_sum(s);
  if (lane == 0) s_red[warp] = s;
  __syncthreads();
  if (tid == 0) {
    s = 0.0f;
    for (int i = 0; i < 32; ++i) s += s_red[i];
    out[b] = __expf(lg[no_speech] - mx) / s;
  }
}

int launch_no_speech(const float* logits, int logits_ld, int n_seq, int n_vocab, int no_speech, float* out,
                     cudaStream_t stream) {
  ProfScope prof_("no_speech", stream);
  B200W_CUDA_OK(launch_k(no_speech_kernel, dim3(n_seq), dim3(1024), 0, stream, logits, logits_ld, n_vocab, no_speech, out));
  count_launch();
  return kOk;
}

// language id: argmax + softmax restricted to the language tokens (<= 128 of them): one warp per sequence
__global__ void language_kernel(const float* __restrict__ logits, int logits_ld, int lang_begin, int n_lang,
                                int* __restrict__ lang_token, float* __restrict__ lang_probs) {
  const int b = blockIdx.x, lane = threadIdx.x;
  const float* lg = logits + (long long)b * logits_ld + lang_begin;
  ArgMax am{-INFINITY, 0x7fffffff};
  for (int i = lane; i < n_lang; i += 32) am = better(am, ArgMax{lg[i], i});
  am = warp_argmax(am);
  float s = 0.0f;
  for (int i = lane; i < n_lang; i += 32) s += __expf(lg[i] - am.v);
  s = warp_sum(s);
  for (int i = lane; i < n_lang; i += 32) lang_probs[(long long)b * n_lang + i] = __expf(lg[i] - am.v) / s;
  if (lane == 0) lang_token[b] = lang_begin + am.i;
}

int launch_language(const float* logits, int logits_ld, int n_seq, int lang_begin, int n_lang, int* lang_token,
                    float* lang_probs, cudaStream_t stream) {
  ProfScope prof_("language", stream);
  language_kernel<<<n_seq, 32, 0, stream>>>(logits, logits_ld, lang_begin, n_lang, lang_token, lang_probs);
  B200W_LAUNCH_OK();
  count_launch();
  return kOk;
}

}  // namespace b200w
