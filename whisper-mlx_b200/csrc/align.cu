// K12: word-level alignment from cross-attention weights (reference: mlx_whisper/timing.py::find_alignment / dtw /
// median_filter, UPSTREAM; the `--word-timestamps` neighbour of the path; it is what makes the
// `--hallucination-silence-threshold 1` of /root/reference/run:6 take effect).
//
//   probs  : f32 [layer][seq][token][head][T] attention probabilities stored by K8<kProbs>
//   K12a   : per selected (layer, head) and token: 1 / sum of the first n_frames probabilities (the reference takes the
//            softmax over the kept frames only); per (layer, head) and frame: mean and 1/std over the token axis
//   K12b   : matrix[token][frame] = mean over selected heads of median7((w - mean) / std) along frames, reflect-padded
//   K12c   : dynamic time warping over -matrix: anti-diagonal wavefront in one CTA, then the backtrace
#include "common.cuh"
#include "kernels.h"

namespace b200w {

// one warp per (selected head, token)
__global__ void align_rowscale_kernel(const float* __restrict__ probs, long long layer_stride, long long seq_off, int n_tok,
                                      int n_head, int T, const int* __restrict__ heads, int n_frames,
                                      float* __restrict__ scale) {
  const int tok = blockIdx.x, s = blockIdx.y, lane = threadIdx.x;
  const float* w = probs + heads[2 * s] * layer_stride + seq_off + ((long long)tok * n_head + heads[2 * s + 1]) * T;
  float sum = 0.0f;
  for (int f = lane; f < n_frames; f += 32) sum += w[f];
  sum = warp_sum(sum);
  if (lane == 0) scale[s * n_tok + tok] = 1.0f / sum;
}

__global__ void align_stats_kernel(const float* __restrict__ probs, long long layer_stride, long long seq_off, int n_tok,
                                   int n_head, int T, const int* __restrict__ heads, int n_frames,
                                   const float* __restrict__ scale, float* __restrict__ mean, float* __restrict__ rstd) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x, s = blockIdx.y;
  if (f >= n_frames) return;
  const float* w = probs + heads[2 * s] * layer_stride + seq_off + (long long)heads[2 * s + 1] * T + f;
  const float* sc = scale + s * n_tok;
  const long long tok_stride = (long long)n_head * T;
  float sum = 0.0f;
  for (int t = 0; t < n_tok; ++t) sum += w[t * tok_stride] * sc[t];
  const float m = sum / (float)n_tok;
  float var = 0.0f;
  for (int t = 0; t < n_tok; ++t) {
    const float dlt = w[t * tok_stride] * sc[t] - m;
    var = fmaf(dlt, dlt, var);
  }
  mean[s * n_frames + f] = m;
  rstd[s * n_frames + f] = rsqrtf(var / (float)n_tok);
}

__device__ __forceinline__ void cswap(float& a, float& b) {
  const float lo = fminf(a, b), hi = fmaxf(a, b);
  a = lo;
  b = hi;
}

// median of seven by a sorting network (only v[3] is needed)
__device__ __forceinline__ float median7(float (&v)[7]) {
  cswap(v[0], v[6]); cswap(v[2], v[3]); cswap(v[4], v[5]);
  cswap(v[0], v[2]); cswap(v[1], v[4]); cswap(v[3], v[6]);
  cswap(v[0], v[1]); cswap(v[2], v[5]); cswap(v[3], v[4]);
  cswap(v[1], v[2]); cswap(v[4], v[6]);
  cswap(v[2], v[3]); cswap(v[4], v[5]);
  cswap(v[1], v[2]); cswap(v[3], v[4]); cswap(v[5], v[6]);
  return v[3];
}

__global__ void align_matrix_kernel(const float* __restrict__ probs, long long layer_stride, long long seq_off, int n_head,
                                    int T, const int* __restrict__ heads, int n_sel, int n_frames, int n_tok,
                                    const float* __restrict__ scale, const float* __restrict__ mean,
                                    const float* __restrict__ rstd, float* __restrict__ matrix) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x, tok = blockIdx.y;
  if (f >= n_frames) return;
  int idx[7];
#pragma unroll
  for (int k = 0; k < 7; ++k) {  // np.pad(..., mode="reflect"): -1 -> 1, n -> n - 2
    int g = f + k - 3;
    if (g < 0) g = -g;
    if (g >= n_frames) g = 2 * (n_frames - 1) - g;
    idx[k] = max(0, min(n_frames - 1, g));
  }
  float acc = 0.0f;
  for (int s = 0; s < n_sel; ++s) {
    const float* w = probs + heads[2 * s] * layer_stride + seq_off + ((long long)tok * n_head + heads[2 * s + 1]) * T;
    const float sc = scale[s * n_tok + tok];
    float v[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) v[k] = (w[idx[k]] * sc - mean[s * n_frames + idx[k]]) * rstd[s * n_frames + idx[k]];
    acc += median7(v);
  }
  matrix[(long long)tok * n_frames + f] = acc / (float)n_sel;
}

// cost[i][j] = x[i-1][j-1] + min(cost[i-1][j-1], cost[i-1][j], cost[i][j-1]) with the reference's tie rule
// (diagonal only if strictly smallest, then up only if strictly smallest, else left); x = -matrix rows [row0, row0 + N)
__global__ void __launch_bounds__(1024)
dtw_kernel(const float* __restrict__ matrix, long long ld, int N, int M, float* __restrict__ cost,
           signed char* __restrict__ trace, int* __restrict__ text_idx, int* __restrict__ time_idx,
           int* __restrict__ path_len) {
  const int W = M + 1;
  for (int k = threadIdx.x; k < (N + 1) * W; k += blockDim.x) {
    const int i = k / W, j = k - i * W;
    cost[k] = (i == 0 && j == 0) ? 0.0f : INFINITY;
    trace[k] = (i == 0) ? 2 : (j == 0 ? 1 : -1);
  }
  __syncthreads();
  for (int dgl = 2; dgl <= N + M; ++dgl) {
    const int i_lo = max(1, dgl - M), i_hi = min(N, dgl - 1);
    for (int i = i_lo + threadIdx.x; i <= i_hi; i += blockDim.x) {
      const int j = dgl - i;
      const float c0 = cost[(i - 1) * W + j - 1], c1 = cost[(i - 1) * W + j], c2 = cost[i * W + j - 1];
      float c;
      signed char t;
      if (c0 < c1 && c0 < c2) {
        c = c0;
        t = 0;
      } else if (c1 < c0 && c1 < c2) {
        c = c1;
        t = 1;
      } else {
        c = c2;
        t = 2;
      }
      cost[i * W + j] = -matrix[(long long)(i - 1) * ld + (j - 1)] + c;
      trace[i * W + j] = t;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    int i = N, j = M, n = 0;
    while (i > 0 || j > 0) {  // written back to front; the host reverses
      text_idx[n] = i - 1;
      time_idx[n] = j - 1;
      ++n;
      const signed char t = trace[i * W + j];
      if (t == 0) {
        --i;
        --j;
      } else if (t == 1) {
        --i;
      } else {
        --j;
      }
    }
    *path_len = n;
  }
}

int launch_alignment_matrix(const float* probs, long long layer_stride, long long seq_off, int n_tok, int n_head, int T,
                            const int* heads, int n_sel, int n_frames, float* stats, float* matrix, cudaStream_t stream) {
  B200W_CHECK_ARG(n_tok > 0 && n_sel > 0 && n_frames > 0 && n_frames <= T, "alignment_matrix: bad sizes");
  // stats: mean [n_sel][n_frames], rstd [n_sel][n_frames], scale [n_sel][n_tok]
  float* mean = stats;
  float* rstd = stats + (size_t)n_sel * n_frames;
  float* scale = stats + 2 * (size_t)n_sel * n_frames;
  ProfScope prof_("alignment_matrix", stream);
  align_rowscale_kernel<<<dim3(n_tok, n_sel), 32, 0, stream>>>(probs, layer_stride, seq_off, n_tok, n_head, T, heads, n_frames,
                                                             scale);
  B200W_LAUNCH_OK();
  align_stats_kernel<<<dim3(ceil_div(n_frames, 128), n_sel), 128, 0, stream>>>(probs, layer_stride, seq_off, n_tok, n_head, T,
                                                                            heads, n_frames, scale, mean, rstd);
  B200W_LAUNCH_OK();
  align_matrix_kernel<<<dim3(ceil_div(n_frames, 128), n_tok), 128, 0, stream>>>(probs, layer_stride, seq_off, n_head, T, heads,
                                                                             n_sel, n_frames, n_tok, scale, mean, rstd, matrix);
  B200W_LAUNCH_OK();
  count_launch(3);
  return kOk;
}

int launch_dtw(const float* matrix, long long ld, int N, int M, float* cost, signed char* trace, int* text_idx,
               int* time_idx, int* path_len, cudaStream_t stream) {
  B200W_CHECK_ARG(N > 0 && M > 0 && (long long)(N + 1) * (M + 1) < (1ll << 30), "dtw: bad sizes");
  ProfScope prof_("dtw", stream);
  dtw_kernel<<<1, 1024, 0, stream>>>(matrix, ld, N, M, cost, trace, text_idx, time_idx, path_len);
  B200W_LAUNCH_OK();
  count_launch();
  return kOk;
}

}  // namespace b200w
