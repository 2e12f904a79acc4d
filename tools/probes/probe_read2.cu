// Microbenchmark: read-only bandwidth in the ACCESS PATTERN of the decoder cross-attention -- (sequence, head) units of
// 1500 rows x 128 B at a 5120-byte stride (K), then the same for V -- with no arithmetic, to separate what the pattern
// and the unit-per-CTA launch shape cost from what the kernel's own phases cost.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o probe_read2 probe_read2.cu && ./probe_read2
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int T = 1500, H = 20, D = 1280, B = 120;
constexpr long long LD = 2 * D * 2;  // bytes per row: [K | V] bf16

// one CTA per unit, U loads in flight per thread, K pass then V pass (two dependent sweeps like the kernel)
template <int U>
__global__ void unit_kernel(const unsigned char* __restrict__ kv, unsigned long long* sink) {
  const int h = blockIdx.x % H, b = blockIdx.x / H;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, sub = lane & 7, kg = lane >> 3;
  const unsigned char* base = kv + (long long)b * T * LD + h * 128 + sub * 16;
  uint4 acc = make_uint4(0, 0, 0, 0);
  for (int pass = 0; pass < 2; ++pass) {
    const unsigned char* p = base + pass * (D * 2);
    for (int j0 = warp * 4; j0 < T; j0 += U * 32) {
      uint4 u[U];
#pragma unroll
      for (int i = 0; i < U; ++i) {
        const int j = j0 + kg + i * 32;
        u[i] = (j < T) ? __ldg(reinterpret_cast<const uint4*>(p + (long long)j * LD)) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int i = 0; i < U; ++i) acc.x ^= u[i].x ^ u[i].y ^ u[i].z ^ u[i].w;
    }
  }
  if (acc.x == 0x1234567) sink[0] = acc.x;
}

// persistent: grid = SMs x ctas, units dealt round-robin; otherwise as above
template <int U>
__global__ void persistent_kernel(const unsigned char* __restrict__ kv, unsigned long long* sink) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, sub = lane & 7, kg = lane >> 3;
  uint4 acc = make_uint4(0, 0, 0, 0);
  for (int unit = blockIdx.x; unit < B * H; unit += gridDim.x) {
    const int h = unit % H, b = unit / H;
    const unsigned char* base = kv + (long long)b * T * LD + h * 128 + sub * 16;
    for (int pass = 0; pass < 2; ++pass) {
      const unsigned char* p = base + pass * (D * 2);
      for (int j0 = warp * 4; j0 < T; j0 += U * 32) {
        uint4 u[U];
#pragma unroll
        for (int i = 0; i < U; ++i) {
          const int j = j0 + kg + i * 32;
          u[i] = (j < T) ? __ldg(reinterpret_cast<const uint4*>(p + (long long)j * LD)) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int i = 0; i < U; ++i) acc.x ^= u[i].x ^ u[i].y ^ u[i].z ^ u[i].w;
      }
    }
  }
  if (acc.x == 0x1234567) sink[0] = acc.x;
}

// rows of ALL heads by one CTA: a CTA streams whole 5120-byte rows of one sequence (contiguous memory), 16 B per thread
template <int U>
__global__ void rows_kernel(const uint4* __restrict__ kv, size_t n16, unsigned long long* sink) {
  uint4 acc = make_uint4(0, 0, 0, 0);
  const size_t per = (n16 + gridDim.x - 1) / gridDim.x;
  const size_t lo = per * blockIdx.x, hi = lo + per < n16 ? lo + per : n16;
  for (size_t i = lo + threadIdx.x; i < hi; i += (size_t)U * blockDim.x) {
    uint4 u[U];
#pragma unroll
    for (int k = 0; k < U; ++k) u[k] = (i + k * blockDim.x < hi) ? __ldg(kv + i + k * blockDim.x) : make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int k = 0; k < U; ++k) acc.x ^= u[k].x ^ u[k].y ^ u[k].z ^ u[k].w;
  }
  if (acc.x == 0x1234567) sink[0] = acc.x;
}

// one CTA per (sequence, group of G heads): a warp-instruction reads 32 / (8 * G) rows x (G x 128) contiguous bytes
template <int U, int G>
__global__ void group_kernel(const unsigned char* __restrict__ kv, unsigned long long* sink) {
  constexpr int HG = H / G;
  const int hg = blockIdx.x % HG, b = blockIdx.x / HG;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int kLanesPerRow = 8 * G, kRowsPerWarp = 32 / kLanesPerRow, kRowsPerIt = 8 * kRowsPerWarp;
  const int col = lane % kLanesPerRow, kg = lane / kLanesPerRow;
  const unsigned char* base = kv + (long long)b * T * LD + hg * (G * 128) + col * 16;
  uint4 acc = make_uint4(0, 0, 0, 0);
  for (int pass = 0; pass < 2; ++pass) {
    const unsigned char* p = base + pass * (D * 2);
    for (int j0 = warp * kRowsPerWarp; j0 < T; j0 += U * kRowsPerIt) {
      uint4 u[U];
#pragma unroll
      for (int i = 0; i < U; ++i) {
        const int j = j0 + kg + i * kRowsPerIt;
        u[i] = (j < T) ? __ldg(reinterpret_cast<const uint4*>(p + (long long)j * LD)) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int i = 0; i < U; ++i) acc.x ^= u[i].x ^ u[i].y ^ u[i].z ^ u[i].w;
    }
  }
  if (acc.x == 0x1234567) sink[0] = acc.x;
}

template <typename F>
static void timeit(const char* name, double bytes, F launch) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e9f;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (r > 0 && ms < best) best = ms;
  }
  printf("%-64s : %7.1f us  %7.1f GB/s  (%s)\n", name, best * 1e3, bytes / best / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const size_t layer = (size_t)B * T * LD;  // 921.6 MB
  unsigned char* kv;
  unsigned long long* sink;
  cudaMalloc(&kv, 3 * layer);  // three layers, rotated: nothing survives in L2
  cudaMalloc(&sink, 8);
  cudaMemset(kv, 1, 3 * layer);
  int rot = 0;
  auto buf = [&]() { rot = (rot + 1) % 3; return kv + (size_t)rot * layer; };
  timeit("unit per CTA (2400 x 256 thr), 8 loads, K then V", (double)layer, [&] { unit_kernel<8><<<B * H, 256>>>(buf(), sink); });
  timeit("unit per CTA (2400 x 256 thr), 4 loads", (double)layer, [&] { unit_kernel<4><<<B * H, 256>>>(buf(), sink); });
  timeit("unit per CTA (2400 x 256 thr), 16 loads", (double)layer, [&] { unit_kernel<16><<<B * H, 256>>>(buf(), sink); });
  timeit("persistent 4 CTAs/SM x 256 thr, 8 loads, units round-robin", (double)layer, [&] { persistent_kernel<8><<<sms * 4, 256>>>(buf(), sink); });
  timeit("persistent 8 CTAs/SM x 256 thr, 4 loads", (double)layer, [&] { persistent_kernel<4><<<sms * 8, 256>>>(buf(), sink); });
  timeit("persistent 8 CTAs/SM x 256 thr, 8 loads", (double)layer, [&] { persistent_kernel<8><<<sms * 8, 256>>>(buf(), sink); });
  timeit("CTA per (seq, 2 heads) (1200 x 256 thr), 4 loads", (double)layer, [&] { group_kernel<4, 2><<<B * H / 2, 256>>>(buf(), sink); });
  timeit("CTA per (seq, 2 heads) (1200 x 256 thr), 8 loads", (double)layer, [&] { group_kernel<8, 2><<<B * H / 2, 256>>>(buf(), sink); });
  timeit("CTA per (seq, 4 heads) (600 x 256 thr), 4 loads", (double)layer, [&] { group_kernel<4, 4><<<B * H / 4, 256>>>(buf(), sink); });
  timeit("CTA per (seq, 4 heads) (600 x 256 thr), 8 loads", (double)layer, [&] { group_kernel<8, 4><<<B * H / 4, 256>>>(buf(), sink); });
  timeit("CTA per (seq, 4 heads) (600 x 256 thr), 16 loads", (double)layer, [&] { group_kernel<16, 4><<<B * H / 4, 256>>>(buf(), sink); });
  timeit("CTA per (seq, 4 heads) (600 x 512 thr), 8 loads", (double)layer, [&] { group_kernel<8, 4><<<B * H / 4, 512>>>(buf(), sink); });
  timeit("contiguous rows, 4 CTAs/SM x 256 thr, 8 loads (same bytes)", (double)layer, [&] { rows_kernel<8><<<sms * 4, 256>>>((const uint4*)buf(), layer / 16, sink); });
  timeit("contiguous rows, 2400 CTAs x 256 thr, 8 loads", (double)layer, [&] { rows_kernel<8><<<B * H, 256>>>((const uint4*)buf(), layer / 16, sink); });
  return 0;
}
