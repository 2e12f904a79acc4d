// Microbenchmark: read-only HBM bandwidth of a B200 as a function of the bytes in flight per SM (the ceiling the
// decoder cross-attention, which only reads, can be held against; MEASURED_PEAKS.json is a COPY: half reads, half writes).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o probe_read probe_read.cu && ./probe_read
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int U>
__global__ void read_kernel(const uint4* __restrict__ src, size_t n16, unsigned long long* sink) {
  uint4 acc = make_uint4(0, 0, 0, 0);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + (U - 1) * stride < n16; i += U * stride) {
    uint4 u[U];
#pragma unroll
    for (int k = 0; k < U; ++k) u[k] = __ldg(src + i + k * stride);
#pragma unroll
    for (int k = 0; k < U; ++k) acc.x ^= u[k].x ^ u[k].y ^ u[k].z ^ u[k].w;
  }
  if (acc.x == 0x1234567) sink[0] = acc.x;
}

template <int U>
static void run(const uint4* src, size_t n16, unsigned long long* sink, int sms, int ctas_per_sm, int threads) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e9f;
  for (int r = 0; r < 4; ++r) {
    cudaEventRecord(e0);
    read_kernel<U><<<sms * ctas_per_sm, threads>>>(src, n16, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (r > 0 && ms < best) best = ms;
  }
  printf("ldg.128  %d CTAs/SM x %4d thr x %2d loads = %4d KB in flight per SM : %7.3f ms  %7.1f GB/s  (%s)\n", ctas_per_sm, threads, U,
         ctas_per_sm * threads * U * 16 / 1024, best, n16 * 16.0 / best / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const size_t bytes = (size_t)4 << 30;
  uint4* src;
  unsigned long long* sink;
  cudaMalloc(&src, bytes);
  cudaMalloc(&sink, 8);
  cudaMemset(src, 1, bytes);
  const size_t n16 = bytes / 16;
  for (int c : {1, 2, 4}) {
    for (int t : {256, 512}) {
      if (c * t > 2048) continue;
      run<4>(src, n16, sink, sms, c, t);
      run<8>(src, n16, sink, sms, c, t);
      run<16>(src, n16, sink, sms, c, t);
    }
  }
  run<8>(src, n16, sink, sms, 8, 256);
  run<4>(src, n16, sink, sms, 8, 256);
  // copy for comparison (read + write bytes)
  uint4* dst;
  cudaMalloc(&dst, bytes / 2);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e9f;
  for (int r = 0; r < 4; ++r) {
    cudaEventRecord(e0);
    cudaMemcpyAsync(dst, src, bytes / 2, cudaMemcpyDeviceToDevice);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (r > 0 && ms < best) best = ms;
  }
  printf("cudaMemcpy D2D 2 GiB (read + write bytes) : %7.3f ms  %7.1f GB/s\n", best, (double)bytes / best / 1e6);
  return 0;
}
