// Microbenchmark: how fast can ONE CTA per SM stream global memory into shared memory with cp.async.bulk (TMA 1-D)
// as a function of the copy size and the number of copies in flight -- the weight pipe of the small-batch decode step.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o probe_bulk probe_bulk.cu && ./probe_bulk
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint64_t* b, uint32_t ph) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.b32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(b)), "r"(ph) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void bulk(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// one thread per CTA: `depth` stages of `chunk` bytes, each stage filled by `pieces` bulk copies
__global__ void stream_kernel(const unsigned char* src, size_t per_cta, int chunk, int depth, int pieces, unsigned long long* sink) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bars[16];
  if (threadIdx.x == 0) {
    for (int i = 0; i < depth; ++i) mbar_init(&bars[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  const unsigned char* base = src + (size_t)blockIdx.x * per_cta;
  const int n = (int)(per_cta / chunk);
  const int piece = chunk / pieces;
  auto issue = [&](int i) {
    const int s = i % depth;
    mbar_expect(&bars[s], chunk);
    for (int p = 0; p < pieces; ++p) bulk(smem + (size_t)s * chunk + p * piece, base + (size_t)i * chunk + p * piece, piece, &bars[s]);
  };
  for (int i = 0; i < depth && i < n; ++i) issue(i);
  unsigned long long acc = 0;
  for (int i = 0; i < n; ++i) {
    const int s = i % depth;
    while (!mbar_try(&bars[s], (i / depth) & 1)) {}
    acc += *reinterpret_cast<const unsigned long long*>(smem + (size_t)s * chunk);
    if (i + depth < n) issue(i + depth);
  }
  if (acc == 0x1234567) sink[0] = acc;
}

// the same bytes with plain 16-byte loads: 256 threads, 8 loads in flight each
__global__ void ldg_kernel(const uint4* src, size_t per_cta16, unsigned long long* sink) {
  const uint4* base = src + (size_t)blockIdx.x * per_cta16;
  uint4 acc = make_uint4(0, 0, 0, 0);
  for (size_t i = threadIdx.x; i + 7 * 256 < per_cta16; i += 8 * 256) {
    uint4 u[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) u[k] = __ldg(base + i + k * 256);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc.x ^= u[k].x ^ u[k].y ^ u[k].z ^ u[k].w;
  }
  if (acc.x == 0x1234567) sink[0] = acc.x;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const size_t per_cta = (size_t)12 * 1024 * 1024;  // 12 MB per CTA -> 1.8 GB in total, like one decoder step
  unsigned char* src;
  unsigned long long* sink;
  cudaMalloc(&src, per_cta * sms);
  cudaMalloc(&sink, 8);
  cudaMemset(src, 1, per_cta * sms);
  cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int chunks[] = {4096, 8192, 20480, 40960};
  const int depths[] = {2, 4, 8};
  const int piecess[] = {1, 4};
  for (int chunk : chunks)
    for (int depth : depths)
      for (int pieces : piecess) {
        if ((size_t)chunk * depth > 200 * 1024) continue;
        stream_kernel<<<sms, 32, chunk * depth>>>(src, per_cta, chunk, depth, pieces, sink);
        cudaEventRecord(e0);
        stream_kernel<<<sms, 32, chunk * depth>>>(src, per_cta, chunk, depth, pieces, sink);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        printf("bulk chunk %6d depth %d pieces %d : %7.3f ms  %7.1f GB/s total  %6.1f GB/s per SM  (%s)\n", chunk, depth, pieces, ms,
               per_cta * sms / ms / 1e6, (double)per_cta / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
      }
  ldg_kernel<<<sms, 256>>>((const uint4*)src, per_cta / 16, sink);
  cudaEventRecord(e0);
  ldg_kernel<<<sms, 256>>>((const uint4*)src, per_cta / 16, sink);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  printf("ldg 256 thr x 8 x 16 B            : %7.3f ms  %7.1f GB/s total  %6.1f GB/s per SM\n", ms, per_cta * sms / ms / 1e6, (double)per_cta / ms / 1e6);
  return 0;
}
