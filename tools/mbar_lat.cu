// microbenchmark: cost of mbarrier.try_wait on an already-completed phase, and arrive->wake latency between warps
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__global__ void k(long long* out) {
  __shared__ uint64_t bars[4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[i]))); asm volatile("fence.mbarrier_init.release.cluster;"); }
  __syncthreads();
  if (threadIdx.x == 0) mbar_arrive(smem_u32(&bars[0]));   // phase 0 of bar0 complete
  __syncthreads();
  if (warp == 0 && lane == 0) {
    long long t0 = clock64();
    for (int i = 0; i < 100; ++i) mbar_wait(smem_u32(&bars[0]), 0);
    out[0] = clock64() - t0;           // 100 completed waits, single lane
  }
  __syncthreads();
  if (warp == 0) {
    long long t0 = clock64();
    for (int i = 0; i < 100; ++i) mbar_wait(smem_u32(&bars[0]), 0);
    if (lane == 0) out[1] = clock64() - t0;   // all 32 lanes polling
  }
  __syncthreads();
  // ping-pong between warp 0 (lane 0) and warp 1 (lane 0): bars[1] w0->w1, bars[2] w1->w0
  if (lane == 0 && warp < 2) {
    uint32_t ph = 0;
    long long t0 = clock64();
    for (int i = 0; i < 100; ++i) {
      if (warp == 0) { mbar_arrive(smem_u32(&bars[1])); mbar_wait(smem_u32(&bars[2]), ph); }
      else { mbar_wait(smem_u32(&bars[1]), ph); mbar_arrive(smem_u32(&bars[2])); }
      ph ^= 1;
    }
    if (warp == 0) out[2] = clock64() - t0;   // 100 round trips
  }
}
int main() {
  long long* d; cudaMalloc(&d, 64); cudaMemset(d, 0, 64);
  k<<<1, 64>>>(d); cudaDeviceSynchronize();
  k<<<1, 64>>>(d); cudaError_t e = cudaDeviceSynchronize();
  long long h[3]; cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
  printf("completed try_wait: %.1f cycles (1 lane), %.1f cycles (32 lanes); arrive->wake round trip %.1f cycles (%s)\n", h[0] / 100.0, h[1] / 100.0, h[2] / 100.0, cudaGetErrorString(e));
  return 0;
}
