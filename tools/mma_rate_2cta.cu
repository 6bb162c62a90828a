// microbenchmark: issue rate of tcgen05.mma.cta_group::2 (TS, tf32) with optional background shared-memory traffic
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mma2 tools/mma_rate_2cta.cu && /tmp/mma2
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t elect_one() { uint32_t pred; asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred)); return pred; }
__device__ __forceinline__ uint64_t desc(uint32_t saddr) {
  uint64_t d = 0; d |= (uint64_t)((saddr & 0x3FFFF) >> 4); d |= (uint64_t)1 << 16; d |= (uint64_t)(1024 >> 4) << 32; d |= (uint64_t)1 << 46; d |= (uint64_t)2 << 61; return d; }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ uint32_t ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void csync() { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }

// CG = 1 or 2 (cta_group), N = MMA N, BG = number of background warps streaming LDS.128 from shared memory
template <int CG, int N, int F16 = 0>
__global__ void __launch_bounds__(512, 1) k(long long* out, int rounds, int per_round, int bg, float* sink) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint64_t bar; __shared__ uint32_t slot; __shared__ int stop;
  uint8_t* base = (uint8_t*)(((uintptr_t)sm + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 512) ((float*)base)[i] = 0.001f * (i % 7);
  if (threadIdx.x == 0) { stop = 0; asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (warp == 0) {
    if (CG == 1) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot))); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;"); }
    else { asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot))); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;"); }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  if (CG == 2) csync(); else __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tm = slot;
  constexpr uint32_t idesc = (1u << 4) | (F16 ? 0u : ((2u << 7) | (2u << 10))) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)((128 * CG) >> 4) << 24);
  const bool leader = CG == 1 || ctarank() == 0;
  if (warp == 1) {
    long long t0 = 0, t1 = 0; uint32_t ph = 0;
    const uint32_t b = smem_u32(base + 16384);
    for (int r = 0; r < rounds + 1; ++r) {
      if (r == 1) t0 = clock64();
      if (leader && elect_one()) {
        for (int i = 0; i < per_round; ++i) {
          const uint32_t koff = (i & 3) * 32;
          if (CG == 1) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tm), "r"(tm + 256 + (i & 3) * 8), "l"(desc(b + koff)), "r"(idesc), "r"(1u) : "memory");
          else if (F16) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tm), "r"(tm + 256 + (i & 3) * 8), "l"(desc(b + koff)), "r"(idesc), "r"(1u) : "memory");
          else asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tm), "r"(tm + 256 + (i & 3) * 8), "l"(desc(b + koff)), "r"(idesc), "r"(1u) : "memory");
        }
        if (CG == 1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        else asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "h"((uint16_t)3) : "memory");
      }
      __syncwarp();
      mbar_wait(smem_u32(&bar), ph); ph ^= 1;
    }
    t1 = clock64();
    if (lane == 0) { if (blockIdx.x == 0) out[0] = t1 - t0; *(volatile int*)&stop = 1; }
  } else if (warp >= 4 && warp < 4 + bg) {
    // background: each warp streams 16 KiB tiles out of shared memory with LDS.128 until the MMA warp is done
    float acc = 0.f; long long n = 0;
    const float4* p = reinterpret_cast<const float4*>(base + 49152) + lane;
    while (!*(volatile int*)&stop) {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float4 v;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_u32(p + j * 32)));
        acc += v.x + v.y + v.z + v.w;
      }
      n += 32;
    }
    if (acc == 123.456f) sink[0] = acc;
    if (lane == 0 && blockIdx.x == 0) out[1 + warp] = n * 512;   // bytes moved by this warp
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  if (CG == 2) csync(); else __syncthreads();
  if (warp == 0) {
    if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm));
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tm));
  }
}
template <int CG, int N, int F16 = 0> void run(int bg) {
  long long* d; cudaMalloc(&d, 8 * 32); cudaMemset(d, 0, 8 * 32);
  float* sink; cudaMalloc(&sink, 4);
  auto fn = k<CG, N, F16>;
  cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
  const int rounds = 200, per = 48;
  cudaLaunchConfig_t lc = {};
  lc.gridDim = dim3(148); lc.blockDim = dim3(512); lc.dynamicSmemBytes = 128 * 1024;
  cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  lc.attrs = at; lc.numAttrs = 1;
  void* args[] = {(void*)&d, (void*)&rounds, (void*)&per, (void*)&bg, (void*)&sink};
  cudaLaunchKernelExC(&lc, (const void*)fn, args);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[32]; cudaMemcpy(h, d, 8 * 32, cudaMemcpyDeviceToHost);
  long long bytes = 0; for (int i = 1; i < 32; ++i) bytes += h[i];
  const double cyc = (double)h[0] / (rounds * per);
  printf("%s cta_group::%d M=%d N=%3d bg_warps=%2d: %7.1f cycles per MMA, MMA B reads %5.1f B/clk, background LDS %5.1f B/clk (%s)\n", F16 ? "f16 " : "tf32", CG, 128 * CG, N, bg, cyc,
         (N / CG) * 32.0 / cyc, (double)bytes / (double)h[0], cudaGetErrorString(e));
  cudaFree(d); cudaFree(sink);
}
int main() {
  for (int bg : {0, 1, 2, 4, 8}) run<1, 128>(bg);
  for (int bg : {0, 1, 2, 4, 8}) run<2, 128>(bg);
  for (int bg : {0, 4}) run<2, 256>(bg);
  for (int bg : {0, 4}) run<1, 256>(bg);
  for (int bg : {0, 4, 8}) run<2, 128, 1>(bg);     // kind::f16, A in TMEM (the fp16-split GEMM's instruction)
  for (int bg : {0, 4}) run<2, 256, 1>(bg);
  return 0;
}
