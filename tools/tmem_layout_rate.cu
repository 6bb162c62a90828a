// microbenchmark: tcgen05.mma (tf32, A from TMEM, 128x128x8) rate as a function of WHERE in TMEM the accumulator
// and the A operand live (columns), and of alternating between two accumulators
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
__global__ void __launch_bounds__(128, 1) k(long long* out, int rounds, int per_round, int d0, int d1, int a0, int ss) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint64_t bar; __shared__ uint32_t slot;
  uint8_t* base = (uint8_t*)(((uintptr_t)sm + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) ((float*)base)[i] = 0.001f * (i % 7);
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (warp == 0) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot))); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;"); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads(); asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tm = slot;
  constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  if (warp == 1) {
    long long t0 = 0; uint32_t ph = 0;
    const uint32_t a = smem_u32(base), b = smem_u32(base + 16384);
    for (int r = 0; r < rounds + 1; ++r) {
      if (r == 1) t0 = clock64();
      if (elect_one()) {
        for (int i = 0; i < per_round; ++i) {
          const uint32_t koff = (i & 3) * 32;
          const uint32_t dst = tm + (((i >> 2) % 3 == 2) ? d0 : d1);     // 8 MMAs into d1, then 4 into d0 (like the GEMM)
          if (ss) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(dst), "l"(desc(a + koff)), "l"(desc(b + koff)), "r"(idesc), "r"(1u) : "memory");
          else asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(dst), "r"(tm + a0 + (i & 3) * 8), "l"(desc(b + koff)), "r"(idesc), "r"(1u) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      }
      __syncwarp();
      mbar_wait(smem_u32(&bar), ph); ph ^= 1;
    }
    if (threadIdx.x == 32 && blockIdx.x == 0) out[0] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm));
}
int main() {
  long long* d; cudaMalloc(&d, 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int rounds = 100, per = 48;
  int cfgs[][4] = {{0, 0, 256, 0}, {0, 128, 256, 0}, {0, 128, 384, 0}, {0, 256, 128, 0}, {0, 64, 256, 0}, {128, 128, 256, 0}, {0, 256, 384, 0},
                   {0, 128, 448, 0}, {0, 0, 0, 1}, {0, 128, 0, 1}, {0, 256, 0, 1}, {128, 256, 0, 1}};
  for (auto& c : cfgs) {
    cudaMemset(d, 0, 8);
    k<<<148, 128, 64 * 1024>>>(d, rounds, per, c[0], c[1], c[2], c[3]);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("%s main@%3d corr@%3d A@%3d : %7.1f cycles per MMA (%s)\n", c[3] ? "SS" : "TS", c[0], c[1], c[2], (double)h / (rounds * per), cudaGetErrorString(e));
  }
  return 0;
}
