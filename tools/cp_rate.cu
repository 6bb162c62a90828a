// microbenchmark + layout check for feeding the A operand of tcgen05.mma.cta_group::2.kind::f16 (tools only):
//   (1) layout: A tile [256 rows][64 fp16] K-major SWIZZLE_128B in shared memory (128 rows per CTA) is multiplied
//       with B three ways -- SS (A descriptor), cp+TS (tcgen05.cp.128x256b into TMEM, A read from TMEM) -- and both are
//       compared with the host product: proves the TMEM image tcgen05.cp leaves is the one a TS MMA expects;
//   (2) rate: cycles per "K block" (12 MMAs, the fp16-split GEMM's unit of work) for
//         TS   : A already in TMEM (today's kernel once the splitter warps have written it)
//         CPTS : 8 tcgen05.cp (h0 + h1 halves, 32 KiB per CTA) into the other TMEM slot, then the 12 MMAs
//         SS   : all 12 MMAs read A through shared-memory descriptors
//       each with 0 / 4 background warps streaming LDS.128 (stand-in for TMA writes + epilogue traffic).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/cp_rate tools/cp_rate.cu && /tmp/cp_rate
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_fp16.h>
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
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void cp_128x256b(uint32_t taddr, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::2.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}
__device__ __forceinline__ void commit2(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);   // f16 x f16 -> f32, M 256, N 128
// shared-memory map (per CTA): A stages 3 x 32 KiB (h0 16 KiB | h1 16 KiB), B 2 x 8 KiB, background area 48 KiB
constexpr int A_STAGE = 32768, B_OFF = 3 * A_STAGE, BG_OFF = B_OFF + 16384, SMEM = BG_OFF + 49152 + 1024;

// ---- (1) layout check ---------------------------------------------------------------------------------
// gA [256][64] fp16, gB [128][64] fp16 (row major); out[mode][256][128] fp32, mode 0 = SS, 1 = cp + TS
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) check_kernel(const __half* gA, const __half* gB, float* out) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint64_t bar; __shared__ uint32_t slot;
  uint8_t* base = (uint8_t*)(((uintptr_t)sm + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cr = ctarank();
  // swizzled K-major images: row r (128 B) in 8-row groups of 1024 B, 16-byte chunk c stored at chunk c ^ (r & 7)
  for (int i = threadIdx.x; i < 128 * 8; i += 128) {
    const int r = i >> 3, c = i & 7;
    *reinterpret_cast<uint4*>(base + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(gA + (cr * 128 + r) * 64 + c * 8);
  }
  for (int i = threadIdx.x; i < 64 * 8; i += 128) {
    const int r = i >> 3, c = i & 7;
    *reinterpret_cast<uint4*>(base + B_OFF + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(gB + (cr * 64 + r) * 64 + c * 8);
  }
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (warp == 0) { asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot))); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;"); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads(); csync();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tm = slot;
  const uint32_t a = smem_u32(base), b = smem_u32(base + B_OFF);
  if (warp == 1 && cr == 0 && elect_one()) {
    for (int k = 0; k < 4; ++k) mma_ss(tm, desc(a + 32 * k), desc(b + 32 * k), IDESC, k != 0);          // D0: columns 0..127
    for (int k = 0; k < 4; ++k) cp_128x256b(tm + 256 + 8 * k, desc(a + 32 * k));                         // A image: columns 256..287
    for (int k = 0; k < 4; ++k) mma_ts(tm + 128, tm + 256 + 8 * k, desc(b + 32 * k), IDESC, k != 0);     // D1: columns 128..255
    commit2(smem_u32(&bar));
  }
  __syncwarp();
  mbar_wait(smem_u32(&bar), 0);
  asm volatile("tcgen05.fence::after_thread_sync;");
  for (int mode = 0; mode < 2; ++mode)
    for (int ch = 0; ch < 4; ++ch) {
      uint32_t v[32];
      ld32(tm + ((uint32_t)(warp * 32) << 16) + mode * 128 + ch * 32, v);
      for (int e = 0; e < 32; ++e) out[(mode * 256 + cr * 128 + warp * 32 + lane) * 128 + ch * 32 + e] = __uint_as_float(v[e]);
    }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads(); csync();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tm));
}

// ---- (2) rates ------------------------------------------------------------------------------------------
// MODE 0 TS, 1 CPTS, 2 SS, 3 CP only
template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(512, 1) rate_kernel(long long* out, int rounds, int kb_per_round, int bg, float* sink) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint64_t bar; __shared__ uint32_t slot; __shared__ int stop;
  uint8_t* base = (uint8_t*)(((uintptr_t)sm + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (SMEM - 1024) / 4; i += 512) ((uint32_t*)base)[i] = 0x14001400u + (i % 7);     // small fp16 values
  if (threadIdx.x == 0) { stop = 0; asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (warp == 0) { asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot))); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;"); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads(); csync();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tm = slot;
  const bool leader = ctarank() == 0;
  if (warp == 1) {
    long long t0 = 0, t1 = 0; uint32_t ph = 0;
    const uint32_t b_hi = smem_u32(base + B_OFF), b_lo = b_hi + 8192;
    int kbg = 0;
    for (int r = 0; r < rounds + 1; ++r) {
      if (r == 1) t0 = clock64();
      if (leader && elect_one()) {
        for (int i = 0; i < kb_per_round; ++i, ++kbg) {
          const uint32_t a0 = smem_u32(base + (kbg % 3) * A_STAGE), a1 = a0 + 16384;
          const uint32_t ts = tm + 384 + (kbg & 1) * 64;                       // TMEM operand slot: h0 32 columns | h1 32 columns
          const uint32_t d_main = tm + (uint32_t)((r & 1) * 128), d_corr = tm + 256;
          if (MODE == 1 || MODE == 3) {
            for (int k = 0; k < 4; ++k) cp_128x256b(ts + 8 * k, desc(a0 + 32 * k));
            for (int k = 0; k < 4; ++k) cp_128x256b(ts + 32 + 8 * k, desc(a1 + 32 * k));
          }
          if (MODE == 2) {
            for (int k = 0; k < 4; ++k) mma_ss(d_corr, desc(a1 + 32 * k), desc(b_hi + 32 * k), IDESC, 1);
            for (int k = 0; k < 4; ++k) mma_ss(d_corr, desc(a0 + 32 * k), desc(b_lo + 32 * k), IDESC, 1);
            for (int k = 0; k < 4; ++k) mma_ss(d_main, desc(a0 + 32 * k), desc(b_hi + 32 * k), IDESC, 1);
          } else if (MODE != 3) {
            for (int k = 0; k < 4; ++k) mma_ts(d_corr, ts + 32 + 8 * k, desc(b_hi + 32 * k), IDESC, 1);
            for (int k = 0; k < 4; ++k) mma_ts(d_corr, ts + 8 * k, desc(b_lo + 32 * k), IDESC, 1);
            for (int k = 0; k < 4; ++k) mma_ts(d_main, ts + 8 * k, desc(b_hi + 32 * k), IDESC, 1);
          }
        }
        commit2(smem_u32(&bar));
      }
      __syncwarp();
      mbar_wait(smem_u32(&bar), ph); ph ^= 1;
    }
    t1 = clock64();
    if (lane == 0) { if (blockIdx.x == 0) out[0] = t1 - t0; *(volatile int*)&stop = 1; }
  } else if (warp >= 4 && warp < 4 + bg) {
    float acc = 0.f; long long n = 0;
    const float4* p = reinterpret_cast<const float4*>(base + BG_OFF) + lane;
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
    if (lane == 0 && blockIdx.x == 0) out[1 + warp] = n * 512;
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads(); csync();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tm));
}

template <int MODE> void run(int bg) {
  static const char* names[] = {"TS  (A in TMEM)          ", "CPTS (8 cp + 12 TS MMAs)  ", "SS  (A via descriptors)   ", "CP only (8 cp)            "};
  long long* d; cudaMalloc(&d, 8 * 32); cudaMemset(d, 0, 8 * 32);
  float* sink; cudaMalloc(&sink, 4);
  auto fn = rate_kernel<MODE>;
  cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
  const int rounds = 200, per = 4;
  fn<<<148, 512, SMEM>>>(d, rounds, per, bg, sink);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[32]; cudaMemcpy(h, d, 8 * 32, cudaMemcpyDeviceToHost);
  long long bytes = 0; for (int i = 1; i < 32; ++i) bytes += h[i];
  printf("%s bg_warps=%d: %7.1f cycles per K block (12 MMAs = 852 at the TS rate), background LDS %5.1f B/clk (%s)\n", names[MODE], bg,
         (double)h[0] / (rounds * per), (double)bytes / (double)h[0], cudaGetErrorString(e));
  cudaFree(d); cudaFree(sink);
}

int main() {
  // layout check
  const int M = 256, N = 128, K = 64;
  __half *hA = (__half*)malloc(M * K * 2), *hB = (__half*)malloc(N * K * 2);
  for (int r = 0; r < M; ++r) for (int k = 0; k < K; ++k) hA[r * K + k] = __float2half((float)(((r * 7 + k * 3) % 13) - 6));
  for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) hB[n * K + k] = __float2half((float)(((n * 5 + k * 11) % 9) - 4));
  __half *dA, *dB; float* dO;
  cudaMalloc(&dA, M * K * 2); cudaMalloc(&dB, N * K * 2); cudaMalloc(&dO, 2 * M * N * 4);
  cudaMemcpy(dA, hA, M * K * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, N * K * 2, cudaMemcpyHostToDevice);
  cudaMemset(dO, 0xff, 2 * M * N * 4);
  cudaFuncSetAttribute(check_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
  check_kernel<<<2, 128, SMEM>>>(dA, dB, dO);
  cudaError_t e = cudaDeviceSynchronize();
  float* hO = (float*)malloc(2 * M * N * 4);
  cudaMemcpy(hO, dO, 2 * M * N * 4, cudaMemcpyDeviceToHost);
  for (int mode = 0; mode < 2; ++mode) {
    int bad = 0; double worst = 0;
    for (int r = 0; r < M; ++r) for (int n = 0; n < N; ++n) {
      double ref = 0; for (int k = 0; k < K; ++k) ref += (double)__half2float(hA[r * K + k]) * (double)__half2float(hB[n * K + k]);
      const double d = fabs(ref - (double)hO[(mode * M + r) * N + n]);
      if (!(d < 1e-3)) { if (bad < 3) printf("  mode %d mismatch r %d n %d: got %g want %g\n", mode, r, n, hO[(mode * M + r) * N + n], ref); ++bad; }
      if (d > worst) worst = d;
    }
    printf("layout check %s: %d mismatches of %d (max |diff| %.3g) (%s)\n", mode == 0 ? "SS     " : "cp + TS", bad, M * N, worst, cudaGetErrorString(e));
  }
  for (int bg : {0, 4}) run<0>(bg);
  for (int bg : {0, 4}) run<1>(bg);
  for (int bg : {0, 4}) run<2>(bg);
  for (int bg : {0, 4}) run<3>(bg);
  return 0;
}
