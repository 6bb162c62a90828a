// Microbenchmark (tools only): per-SM throughput of the FP32 FMA pipe and of the LEGACY warp-level tensor path
// (mma.sync m16n8k16 f16 / m16n8k8 tf32, SASS HMMA) on sm_100a.  Decides how the N >= 10 attention kernel does its
// per-channel score / output products.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rates pipe_rates.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

__global__ void ffma_kernel(float* out, int iters) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
  float x = out[0], y = out[1];
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], x, y);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  if (s == 123.456f) out[2] = s;
}

__global__ void hmma_f16_kernel(float* out, int iters) {
  unsigned a0 = threadIdx.x, a1 = 2, a2 = 3, a3 = 4, b0 = 5, b1 = 6;
  float c[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  if (s == 123.456f) out[2] = s;
}

__global__ void hmma_tf32_kernel(float* out, int iters) {
  unsigned a0 = threadIdx.x, a1 = 2, a2 = 3, a3 = 4, b0 = 5, b1 = 6;
  float c[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  if (s == 123.456f) out[2] = s;
}

template <typename F>
float time_ms(F f) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  f();
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  f();
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  return ms;
}

int main() {
  float* out;
  cudaMalloc(&out, 64);
  cudaMemset(out, 0, 64);
  int sms = 0, khz = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const int iters = 20000;
  for (int warps : {4, 8, 16, 32}) {
    const int threads = warps * 32;
    float t = time_ms([&] { ffma_kernel<<<sms, threads>>>(out, iters); });
    double fl = 2.0 * 16 * iters * (double)threads * sms;
    printf("FFMA  %2d warps/SM: %.3f ms  %.1f TFLOP/s  (%.1f FMA/clk/SM at %d MHz nominal)\n", warps, t, fl / t / 1e9, fl / 2 / (t * 1e-3) / sms / (khz * 1e3), khz / 1000);
    t = time_ms([&] { hmma_f16_kernel<<<sms, threads>>>(out, iters / 4); });
    fl = 2.0 * 16 * 8 * 16 * 8 * (iters / 4) * (double)warps * sms;
    printf("HMMA.f16 m16n8k16  %2d warps/SM: %.3f ms  %.1f TFLOP/s  (%.0f FLOP/clk/SM)\n", warps, t, fl / t / 1e9, fl / (t * 1e-3) / sms / (khz * 1e3));
    t = time_ms([&] { hmma_tf32_kernel<<<sms, threads>>>(out, iters / 4); });
    fl = 2.0 * 16 * 8 * 8 * 8 * (iters / 4) * (double)warps * sms;
    printf("HMMA.tf32 m16n8k8  %2d warps/SM: %.3f ms  %.1f TFLOP/s  (%.0f FLOP/clk/SM)\n", warps, t, fl / t / 1e9, fl / (t * 1e-3) / sms / (khz * 1e3));
  }
  return 0;
}
