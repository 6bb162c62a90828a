// Shared device/host helpers for libpsiformer_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <atomic>
#include <cstdio>
#include <cstring>

#include "../../include/psiformer_b200.h"

namespace psif {

// ---- thread-local error string + launch counter -----------------------------------------
extern thread_local char g_err[512];
extern std::atomic<long long> g_launches;

inline int32_t fail(int32_t code, const char* fmt, const char* a = "", long long b = 0, long long c = 0) {
  snprintf(g_err, sizeof(g_err), fmt, a, b, c);
  return code;
}

#define PSIF_CUDA_CHECK(expr)                                                              \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      snprintf(psif::g_err, sizeof(psif::g_err), "%s failed: %s (%s:%d)", #expr,          \
               cudaGetErrorString(_e), __FILE__, __LINE__);                                \
      return PSIF_E_CUDA;                                                                  \
    }                                                                                      \
  } while (0)

// every kernel launch in the library goes through this macro: it counts the launch
// (bench.py's gpu_launches) and turns launch-configuration errors into return codes.
#define PSIF_LAUNCH(kernel, grid, block, smem, stream, ...)                                \
  do {                                                                                     \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                            \
    psif::g_launches.fetch_add(1, std::memory_order_relaxed);                              \
    cudaError_t _e = cudaGetLastError();                                                   \
    if (_e != cudaSuccess) {                                                               \
      snprintf(psif::g_err, sizeof(psif::g_err), "launch of %s failed: %s (%s:%d)",       \
               #kernel, cudaGetErrorString(_e), __FILE__, __LINE__);                       \
      return PSIF_E_CUDA;                                                                  \
    }                                                                                      \
  } while (0)

#define PSIF_TRY(expr)                 \
  do {                                 \
    int32_t _r = (expr);               \
    if (_r != PSIF_OK) return _r;      \
  } while (0)

// ---- constants of the reference ---------------------------------------
constexpr double kDetJitter = 1e-4;     // logdet_matmul.py:18
constexpr double kMinSingular = 1e-6;   // logdet_matmul.py:16
constexpr double kOutputFloor = 1e-12;  // logdet_matmul.py:17
constexpr float kLnEps = 1e-5f;         // nn.LayerNorm default, psiformer.py:86-87
constexpr double kCoulombEps = 1e-5;    // hamiltonian.py:16
constexpr double kJastrowEps = 1e-12;   // jastrow.py:34,60

// ---- warp helpers ------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

#ifdef __CUDACC__
// ---- mbarrier helpers (tensor-core GEMM pipeline, bulk-copy rings of the attention kernel) ---------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// try_wait's suspend-time hint: the waiting thread is parked in hardware until the phase flips (or this many ns pass)
// instead of spinning.  ncu (source page, round 1): without it the polling loops of the waiting roles made up more
// than half of all executed warp instructions of the cta_group::2 GEMM, which had become issue bound.
constexpr uint32_t MBAR_SUSPEND_HINT_NS = 0x989680u;
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  uint32_t spins = 0;
  unsigned long long t0 = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity), "r"(MBAR_SUSPEND_HINT_NS)
        : "memory");
    if (done) break;
    if ((++spins & 1023u) == 0) {  // a protocol bug must not hang the GPU box: give up after 4 s
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ull) __trap();
    }
  }
}
// non-blocking test; the result can be consumed much later, which hides the ~250-cycle latency every mbarrier
// operation has while the tensor core and TMA keep shared memory busy
__device__ __forceinline__ uint32_t mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  return done;
}
// Warp-wide wait: ONE lane polls, the rest of the warp joins through __syncwarp (which orders memory among the
// participating lanes).  32 lanes polling the same mbarrier serialise in the shared-memory sync unit: the clock64
// timeline of round 1 showed ~430 cycles for a try_wait on a barrier that had completed long before.
__device__ __forceinline__ void mbar_wait_warp(uint32_t bar, uint32_t parity) {
  if ((threadIdx.x & 31) == 0) mbar_wait(bar, parity);
  __syncwarp();
}
#endif

// GELU(tanh) and its first two derivatives (psiformer.py:70).
// tanh through E = exp(-2 |inner|) <= 1:  tanh = +-(1 - E) / (1 + E),  1 + tanh = 2 / (1 + E) or 2 E / (1 + E),
// sech^2 = 4 E / (1 + E)^2 -- no cancellation in 1 + tanh and 1 - tanh^2 for large |u| (where tanhf-based code loses
// digits), and ~12 instructions (MUFU.EX2 + MUFU.RCP) instead of ~50: the payload-GELU epilogue of the FC GEMM is
// instruction bound.  Absolute error <= 7e-7 in g, 2e-7 in g' and g'' over |u| <= 12 (checked against fp64 autograd).
__device__ __forceinline__ void gelu_tanh_parts(float inner, float& t, float& one_plus_t, float& sech2) {
  float E, r;
  const float a = -2.8853900817779268f * fabsf(inner);          // -2 |inner| log2(e)
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(E) : "f"(a));
  const float den = 1.0f + E;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(den));
  const float tp = (1.0f - E) * r;
  const bool pos = inner >= 0.0f;
  t = pos ? tp : -tp;
  one_plus_t = pos ? 2.0f * r : 2.0f * E * r;
  sech2 = 4.0f * E * r * r;
}
__device__ __forceinline__ void gelu_tanh_d2(float u, float& g, float& g1, float& g2) {
  const float kap = 0.7978845608028654f;  // sqrt(2/pi)
  const float c3 = 0.044715f;
  const float u2 = u * u;
  float t, opt, sech2;
  gelu_tanh_parts(kap * (u + c3 * u * u2), t, opt, sech2);
  const float q = kap * (1.0f + 3.0f * c3 * u2);
  const float hs = 0.5f * u * sech2;
  g = 0.5f * u * opt;
  g1 = fmaf(hs, q, 0.5f * opt);
  g2 = fmaf(sech2, q, hs * (kap * 6.0f * c3 * u - 2.0f * t * q * q));
}
__device__ __forceinline__ float gelu_tanh(float u) {
  const float kap = 0.7978845608028654f;
  const float u2 = u * u;                                         // the same expression tree as gelu_tanh_d2: bit-identical values
  float t, opt, sech2;
  gelu_tanh_parts(kap * (u + 0.044715f * u * u2), t, opt, sech2);
  return 0.5f * u * opt;
}

inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- packed fp16 pair: the A-operand format of the fp16-split GEMM, written by its PRODUCERS -----------------------
// x = h0 + 2^-11 h1, h0 = fp16(x), h1 = fp16(2^11 (x - h0)) (gemm_tcgen05.cuh).  A payload row of w fp32 values is
// stored in the same 4 w bytes as two fp16 planes, [h0[0..w) | h1[0..w)], so a TMA box of 64 fp16 columns of either
// plane is directly a K-major tensor-core operand tile: the GEMM needs no conversion pass over its A operand.
// amax tracks the largest |x| packed (fp16 tops out at 65504: the producer raises the handle's range flag).
constexpr float kSplitLoScale = 2048.f;
#ifdef __CUDACC__
__device__ __forceinline__ void pack_split4(const float4 v, uint2& h0, uint2& h1, float& amax) {
  const __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
  const float2 fa = __half22float2(a), fb = __half22float2(b);
  const __half2 la = __floats2half2_rn((v.x - fa.x) * kSplitLoScale, (v.y - fa.y) * kSplitLoScale);
  const __half2 lb = __floats2half2_rn((v.z - fb.x) * kSplitLoScale, (v.w - fb.y) * kSplitLoScale);
  h0.x = *reinterpret_cast<const uint32_t*>(&a); h0.y = *reinterpret_cast<const uint32_t*>(&b);
  h1.x = *reinterpret_cast<const uint32_t*>(&la); h1.y = *reinterpret_cast<const uint32_t*>(&lb);
  amax = fmaxf(fmaxf(amax, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
}
// store 4 consecutive values at column `col` of the payload row starting at `row` (w columns): fp32, or packed
template <bool PK>
__device__ __forceinline__ void st_row4(float* row, int w, int col, const float4 v, float& amax) {
  if constexpr (PK) {
    uint2 h0, h1;
    pack_split4(v, h0, h1, amax);
    __half* rh = reinterpret_cast<__half*>(row);
    *reinterpret_cast<uint2*>(rh + col) = h0;
    *reinterpret_cast<uint2*>(rh + w + col) = h1;
  } else {
    *reinterpret_cast<float4*>(row + col) = v;
  }
}
__device__ __forceinline__ void raise_range_flag(unsigned* ovf, float amax) {
  if (ovf != nullptr && !(amax < 65504.f)) atomicOr(ovf, 1u);      // also catches NaN / inf
}
#endif

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a per-DEVICE setting: the launchers remember, per device, how
// much each kernel has been granted so far.  Entries only ever grow and setting an attribute twice is harmless, so
// two host threads driving the same device at worst repeat a call.
struct DevSmemCfg {
  size_t att1 = 0, att2 = 0, attw[8] = {}, attp[8] = {}, det[9][2] = {}, detfix[9] = {}, afl = 0;
  bool att4 = false;
};
inline DevSmemCfg& dev_smem_cfg() {
  static DevSmemCfg table[64];
  int dev = 0;
  cudaGetDevice(&dev);
  return table[dev & 63];
}

}  // namespace psif
