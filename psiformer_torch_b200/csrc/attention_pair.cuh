// Attention on the (value, tangents, Laplacian) payload for 5 .. 14 electrons, head_dim 64: TWO warps per (walker, head).
//
// attention_payload_warp_kernel (attention.cuh) gives a unit to ONE warp that walks through score products, the
// reduce-scatter, the softmax-derivative arithmetic and the output products of a channel one after the other.  Its
// register tiles (a, b, y, cr: 250+ registers) and its 27 - 36 KiB of staged q / k / v per unit leave 6 - 8 warps per SM:
// 1.5 - 2 per scheduler, so every dependent phase boundary is exposed (ncu round 2: issue slots 63 % used, FMA pipe 37 %,
// 3.0 TB/s on Ne).  The staged data cannot shrink, but the WORK of a unit splits cleanly in two:
//
//   score warp   (a, b tiles; p, sum dv^2, cross terms)   channel c: scores -> reduce-scatter -> weights p~_c -> PT[c & 1]
//   output warp  (y, cr tiles)                            channel c: y_c = p~_c v_0 + p v_c, cr += p~_c v_c -> HBM, refill
//
// The score warp of channel c + 1 runs beside the output warp of channel c: the same staged bytes feed twice as many
// warps (16 per SM for N <= 10, 12 for N <= 14), each needing about half the registers, and neither repeats a load of
// the other.  Hand-over through two shared-memory weight tiles and four mbarriers per unit (PT full / PT empty x 2);
// each warp stages, and refills as soon as it has consumed them, the operands it reads itself (score warp: q, k;
// output warp: v), so a slot's next copies are in flight a whole channel of that warp's work before they are needed.
// Lane layouts, products and arithmetic are those of the one-warp kernel (attw_scores, attw_reduce_scatter, same
// summation order): results are bit-identical to it.
#pragma once
#include <cuda.h>

#include "attention.cuh"

namespace psif {

// One of q / k / v of a channel: 2 TI rows of ATTW_RS = 68 floats, rounded up to 128 bytes (TMA destinations).  A row is
// filled by ONE box row of 68 floats: the head's 64 columns plus 4 columns of padding that are never read (the next head's
// first columns, or TMA zero fill behind the row end), so a whole [N][68] array is a single tensor copy.
__host__ __device__ constexpr int attp_arr_floats(int TI) { return (2 * TI * ATTW_RS + 31) / 32 * 32; }
// q0 k0 v0 | ring of (qc kc vc) | p | p~ x 2 | 10 mbarriers: SQ[0 .. RING], SV[0 .. RING], PTF[2], PTE[2]  (+ pad to 128 bytes)
__host__ __device__ constexpr int attp_unit_floats(int TI) { return 3 * (1 + ATTW_RING) * attp_arr_floats(TI) + 3 * 256 + 32; }
__host__ __device__ constexpr int attp_units(int TI) {
  int u = (226 * 1024 - 128) / (attp_unit_floats(TI) * 4);
  const int cap = TI <= 5 ? 8 : 6;            // 128 registers per thread carry the TI <= 5 tiles, 170 the larger ones
  return u > cap ? cap : u;
}


// ---- waiting without splitting the warp -------------------------------------------------------------------------------
// `if (lane == 0) mbar_wait(..); __syncwarp();` is fatal in this kernel: when lane 0 blocks for long (here a role waits
// for the OTHER warp or for HBM, microseconds) the 31 lanes parked at the compiler's convergence barrier are released for
// forward progress, meet lane 0 again at the __syncwarp -- and from then on the two fragments run the rest of the role
// apart: ncu (round 2, first two versions) showed every FFMA issued twice, with 1 + 31 threads, and the kernel 2.4x
// slower than the one-warp kernel.  So ALL lanes poll, and leave the loop together on a vote; arrivals are made by all 32
// lanes on barriers initialised to 32.  Nothing in a role is executed by a single lane except the bulk-copy issue, a short
// loop with a known trip count that does reconverge.
__device__ __forceinline__ void attp_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  unsigned long long t0 = 0;
  while (true) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity), "r"(MBAR_SUSPEND_HINT_NS) : "memory");
    if (__all_sync(0xffffffffu, done != 0)) break;
    if ((++spins & 1023u) == 0) {  // a protocol bug must not hang the GPU box: give up after 4 s
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      if (__any_sync(0xffffffffu, now - t0 > 4000000000ull)) __trap();
    }
  }
}
__device__ __forceinline__ void attp_arrive(uint32_t bar) {      // all 32 lanes (release): the barrier counts 32
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// rows [0, N) of parts p0 .. p1 (q 0, k 1, v 2) of channel c -> dst: ONE 3-D tensor copy per part, box = 68 columns x 1
// channel x N tokens of the [tokens][C][3 d] view of qkv.  (As N row copies of 256 bytes per part, issued from a loop, the
// staging cost 12 % of the kernel's instructions: every cp.async.bulk drags ELECT / uniform-register moves along.)
// Each role stages, and later refills, only what it reads itself.
__device__ __forceinline__ void attp_issue(float* dst, int arr, uint32_t bar, const CUtensorMap* tm, int tok0, int N, int c, int d,
                                           int col, int p0, int p1) {
  __syncwarp();
  uint32_t elected;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(elected));
  if (elected) {
    mbar_arrive_expect_tx(bar, (uint32_t)((p1 - p0) * N * ATTW_RS * 4));
    for (int part = p0; part < p1; ++part)
      asm volatile(
          "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
          ::"r"(smem_u32(dst + part * arr)), "l"(tm), "r"(bar), "r"(part * d + col), "r"(c), "r"(tok0) : "memory");
  }
  __syncwarp();
}

template <int TI, bool PK>
__global__ void __launch_bounds__(attp_units(TI) * 64, 1)
attention_payload_pair_kernel(const __grid_constant__ CUtensorMap tmQKV, float* __restrict__ out, long long units, int N, int C, int d,
                              int H, unsigned* ovf) {
  extern __shared__ __align__(128) float smw_raw[];
  float* smw = smw_raw + ((128u - (smem_u32(smw_raw) & 127u)) & 127u) / 4;      // TMA destinations: 128-byte aligned
  constexpr int RS = ATTW_RS, ARR = attp_arr_floats(TI), TRI = 3 * ARR, U = attp_units(TI);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;    // warp-uniform for the compiler, too
  const int us = warp >> 1, role = warp & 1;                    // role 0: scores, 1: outputs
  const long long u = (long long)blockIdx.x * U + us;
  const bool live = u < units;
  const long long b = live ? u / H : 0;
  const int h = (int)(live ? u - b * H : 0);
  const long long tok0 = b * N;
  const int tk0 = (int)tok0;
  const int col = h * 64;
  float* base0 = smw + (size_t)us * attp_unit_floats(TI);       // q0 | k0 | v0
  float* ring = base0 + TRI;
  float* Pm = ring + ATTW_RING * TRI;                           // p  [j][16]: column j of p, rows ig * 8 + r
  float* PT = Pm + 256;                                         // p~ (tangent) / softmax Laplacian weights, two tiles
  const uint32_t bar0 = smem_u32(PT + 512);
  auto SQ = [&](int s) { return bar0 + 8u * s; };                      // q, k of channel 0 (s = 0) / ring slot s - 1 landed
  auto SV = [&](int s) { return bar0 + 8u * (1 + ATTW_RING + s); };    // v likewise
  auto PTF = [&](int t) { return bar0 + 8u * (2 + 2 * ATTW_RING + t); };   // weight tile t written (channel parity t)
  auto PTE = [&](int t) { return bar0 + 8u * (4 + 2 * ATTW_RING + t); };   // ... read
  const float scale = 0.125f;                                   // 1 / sqrt(64)
  const unsigned FULL = 0xffffffffu;
  // rows N .. 2 TI - 1 (an odd N) and the weight tiles start as zeros: padding contributes exact zeros.  The bulk copies
  // never write these locations, so the two proxies do not meet.
  if (role == 0) {
    if (N < 2 * TI)
      for (int i = lane; i < 3 * (1 + ATTW_RING) * RS; i += 32) base0[(i / RS) * ARR + N * RS + (i % RS)] = 0.f;
    for (int i = lane; i < 768; i += 32) Pm[i] = 0.f;
#pragma unroll
    for (int s = 0; s < 6 + 2 * ATTW_RING; ++s)   // SQ / SV: one arrival (+ the bytes); PTF / PTE: the 32 lanes of the signalling warp
      asm volatile(
          "{\n\t.reg .pred q;\n\t.reg .b32 l;\n\tmov.u32 l, %%laneid;\n\tsetp.eq.u32 q, l, 0;\n\t"
          "@q mbarrier.init.shared::cta.b64 [%0], %1;\n\t}" ::"r"(bar0 + 8u * s), "r"(s < 2 + 2 * ATTW_RING ? 1u : 32u) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (!live) return;

  if (role == 0) {
    // ================================================= score warp =====================================================
    const int eh = lane & 7, jg = (lane >> 3) & 1, ig = lane >> 4;
    const int qoff = ig * TI * RS + 2 * eh, koff = ARR + jg * TI * RS + 2 * eh;
    float p[TI], quad[TI];          // row eh of p and of sum_c dv^2, columns jg TI + j
    float bacc[TI][TI];             // cross terms sum_c q_c k_c^T (partial over this lane's column slice)
#pragma unroll
    for (int i = 0; i < TI; ++i) {
      quad[i] = 0.f;
#pragma unroll
      for (int j = 0; j < TI; ++j) bacc[i][j] = 0.f;
    }
    attp_issue(base0, ARR, SQ(0), &tmQKV, tk0, N, 0, d, col, 0, 2);
#pragma unroll
    for (int s = 0; s < ATTW_RING; ++s)
      if (1 + s < C) attp_issue(ring + s * TRI, ARR, SQ(1 + s), &tmQKV, tk0, N, 1 + s, d, col, 0, 2);
    attp_wait(SQ(0), 0);
    {
      float a[TI][TI];
#pragma unroll
      for (int i = 0; i < TI; ++i)
#pragma unroll
        for (int j = 0; j < TI; ++j) a[i][j] = 0.f;
      attw_scores<TI, 0>(base0 + qoff, base0 + koff, nullptr, nullptr, a, bacc);
      float s[TI];
      attw_reduce_scatter<TI>(a, eh, s);
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < TI; ++j) {
        s[j] = (jg * TI + j < N) ? s[j] * scale : -INFINITY;      // padding columns take no weight
        mx = fmaxf(mx, s[j]);
      }
      mx = fmaxf(mx, __shfl_xor_sync(FULL, mx, 8));
      float den = 0.f;
#pragma unroll
      for (int j = 0; j < TI; ++j) {
        s[j] = expf(s[j] - mx);
        den += s[j];
      }
      den += __shfl_xor_sync(FULL, den, 8);
      const float inv = 1.0f / den;
#pragma unroll
      for (int j = 0; j < TI; ++j) {
        p[j] = s[j] * inv;
        if (eh < TI) Pm[(jg * TI + j) * 16 + ig * 8 + eh] = p[j];
      }
    }
    attp_arrive(PTF(0));                          // channel 0: p is in shared memory (each lane releases its own writes)
    for (int c = 1; c < C; ++c) {
      const int slot = (c - 1) % ATTW_RING;
      attp_wait(SQ(1 + slot), (uint32_t)(((c - 1) / ATTW_RING) & 1));
      const float* cb = ring + slot * TRI;
      const bool lapc = c == C - 1;
      float a[TI][TI];
#pragma unroll
      for (int i = 0; i < TI; ++i)
#pragma unroll
        for (int j = 0; j < TI; ++j) a[i][j] = 0.f;
      if (!lapc) attw_scores<TI, 1>(base0 + qoff, base0 + koff, cb + qoff, cb + koff, a, bacc);
      else attw_scores<TI, 2>(base0 + qoff, base0 + koff, cb + qoff, cb + koff, a, bacc);
      // q_c, k_c are consumed: their slot takes channel c + RING now, a whole channel of this warp's work ahead of its use
      if (c + ATTW_RING < C) attp_issue(ring + slot * TRI, ARR, SQ(1 + slot), &tmQKV, tk0, N, c + ATTW_RING, d, col, 0, 2);
      float st[TI];
      attw_reduce_scatter<TI>(a, eh, st);
      float w[TI];                    // weight of v_0 in this channel's output: p~ (tangent) or the softmax Laplacian
      if (!lapc) {
        float m = 0.f;
#pragma unroll
        for (int j = 0; j < TI; ++j) {
          st[j] *= scale;
          m = fmaf(p[j], st[j], m);
        }
        m += __shfl_xor_sync(FULL, m, 8);
#pragma unroll
        for (int j = 0; j < TI; ++j) {
          const float dv = st[j] - m;
          w[j] = p[j] * dv;
          quad[j] = fmaf(dv, dv, quad[j]);
        }
      } else {
        float cross[TI];
        attw_reduce_scatter<TI>(bacc, eh, cross);
        float ma = 0.f, mq = 0.f;
#pragma unroll
        for (int j = 0; j < TI; ++j) {
          st[j] = st[j] * scale + 2.0f * scale * cross[j];
          ma = fmaf(p[j], st[j], ma);
          mq = fmaf(p[j], quad[j], mq);
        }
        ma += __shfl_xor_sync(FULL, ma, 8);
        mq += __shfl_xor_sync(FULL, mq, 8);
#pragma unroll
        for (int j = 0; j < TI; ++j) w[j] = p[j] * ((st[j] - ma) + quad[j] - mq);
      }
      const int t = c & 1;
      // the output warp has read tile t for channel c - 2 (its (c - 2) / 2-th use; channel 0 counts as a use of tile 0)
      if (c >= 2) attp_wait(PTE(t), (uint32_t)((((c - 2) >> 1)) & 1));
      float* PTm = PT + t * 256;
#pragma unroll
      for (int j = 0; j < TI; ++j)
        if (eh < TI) PTm[(jg * TI + j) * 16 + ig * 8 + eh] = w[j];
      __syncwarp();
      attp_arrive(PTF(t));
    }
  } else {
    // ================================================= output warp ====================================================
    const int eg = lane & 15, ig = lane >> 4;
    const int voff = 2 * ARR + 4 * eg;
    float* orow0 = out + ((tok0 + ig * TI) * C) * (long long)d;   // payload row (electron ig TI, channel 0)
    const long long rstep = (long long)C * d;
    const int ocol = col + 4 * eg;
    float amax = 0.f;
    float4 cr[TI];                  // sum_c p~_c v_c
#pragma unroll
    for (int i = 0; i < TI; ++i) cr[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    attp_issue(base0, ARR, SV(0), &tmQKV, tk0, N, 0, d, col, 2, 3);
#pragma unroll
    for (int s = 0; s < ATTW_RING; ++s)
      if (1 + s < C) attp_issue(ring + s * TRI, ARR, SV(1 + s), &tmQKV, tk0, N, 1 + s, d, col, 2, 3);
    attp_wait(SV(0), 0);
    attp_wait(PTF(0), 0);
    {
      float4 y[TI];
#pragma unroll
      for (int i = 0; i < TI; ++i) y[i] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
      for (int j = 0; j < 2 * TI; ++j) {
        const float4 v0j = *reinterpret_cast<const float4*>(base0 + voff + j * RS);
        const float4 pa = *reinterpret_cast<const float4*>(Pm + j * 16 + ig * 8);
        const float4 pb = *reinterpret_cast<const float4*>(Pm + j * 16 + ig * 8 + 4);
        const float pv[8] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w};
#pragma unroll
        for (int i = 0; i < TI; ++i) axpy4(y[i], pv[i], v0j);
      }
#pragma unroll
      for (int i = 0; i < TI; ++i)
        if (ig * TI + i < N) st_row4<PK>(orow0 + i * rstep, d, ocol, y[i], amax);
    }
    __syncwarp();
    attp_arrive(PTE(0));
    for (int c = 1; c < C; ++c) {
      const int slot = (c - 1) % ATTW_RING, t = c & 1;
      attp_wait(PTF(t), (uint32_t)((c >> 1) & 1));
      attp_wait(SV(1 + slot), (uint32_t)(((c - 1) / ATTW_RING) & 1));
      const float* cb = ring + slot * TRI;
      const float* PTm = PT + t * 256;
      const bool lapc = c == C - 1;
      float4 y[TI];
#pragma unroll
      for (int i = 0; i < TI; ++i)
        y[i] = lapc ? make_float4(2.0f * cr[i].x, 2.0f * cr[i].y, 2.0f * cr[i].z, 2.0f * cr[i].w) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
      for (int j = 0; j < 2 * TI; ++j) {
        const float4 v0j = *reinterpret_cast<const float4*>(base0 + voff + j * RS);
        const float4 vcj = *reinterpret_cast<const float4*>(cb + voff + j * RS);
        const float4 pa = *reinterpret_cast<const float4*>(Pm + j * 16 + ig * 8);
        const float4 pb = *reinterpret_cast<const float4*>(Pm + j * 16 + ig * 8 + 4);
        const float4 wa = *reinterpret_cast<const float4*>(PTm + j * 16 + ig * 8);
        const float4 wb = *reinterpret_cast<const float4*>(PTm + j * 16 + ig * 8 + 4);
        const float pv[8] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w};
        const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
        for (int i = 0; i < TI; ++i) axpy4(y[i], wv[i], v0j);
        if (!lapc) {
#pragma unroll
          for (int i = 0; i < TI; ++i) axpy4(cr[i], wv[i], vcj);
        }
#pragma unroll
        for (int i = 0; i < TI; ++i) axpy4(y[i], pv[i], vcj);
      }
      // past the last read of this channel's weight tile and v_c: hand the tile back and refill the v slot BEFORE the
      // stores, so that the copies fly while they drain
      attp_arrive(PTE(t));
      if (c + ATTW_RING < C) attp_issue(ring + slot * TRI, ARR, SV(1 + slot), &tmQKV, tk0, N, c + ATTW_RING, d, col, 2, 3);
#pragma unroll
      for (int i = 0; i < TI; ++i)
        if (ig * TI + i < N) st_row4<PK>(orow0 + i * rstep + (long long)c * d, d, ocol, y[i], amax);
    }
    if (PK) raise_range_flag(ovf, amax);
  }
}

typedef CUresult (*PFN_attpEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
// the driver entry point, looked up once per process (immutable afterwards)
inline PFN_attpEncodeTiled attp_encode_fn() {
  static const PFN_attpEncodeTiled fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<PFN_attpEncodeTiled>(p);
  }();
  return fn;
}

inline int32_t attention_pair_launch(const float* qkv, float* out, long long units, int N, int C, int d, int H, cudaStream_t st,
                                     bool packed, unsigned* ovf) {
  DevSmemCfg& cfg = dev_smem_cfg();
  const int TI = (N + 1) / 2;
  const long long tokens = units / H * N;
  PFN_attpEncodeTiled enc = attp_encode_fn();
  if (!enc) return fail(PSIF_E_CUDA, "cuTensorMapEncodeTiled entry point not available%s");
  if (tokens > 0x7fffffffLL) return fail(PSIF_E_INVALID, "attention: too many tokens for one launch%s");
  CUtensorMap tm;
  {
    // qkv as [tokens][C][3 d] fp32; box = 68 columns (the head's 64 + the padding of a staged row) x 1 channel x N tokens
    cuuint64_t dims[3] = {(cuuint64_t)(3 * d), (cuuint64_t)C, (cuuint64_t)tokens};
    cuuint64_t strides[2] = {(cuuint64_t)(3 * d) * 4, (cuuint64_t)C * (3 * d) * 4};
    cuuint32_t box[3] = {(cuuint32_t)ATTW_RS, 1u, (cuuint32_t)N};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(qkv), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(PSIF_E_CUDA, "cuTensorMapEncodeTiled (attention) failed (%s%lld)", "", (long long)r);
  }
#define PSIF_ATTP(T)                                                                                                              \
  case T: {                                                                                                                       \
    constexpr int U = attp_units(T);                                                                                              \
    constexpr size_t smem = (size_t)U * attp_unit_floats(T) * sizeof(float) + 128;                                                \
    if (cfg.attp[T] == 0) {                                                                                                       \
      PSIF_CUDA_CHECK(cudaFuncSetAttribute((attention_payload_pair_kernel<T, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      PSIF_CUDA_CHECK(cudaFuncSetAttribute((attention_payload_pair_kernel<T, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));  \
      cfg.attp[T] = smem;                                                                                                         \
    }                                                                                                                             \
    const long long nb = (units + U - 1) / U;                                                                                     \
    if (nb > 0x7fffffffLL) return fail(PSIF_E_INVALID, "attention: grid too large%s");                                            \
    if (packed) PSIF_LAUNCH((attention_payload_pair_kernel<T, true>), (unsigned)nb, U * 64, smem, st, tm, out, units, N, C, d, H, ovf);     \
    else PSIF_LAUNCH((attention_payload_pair_kernel<T, false>), (unsigned)nb, U * 64, smem, st, tm, out, units, N, C, d, H, ovf);           \
    return PSIF_OK;                                                                                                               \
  }
  switch (TI) { PSIF_ATTP(3) PSIF_ATTP(4) PSIF_ATTP(5) PSIF_ATTP(6) PSIF_ATTP(7) }
#undef PSIF_ATTP
  return fail(PSIF_E_INVALID, "attention: shape not handled by the two-warp kernel%s");
}

}  // namespace psif
