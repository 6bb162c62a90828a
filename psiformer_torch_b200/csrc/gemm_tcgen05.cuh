// Split-precision GEMM on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM):
//     Y[M][N] = X[M][K] * W[N][K]^T  (+ bias on rows r % C == 0) (+ residual) (GELU if act)
//
// One kernel: tc_gemm_2cta_kernel (cta_group::2 pairs; fp16-split operands when HST > 0, tf32-split operands when
// HST = 0 = psif_set_gemm_mode(PSIF_GEMM_TF32_SPLIT)), epilogue through TMA tensor stores / reduce-adds, fused payload
// GELU.  See the comment block in front of it and DESIGN.md, "Tensor-core GEMM".  The two one-CTA-per-tile kernels of
// round 1 live in tools/legacy/ and are no longer compiled.
//
// fp32-grade accuracy from 11-bit-significand tensor-core inputs by splitting both operands, x = x_hi + x_lo, and
// issuing three MMAs per K slice,
//     D_main += X_hi*W_hi,   D_corr += X_lo*W_hi + X_hi*W_lo     (the dropped X_lo*W_lo term is ~2^-22 relative).
// The tensor core truncates when it adds into the fp32 accumulator, a bias that grows with the number of
// accumulation steps; keeping the 2^-11-sized correction products in their OWN accumulator leaves a third of the steps
// on the main one (measured 3x smaller error), and the two are added (round-to-nearest) in the epilogue.
// The weight halves are prepared once per psif_set_params; X is split on the fly (splitter warps -> TMEM).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>

#include <cstdlib>
#include <map>
#include <tuple>

#include "common.cuh"

namespace psif {

constexpr int TC_BM = 128, TC_BK = 32;
constexpr int TC_A_BYTES = TC_BM * TC_BK * 4;  // 16 KiB

// ---- PTX wrappers (the mbarrier helpers live in common.cuh) -------------------------------------------
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// one elected lane of a fully converged warp.  Issuing tcgen05.mma / TMA / commit under `if (elect_one())` lets
// ptxas keep descriptors in uniform registers; under `if (lane == 0)` it wraps EVERY such instruction in an
// ELECT / BRA.U.ANY loop (ncu source page, round 1: ~100 issue cycles per MMA, i.e. the kernel was issue bound).
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred;
}
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
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

// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 format, version 1): 8-row groups of 1024 bytes
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);   // start address            bits [0,14)
  d |= (uint64_t)1 << 16;                     // leading byte offset (unused for swizzled K-major) bits [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;           // stride byte offset: 8 rows * 128 B   bits [32,46)
  d |= (uint64_t)1 << 46;                     // descriptor version (Blackwell)         bits [46,48)
  d |= (uint64_t)2 << 61;                     // SWIZZLE_128B                           bits [61,64)
  return d;
}

// instruction descriptor: D=f32, A=B=tf32, both K-major, M=128, N=256
__host__ __device__ constexpr uint32_t tc_idesc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}


constexpr int TS_BN = 128;   // output columns per tile
__device__ __forceinline__ void tc_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tc_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tc_ld32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_id_x() { uint32_t r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_count_x() { uint32_t r; asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void st_global_v8(float* p, const float (&v)[8]) {
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
               "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
}
__device__ __forceinline__ void st_global_v4_if(float* p, const float4 v, bool pred) {
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %5, 0;\n\t@q st.global.v4.f32 [%0], {%1, %2, %3, %4};\n\t}"
      ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"((int)pred) : "memory");
}
__device__ __forceinline__ void st_global_u4_if(void* p, const uint4 v, bool pred) {
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %5, 0;\n\t@q st.global.v4.b32 [%0], {%1, %2, %3, %4};\n\t}"
      ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"((int)pred) : "memory");
}
__device__ __forceinline__ void ld_global_v8(const float* p, float (&v)[8]) {
  asm volatile("ld.global.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]),
               "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(p) : "memory");
}

// ------------------------------------------------------------------------------------------------
// 2-CTA variant (tcgen05 cta_group::2): the two CTAs of a cluster (one TPC) compute a 256 x 128 tile with ONE MMA
// instruction stream issued by the leader CTA.  Each CTA owns 128 rows of X (its own TMA ring, splitter, TMEM
// operand slots, accumulators and epilogue) but stages only HALF of every weight tile (64 of the 128 rows of W_hi
// and W_lo): the tensor cores of the pair exchange the halves.  Per CTA and K block this halves both the weight
// bytes written by TMA (32 -> 16 KiB) and the weight bytes the tensor core reads from shared memory (48 -> 24 KiB),
// and a stage shrinks from 48 to 32 KiB, so the ring is 6 deep instead of 4.
//   leader-side barriers  : FULL_B (both CTAs' weight halves, 2-SM TMA signals the leader), SPLIT (8 splitter
//                            warps of the pair, remote arrive), TEMPTY (16 epilogue warps of the pair)
//   per-CTA barriers      : FULL_X (own X tile), and EMPTY_S / EMPTY_A / TFULL which the leader's tcgen05.commit
//                            multicasts to both CTAs.
// ------------------------------------------------------------------------------------------------
// payload-GELU epilogue staging (act == 2): two [128 rows][68 floats] half tiles + 8 x 192 floats of per-warp scratch
constexpr int GELU_STAGE_STRIDE = 68;
constexpr int GELU_STAGE_BYTES = 2 * TC_BM * GELU_STAGE_STRIDE * 4 + 8 * 192 * 4;
constexpr int T2_STAGES = 4, T2_THREADS = 640;   // warps 0-3 TMA / MMA / TMEM alloc, 4-7 + 16-19 splitter, 8-15 epilogue
constexpr int T2_BH_BYTES = (TS_BN / 2) * TC_BK * 4;                  // 8 KiB: this CTA's half of one weight tile
constexpr int T2_STAGE_BYTES = TC_A_BYTES + 2 * T2_BH_BYTES;          // 32 KiB
constexpr int T2_SMEM_BYTES = T2_STAGES * T2_STAGE_BYTES + 1024 + 1024;
constexpr int T2_SMEM_BYTES_GELU = T2_SMEM_BYTES + GELU_STAGE_BYTES;   // + the payload-GELU staging tiles (act == 2)

__device__ __forceinline__ uint32_t mapa_rank(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  // default semantics (release at CTA scope), as CUTLASS' ClusterBarrier::arrive(cta_id) does: a cluster-scope release
  // costs ~1300 cycles here (tools/trace_gemm2.py), and what the consumer needs ordered are tcgen05 operations, which
  // tcgen05.wait / tcgen05.fence::before_thread_sync take care of
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {   // acquire at cluster scope
  uint32_t done = 0, spins = 0;
  unsigned long long t0 = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity), "r"(MBAR_SUSPEND_HINT_NS) : "memory");
    if (done) break;
    if ((++spins & 1023u) == 0) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ull) __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t leader_bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_commit_2sm(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32_ts_2sm(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// fp16-split mode of the same kernel (HST = ring depth, 0 = the tf32 mode above).  fp16 has the 11-bit significand of
// tf32 but kind::f16 MMAs take K = 16 per instruction at the same 71 cycles, i.e. twice the FLOPs per tensor cycle:
//     x = h0 + 2^-11 h1,  h0 = fp16(x),  h1 = fp16(2^11 (x - h0))        (both operands, weights once per set_params)
//     D_main += X_h0 W_h0,   D_corr += X_h1 W_h0 + X_h0 W_h1,   Y = D_main + 2^-11 D_corr.
// Scaling the low part keeps it a NORMAL fp16 number whenever x is one (unscaled it would be subnormal for |x| < 0.12),
// so the split carries 22 significand bits like the tf32 one; below fp16's normal range (|x| < 6e-5) the absolute error
// is <= 2^-25 / 2^11 = 1.5e-11.  fp16 tops out at 65504: a larger activation would become inf, so the splitter tracks
// max |x| and raises *ovf; the host side then repeats the call in tf32 mode (never seen on sampled walkers).
// A K block is 64 columns: two 128-byte-swizzled fp32 boxes of X (32 KiB) and one 128-byte-swizzled fp16 box of each
// weight half (64 rows x 128 B = 8 KiB), still 12 MMAs per K block, so every hand-off of the pipeline is amortised
// over twice the work; TMEM operand slots keep their 64 columns (32 of packed h0, 32 of packed h1).
constexpr int H_BK = 64;
constexpr int H_X_BYTES = 2 * TC_A_BYTES;                         // 32 KiB
constexpr int H_STAGE_BYTES = H_X_BYTES + 2 * T2_BH_BYTES;        // 48 KiB
constexpr float H_LO_SCALE = 2048.f;
// + the epilogue's staging: 8 warps x 4 KiB (plain) or two 128 x 64 fp32 tiles (payload GELU)
constexpr int h_smem_bytes(int nst, bool gelu) { return nst * H_STAGE_BYTES + 1024 + 1024 + (gelu ? GELU_STAGE_BYTES : 8 * 4096); }
// instruction descriptor: D = f32, A = B = f16, both K-major
__host__ __device__ constexpr uint32_t tc_idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void tc_mma_f16_ts_2sm(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Four MMAs (one K block's main-accumulator group) with two non-blocking mbarrier tests issued IN FRONT of them and
// their predicates read BEHIND them, in one asm block.  The tensor core's instruction queue is shallow: a
// tcgen05.mma issue blocks until a slot frees, so the issuing thread spends about as long on 12 issues as the MMAs
// take to execute (852 cycles), and two blocking barrier waits per K block (~300 cycles each while shared memory is
// busy) came on top of that: 1440 cycles per K block on the clock64 timeline.  Issued this way the barrier round trips
// overlap the MMA issues; the answers (next K block's operands ready?) are looked at one K block later.
template <bool F16>
__device__ __forceinline__ void tc_mma4_test2_2sm(uint32_t d, uint32_t a, uint64_t b0, uint64_t b1, uint64_t b2, uint64_t b3,
                                                  uint32_t idesc, uint32_t acc0, uint32_t bar1, uint32_t par1, uint32_t bar2,
                                                  uint32_t par2, uint32_t& r1, uint32_t& r2) {
  if constexpr (F16)
    asm volatile(
        "{\n\t.reg .pred p, q1, q2;\n\t.reg .b32 a1, a2, a3;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 q1, [%10], %11;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 q2, [%12], %13;\n\t"
        "add.u32 a1, %3, 8;\n\tadd.u32 a2, %3, 16;\n\tadd.u32 a3, %3, 24;\n\t"
        "setp.ne.b32 p, %9, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%2], [%3], %4, %8, p;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%2], [a1], %5, %8, 1;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%2], [a2], %6, %8, 1;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%2], [a3], %7, %8, 1;\n\t"
        "selp.u32 %0, 1, 0, q1;\n\tselp.u32 %1, 1, 0, q2;\n\t}"
        : "=r"(r1), "=r"(r2)
        : "r"(d), "r"(a), "l"(b0), "l"(b1), "l"(b2), "l"(b3), "r"(idesc), "r"(acc0), "r"(bar1), "r"(par1), "r"(bar2), "r"(par2)
        : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p, q1, q2;\n\t.reg .b32 a1, a2, a3;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 q1, [%10], %11;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 q2, [%12], %13;\n\t"
        "add.u32 a1, %3, 8;\n\tadd.u32 a2, %3, 16;\n\tadd.u32 a3, %3, 24;\n\t"
        "setp.ne.b32 p, %9, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%2], [%3], %4, %8, p;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%2], [a1], %5, %8, 1;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%2], [a2], %6, %8, 1;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%2], [a3], %7, %8, 1;\n\t"
        "selp.u32 %0, 1, 0, q1;\n\tselp.u32 %1, 1, 0, q2;\n\t}"
        : "=r"(r1), "=r"(r2)
        : "r"(d), "r"(a), "l"(b0), "l"(b1), "l"(b2), "l"(b3), "r"(idesc), "r"(acc0), "r"(bar1), "r"(par1), "r"(bar2), "r"(par2)
        : "memory");
}

// PK (fp16 mode only): the A operand arrives ALREADY split, as the packed fp16 pair its producer wrote (common.cuh:
// LayerNorm, attention, the payload-GELU epilogue below).  tmX is then an fp16 map over the [rows][2 Ktot] matrix; the
// two boxes of a K block are 64 columns of the h0 plane and the same 64 columns of the h1 plane (a_h1_col = Ktot further
// right), and the "splitter" warps only move them from shared memory to TMEM: 8 LDS.128 + 2 tcgen05.st per thread and
// K block instead of ~150 conversion instructions.  With PK the payload-GELU epilogue (act == 2) writes packed output.
template <int NMAIN, int HST, bool PK = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(T2_THREADS, 1)
tc_gemm_2cta_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmWhi,
                    const __grid_constant__ CUtensorMap tmWlo, const float* __restrict__ bias, const float* res, float* Y,
                    long long M, int N, int K, int C, int act, long long* trace, int rpt, int dbg, unsigned* ovf,
                    const __grid_constant__ CUtensorMap tmY, int tma_out, int a_h1_col) {
  constexpr bool F16 = HST > 0;
  static_assert(!F16 || NMAIN == 1, "fp16 mode runs K passes of <= 512 columns with one main accumulator");
  static_assert(!PK || F16, "a packed A operand is an fp16 pair");
  constexpr int BK = F16 ? H_BK : TC_BK;                       // K columns per K block
  constexpr int X_BYTES = F16 ? H_X_BYTES : TC_A_BYTES;
  constexpr int STAGE_BYTES = F16 ? H_STAGE_BYTES : T2_STAGE_BYTES;
  constexpr int RING_BYTES = F16 ? HST * H_STAGE_BYTES : T2_STAGES * T2_STAGE_BYTES;
  constexpr int ACC_COLS = (NMAIN + 1) * TS_BN;
  constexpr int TA_STAGES = (512 - ACC_COLS) / 64;
  // shared-memory ring depth.  With a ring no deeper than the TMEM operand ring, FULL_X of K block kb (whose TMA was
  // issued after the commit of K block kb - NST) already implies that TMEM slot kb % TA_STAGES has been consumed, so
  // the splitter needs no EMPTY_A barrier at all: one barrier operation less per K block on its critical path.
  constexpr int NST = F16 ? HST : (NMAIN == 1 ? 4 : T2_STAGES);
  constexpr bool ELIDE_A = TA_STAGES >= NST;
  extern __shared__ uint8_t tc_smem_raw[];
  // round up to 1024 bytes by pointer arithmetic on the __shared__ array (not through an integer cast): the compiler
  // then knows these are shared-memory accesses (LDS / STS instead of generic LD / ST) that cannot alias Y or res
  uint8_t* base = tc_smem_raw + ((1024u - (smem_u32(tc_smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + RING_BYTES);
  const uint32_t bar0 = smem_u32(bars);
  auto FULL_X = [&](int s) { return bar0 + 8u * s; };            // per CTA
  auto FULL_B = [&](int s) { return bar0 + 8u * (8 + s); };      // used in the leader
  auto EMPTY_S = [&](int s) { return bar0 + 8u * (16 + s); };    // per CTA (multicast commit)
  auto SPLIT = [&](int a) { return bar0 + 8u * (24 + a); };      // used in the leader
  auto EMPTY_A = [&](int a) { return bar0 + 8u * (28 + a); };    // per CTA (multicast commit)
  // TFULL per CTA (multicast commit); CEMPTY / TEMPTY (correction / main accumulators drained) live in the leader
  // CFULL: the correction accumulator is complete (committed before the last 4 main MMAs of a tile), so its drain
  // overlaps them
  const uint32_t TFULL = bar0 + 8u * 32, TEMPTY = bar0 + 8u * 33, CEMPTY = bar0 + 8u * 34, CFULL = bar0 + 8u * 35;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 36);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  const bool leader = crank == 0;
  // optional timeline of cluster 0 (tools only): trace[(cta * 11 + role) * 512 + i] = clock64 at event i
  const bool tracing = trace != nullptr && blockIdx.x < 2;
  long long* tr = trace + (long long)blockIdx.x * 18 * 512;
  int tcount = 0;
#define PSIF_TRACE2(role) do { if (tracing && tcount < 512) tr[(role) * 512 + tcount] = clock64(); } while (0)
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmWhi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmWlo) : "memory");
    if (tma_out) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmY) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < NST; ++s) { mbar_init(FULL_X(s), 1); mbar_init(FULL_B(s), 1); mbar_init(EMPTY_S(s), 1); }
    for (int a = 0; a < TA_STAGES; ++a) { mbar_init(SPLIT(a), 16); mbar_init(EMPTY_A(a), 1); }
    mbar_init(TFULL, 1);
    mbar_init(CFULL, 1);
    mbar_init(TEMPTY, 16);
    mbar_init(CEMPTY, 16);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // 640 threads start with 96 registers each; the epilogue holds a 32 x 64 accumulator slab per warp twice over
  // (correction + main) and takes what the single-lane roles and the splitters do not need: 56 + 2 x 80 + 2 x 128 <= 480
  // (setmaxnreg at the top of each warpgroup's branch below)

  const int tiles_n = (N + TS_BN - 1) / TS_BN;                    // a ragged last column tile: W rows >= N are TMA zero fill
  const long long tiles_m = (M + rpt - 1) / rpt;
  const long long groups = ((tiles_m + 1) / 2) * tiles_n;        // a group = 2 row tiles x 1 column tile
  const int nkb = K / BK;
  const uint32_t smem_base = smem_u32(base);
  const long long g0 = (long long)cluster_id_x(), gstep = (long long)cluster_count_x();

  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t full_b_leader0 = mapa_rank(FULL_B(0), 0);
      for (long long grp = g0; grp < groups; grp += gstep) {
        const int m0 = (int)(((grp / tiles_n) * 2 + crank) * rpt), n0 = (int)(grp % tiles_n) * TS_BN + (int)crank * (TS_BN / 2);
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(EMPTY_S(stage), phase ^ 1);
          PSIF_TRACE2(0);
          const uint32_t sa = smem_base + stage * STAGE_BYTES;
          mbar_arrive_expect_tx(FULL_X(stage), X_BYTES);
          tma_load_2d(sa, &tmX, kb * BK, m0, FULL_X(stage));
          if (F16) tma_load_2d(sa + TC_A_BYTES, &tmX, PK ? a_h1_col + kb * BK : kb * BK + TC_BK, m0, FULL_X(stage));
          if (leader) mbar_arrive_expect_tx(FULL_B(stage), 4 * T2_BH_BYTES);   // both halves of W_hi and W_lo
          const uint32_t lb = full_b_leader0 + 8u * stage;
          tma_load_2d_2sm(sa + X_BYTES, &tmWhi, kb * BK, n0, lb);
          tma_load_2d_2sm(sa + X_BYTES + T2_BH_BYTES, &tmWlo, kb * BK, n0, lb);
          PSIF_TRACE2(1);
          ++tcount;
          if (++stage == NST) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = F16 ? tc_idesc_f16(2 * TC_BM, TS_BN) : tc_idesc(2 * TC_BM, TS_BN);
    auto mma = [](uint32_t d, uint32_t a, uint64_t b, uint32_t id, uint32_t accum) {
      if constexpr (F16) tc_mma_f16_ts_2sm(d, a, b, id, accum);
      else tc_mma_tf32_ts_2sm(d, a, b, id, accum);
    };
    if (leader && elect_one()) {
      int stage = 0, ta = 0;
      uint32_t phase = 0, ta_phase = 0, acc_phase = 0;
      const uint32_t d_corr = tmem_base + NMAIN * TS_BN;
      bool ready = false, skip_corr = false;
      for (long long grp = g0; grp < groups; grp += gstep) {
        mbar_wait_cluster(CEMPTY, acc_phase ^ 1);
        tc_fence_after();
        for (int kb = 0; kb < nkb; ++kb) {
          if (!ready) {
            mbar_wait_cluster(FULL_B(stage), phase);
            PSIF_TRACE2(4);
            mbar_wait_cluster(SPLIT(ta), ta_phase);
            tc_fence_after();
          }
          PSIF_TRACE2(5);
          const uint32_t sa = smem_base + stage * STAGE_BYTES;
          const uint32_t b_hi = sa + X_BYTES, b_lo = b_hi + T2_BH_BYTES;
          const uint32_t a_hi = tmem_base + ACC_COLS + ta * 64, a_lo = a_hi + 32;
          const uint32_t d_main = tmem_base + (uint32_t)((kb % NMAIN) * TS_BN);
          // a K slice of one MMA is 32 bytes of a weight row (8 tf32 or 16 fp16) and 8 TMEM columns of the A operand
          if (!skip_corr) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              mma(d_corr, a_lo + 8 * k, tc_smem_desc(b_hi + k * 32), idesc, (kb | k) != 0);
#pragma unroll
            for (int k = 0; k < 4; ++k) mma(d_corr, a_hi + 8 * k, tc_smem_desc(b_lo + k * 32), idesc, 1);
          }
          skip_corr = false;
          int nstage = stage + 1, nta = ta + 1;
          uint32_t nphase = phase, nta_phase = ta_phase;
          if (nstage == NST) { nstage = 0; nphase ^= 1; }
          if (nta == TA_STAGES) { nta = 0; nta_phase ^= 1; }
          ready = false;
          if (kb == nkb - 1) tc_commit_2sm(CFULL);
          if (kb == 0) {
            // Tile start: the main accumulator is still being drained by the epilogue (TFULL -> tcgen05.ld -> TEMPTY is
            // ~2000 cycles end to end on the timeline).  The correction accumulator was drained earlier, so the
            // correction MMAs of the SECOND K block go ahead of the first K block's main MMAs too: 16 instead of 8 MMAs
            // (1140 cycles) of tensor work cover the drain.
            if (nkb >= 2 && !(dbg & 8)) {
              mbar_wait_cluster(FULL_B(nstage), nphase);
              mbar_wait_cluster(SPLIT(nta), nta_phase);
              tc_fence_after();
              const uint32_t nb_hi = smem_base + nstage * STAGE_BYTES + X_BYTES, nb_lo = nb_hi + T2_BH_BYTES;
              const uint32_t na_hi = tmem_base + ACC_COLS + nta * 64, na_lo = na_hi + 32;
#pragma unroll
              for (int k = 0; k < 4; ++k) mma(d_corr, na_lo + 8 * k, tc_smem_desc(nb_hi + k * 32), idesc, 1);
#pragma unroll
              for (int k = 0; k < 4; ++k) mma(d_corr, na_hi + 8 * k, tc_smem_desc(nb_lo + k * 32), idesc, 1);
              skip_corr = true;
            }
            mbar_wait_cluster(TEMPTY, acc_phase ^ 1);
            tc_fence_after();
          }
          // look ahead, also across the tile boundary (the operands of the next tile's first K block are ready long
          // before its accumulators are): the next K block's barriers are tested in front of this K block's main MMAs
          // and the answers read behind them (tc_mma4_test2_2sm)
          if ((dbg & 4) == 0 && (kb + 1 < nkb || (!(dbg & 1) && grp + gstep < groups))) {
            uint32_t r1, r2;
            tc_mma4_test2_2sm<F16>(d_main, a_hi, tc_smem_desc(b_hi), tc_smem_desc(b_hi + 32), tc_smem_desc(b_hi + 64),
                                   tc_smem_desc(b_hi + 96), idesc, kb >= NMAIN ? 1u : 0u, FULL_B(nstage), nphase, SPLIT(nta),
                                   nta_phase, r1, r2);
            ready = (r1 & r2 & 1u) != 0;
            if (ready) tc_fence_after();
          } else {
            if (kb + 1 < nkb || (!(dbg & 1) && grp + gstep < groups)) {
              mbar_wait_cluster(FULL_B(nstage), nphase);
              mbar_wait_cluster(SPLIT(nta), nta_phase);
              tc_fence_after();
              ready = true;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k)
              mma(d_main, a_hi + 8 * k, tc_smem_desc(b_hi + k * 32), idesc, (kb >= NMAIN) || (k != 0));
          }
          tc_commit_2sm(EMPTY_S(stage));
          if (!ELIDE_A) tc_commit_2sm(EMPTY_A(ta));
          if (kb == nkb - 1) tc_commit_2sm(TFULL);
          PSIF_TRACE2(6);
          ++tcount;
          stage = nstage; phase = nphase; ta = nta; ta_phase = nta_phase;
        }
        acc_phase ^= 1;
      }
    }
  }
  } else if ((warp >= 4 && warp < 8) || warp >= 16) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 80;");
    // splitter: 8 warps.  Thread = (tile row, half of the K block's columns): warps 4-7 take the first half, warps
    // 16-19 the second (a warp reaches the TMEM lanes 32 (warp % 4) ..).  One warp doing whole rows needed ~1700 cycles
    // per 64-column K block (LDS latency + ~300 ALU instructions + tcgen05.st + the arrive, all serial) against 852
    // cycles of MMA work: with fp16 operands the splitter, not the tensor core, set the pace.
    int stage = 0, ta = 0;
    uint32_t phase = 0, ta_phase = 0;
    const int q = warp & 3, sh = warp >= 16 ? 1 : 0;
    const int row = q * 32 + lane;
    const uint32_t split_leader0 = mapa_rank(SPLIT(0), 0);
    // lane 0 tests the NEXT K block's FULL_X right behind the loads of the current one and looks at the answer a
    // whole K block later; it only falls back to a blocking wait when the ring has run dry
    uint32_t early = 0;
    float amax = 0.f;                      // fp16 mode: largest |x| this thread has split
    for (long long grp = g0; grp < groups; grp += gstep) {
      for (int kb = 0; kb < nkb; ++kb) {
        if (lane == 0) {
          if (!(early & 1)) mbar_wait(FULL_X(stage), phase);
          if (!ELIDE_A) mbar_wait(EMPTY_A(ta), ta_phase ^ 1);
        }
        __syncwarp();
        if (warp == 4 && lane == 0) PSIF_TRACE2(7);
        tc_fence_after();
        uint32_t hi[16], lo[16];
        if constexpr (PK) {
          // box 0 = h0, box 1 = h1 of the K block's 64 columns, 128-byte rows of fp16; this thread moves columns
          // [32 sh, 32 sh + 32) of both: 16-byte chunks 4 sh .. 4 sh + 3 (un-swizzled), 16 packed TMEM columns each
          const uint8_t* r0 = base + stage * STAGE_BYTES + row * 128;
          uint4 c0[4], c1[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) c0[c] = *reinterpret_cast<const uint4*>(r0 + (((4 * sh + c) ^ (row & 7)) << 4));
#pragma unroll
          for (int c = 0; c < 4; ++c) c1[c] = *reinterpret_cast<const uint4*>(r0 + TC_A_BYTES + (((4 * sh + c) ^ (row & 7)) << 4));
          if (ELIDE_A && lane == 0) {
            const int ns = stage + 1 == NST ? 0 : stage + 1;
            early = mbar_test(FULL_X(ns), ns == 0 ? phase ^ 1 : phase);
          }
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            hi[4 * c] = c0[c].x; hi[4 * c + 1] = c0[c].y; hi[4 * c + 2] = c0[c].z; hi[4 * c + 3] = c0[c].w;
            lo[4 * c] = c1[c].x; lo[4 * c + 1] = c1[c].y; lo[4 * c + 2] = c1[c].z; lo[4 * c + 3] = c1[c].w;
          }
        } else if constexpr (F16) {
          // columns [32 sh, 32 sh + 32) of the K block = TMA box sh; 16 packed TMEM columns each of h0 and h1
          const uint8_t* rp = base + stage * STAGE_BYTES + sh * TC_A_BYTES + row * 128;
          float4 xv[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) xv[c] = *reinterpret_cast<const float4*>(rp + ((c ^ (row & 7)) << 4));
          if (ELIDE_A && lane == 0) {          // behind the loads in the shared-memory pipe, consumed one K block later
            const int ns = stage + 1 == NST ? 0 : stage + 1;
            early = mbar_test(FULL_X(ns), ns == 0 ? phase ^ 1 : phase);
          }
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            // packed pair: even k in the low half-word, as the tensor core reads 16-bit A operands from TMEM
            const __half2 h01 = __floats2half2_rn(xv[c].x, xv[c].y), h23 = __floats2half2_rn(xv[c].z, xv[c].w);
            const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
            const __half2 l01 = __floats2half2_rn((xv[c].x - f01.x) * H_LO_SCALE, (xv[c].y - f01.y) * H_LO_SCALE);
            const __half2 l23 = __floats2half2_rn((xv[c].z - f23.x) * H_LO_SCALE, (xv[c].w - f23.y) * H_LO_SCALE);
            hi[2 * c] = *reinterpret_cast<const uint32_t*>(&h01);
            hi[2 * c + 1] = *reinterpret_cast<const uint32_t*>(&h23);
            lo[2 * c] = *reinterpret_cast<const uint32_t*>(&l01);
            lo[2 * c + 1] = *reinterpret_cast<const uint32_t*>(&l23);
            amax = fmaxf(fmaxf(amax, fmaxf(fabsf(xv[c].x), fabsf(xv[c].y))), fmaxf(fabsf(xv[c].z), fabsf(xv[c].w)));
          }
        } else {
          // columns [16 sh, 16 sh + 16) of the 32-column K block
          const uint8_t* rp = base + stage * STAGE_BYTES + row * 128;
          float4 xv[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) xv[c] = *reinterpret_cast<const float4*>(rp + (((4 * sh + c) ^ (row & 7)) << 4));
          if (ELIDE_A && lane == 0) {
            const int ns = stage + 1 == NST ? 0 : stage + 1;
            early = mbar_test(FULL_X(ns), ns == 0 ? phase ^ 1 : phase);
          }
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float vv[4] = {xv[c].x, xv[c].y, xv[c].z, xv[c].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              // tf32 round-to-nearest (ties away from zero) on the integer pipe: add half a tf32 ulp to the magnitude
              // bits and drop the low 13 (cvt.rna.tf32.f32 runs on a quarter-rate pipe; the splitter sits on the
              // critical path of every K block)
              const uint32_t u = (__float_as_uint(vv[e]) + 0x1000u) & 0xFFFFE000u;
              hi[4 * c + e] = u;
              lo[4 * c + e] = __float_as_uint(vv[e] - __uint_as_float(u));
            }
          }
        }
        const uint32_t ta_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ACC_COLS + ta * 64 + 16 * sh);
        if (warp == 4 && lane == 0) PSIF_TRACE2(11);
        tc_st16(ta_addr, hi);
        tc_st16(ta_addr + 32, lo);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        if (warp == 4 && lane == 0) PSIF_TRACE2(12);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(split_leader0 + 8u * ta);
        if (warp == 4 && lane == 0) PSIF_TRACE2(3);
        ++tcount;
        if (++stage == NST) { stage = 0; phase ^= 1; }
        if (++ta == TA_STAGES) { ta = 0; ta_phase ^= 1; }
      }
    }
    if (F16 && !PK && ovf != nullptr && !(amax < 65504.f)) atomicOr(ovf, 1u);     // also catches NaN / inf inputs
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 128;");
    uint32_t acc_phase = 0;
    float eamax = 0.f;                     // PK: largest |value| this thread has packed for the next GEMM
    const int q = warp & 3, half = (warp - 8) >> 2;
    const uint32_t tempty_leader = mapa_rank(TEMPTY, 0), cempty_leader = mapa_rank(CEMPTY, 0);
    for (long long grp = g0; grp < groups; grp += gstep) {
      const long long m0 = ((grp / tiles_n) * 2 + crank) * rpt;
      const int n0 = (int)(grp % tiles_n) * TS_BN + half * 64;
      const long long r = m0 + q * 32 + lane;
      mbar_wait_warp((dbg & 2) ? TFULL : CFULL, acc_phase);
      if (warp == 8 && lane == 0) PSIF_TRACE2(8);
      tc_fence_after();
      uint32_t v[2][32];
      const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * 64);
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) tc_ld32_nowait(ta + NMAIN * TS_BN + ch * 32, v[ch]);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(cempty_leader);
      mbar_wait_warp(TFULL, acc_phase);
      tc_fence_after();
#pragma unroll
      for (int mj = 0; mj < NMAIN; ++mj) {
        uint32_t w[2][32];
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) tc_ld32_nowait(ta + mj * TS_BN + ch * 32, w[ch]);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (mj == NMAIN - 1) {          // the accumulators are in registers: hand TMEM back before doing arithmetic
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(tempty_leader);
        }
#pragma unroll
        for (int ch = 0; ch < 2; ++ch)
#pragma unroll
          for (int e = 0; e < 32; ++e)
            v[ch][e] = F16 ? __float_as_uint(fmaf(__uint_as_float(v[ch][e]), 1.f / H_LO_SCALE, __uint_as_float(w[ch][e])))
                           : __float_as_uint(__uint_as_float(v[ch][e]) + __uint_as_float(w[ch][e]));
      }
      if (warp == 8 && lane == 0) PSIF_TRACE2(9);
      acc_phase ^= 1;
      if (act == 2) {
        // GELU on the payload: a token's value row gives g, g', g''; its tangent rows are scaled by
        // g' and its Laplacian row becomes g' lap + g'' sum_t t^2.  Rows of a token sit in different threads, so
        // each 64-column half of the tile goes through a shared-memory staging tile [128 rows][68 floats] (the padding
        // makes the row-per-thread writes and the row-segment reads below conflict free without address swizzling).
        // This epilogue must fit under the tile's MMA time (4 K blocks = 3400 cycles for the FC GEMM) in ISSUE slots:
        // 8 warps share 4 schedulers with the other roles.  Per token (C rows x 64 columns) a warp does
        //   pre-pass   lane = 2 columns: g, g', g'' of the value row -> a 3 x 64 scratch row          (2 GELU evaluations)
        //   rows       lane = (row class rc = lane >> 3, 8 columns): tangent rows 1 + rc, 5 + rc, ...: two 16-byte loads,
        //              8 FMA for sum t^2, 8 FMUL, two 16-byte stores (fp32: 32 bytes of the row; PK: 16 bytes of the h0
        //              plane + 16 bytes of the h1 plane of the packed pair the next GEMM consumes)
        //   ends       class 0 writes the value row g, class 1 the Laplacian row g' lap + g'' sum t^2 (one predicated pass)
        constexpr int GS = GELU_STAGE_STRIDE;
        float* gbase = reinterpret_cast<float*>(base + RING_BYTES + 512);
        float* buf = gbase + half * (TC_BM * GS);
        float* gsc = gbase + 2 * (TC_BM * GS) + (warp - 8) * 192;        // this warp's scratch: g | g' | g'' x 64 columns
        const int lr = q * 32 + lane;
        const int tpt = rpt / C;
        const int bar_id = 1 + half;
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");      // previous tile's readers are done
#pragma unroll
        for (int ch = 0; ch < 2; ++ch)
#pragma unroll
          for (int g = 0; g < 8; ++g)
            *reinterpret_cast<float4*>(buf + lr * GS + ch * 32 + 4 * g) =
                make_float4(__uint_as_float(v[ch][4 * g]), __uint_as_float(v[ch][4 * g + 1]), __uint_as_float(v[ch][4 * g + 2]),
                            __uint_as_float(v[ch][4 * g + 3]));
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        if (warp == 8 && lane == 0) PSIF_TRACE2(13);
        const int rc = lane >> 3, c8 = (lane & 7) * 8;
        const float2 b2 = bias ? __ldg(reinterpret_cast<const float2*>(bias + n0 + 2 * lane)) : make_float2(0.f, 0.f);
        constexpr int ESZ = PK ? 2 : 4;                     // bytes per element of the plane a lane's pointer walks
        const long long opitch = (long long)N * 4;          // bytes between payload rows (packed rows are as long as fp32 rows)
        auto st8 = [&](uint8_t* dst, const float (&o)[8], bool pred) {
          if constexpr (PK) {
            uint2 a0, a1, b0, b1;
            pack_split4(make_float4(o[0], o[1], o[2], o[3]), a0, a1, eamax);
            pack_split4(make_float4(o[4], o[5], o[6], o[7]), b0, b1, eamax);
            st_global_u4_if(dst, make_uint4(a0.x, a0.y, b0.x, b0.y), pred);                 // h0 plane
            st_global_u4_if(dst + (long long)N * 2, make_uint4(a1.x, a1.y, b1.x, b1.y), pred);   // h1 plane: N halves further
          } else {
            st_global_v4_if(reinterpret_cast<float*>(dst), make_float4(o[0], o[1], o[2], o[3]), pred);
            st_global_v4_if(reinterpret_cast<float*>(dst) + 4, make_float4(o[4], o[5], o[6], o[7]), pred);
          }
        };
        auto ld8 = [&](const float* src, float (&o)[8]) {
          const float4 a = *reinterpret_cast<const float4*>(src), b = *reinterpret_cast<const float4*>(src + 4);
          o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
        };
        const int jmax = (C - 2 + 3) >> 2;                  // tangent rows per class, rounded up (uniform trip count)
        // warp = every 4th token of the tile
        for (int t = q; t < tpt; t += 4) {
          const long long gr = m0 + (long long)t * C;
          if (gr >= M) break;
          const float* trow = buf + t * C * GS;
          {
            const float2 v0 = *reinterpret_cast<const float2*>(trow + 2 * lane);
            float ga, gb, g1a, g1b, g2a, g2b;
            gelu_tanh_d2(v0.x + b2.x, ga, g1a, g2a);
            gelu_tanh_d2(v0.y + b2.y, gb, g1b, g2b);
            *reinterpret_cast<float2*>(gsc + 2 * lane) = make_float2(ga, gb);
            *reinterpret_cast<float2*>(gsc + 64 + 2 * lane) = make_float2(g1a, g1b);
            *reinterpret_cast<float2*>(gsc + 128 + 2 * lane) = make_float2(g2a, g2b);
          }
          __syncwarp();
          if (warp == 8 && lane == 0 && t == q) PSIF_TRACE2(14);
          float g1v[8], ss[8];
          ld8(gsc + 64 + c8, g1v);
#pragma unroll
          for (int e = 0; e < 8; ++e) ss[e] = 0.f;
          // this lane's position in the token's first output row
          uint8_t* yp = reinterpret_cast<uint8_t*>(Y) + gr * opitch + (long long)(n0 + c8) * ESZ;
          {
            uint8_t* yr = yp + (long long)(1 + rc) * opitch;
            const float* tr_ = trow + (1 + rc) * GS + c8;
            for (int j = 0; j < jmax; ++j) {
              const bool live = 1 + rc + 4 * j < C - 1;
              float tv[8], o[8];
              ld8(live ? tr_ : trow + c8, tv);
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const float x = live ? tv[e] : 0.f;
                ss[e] = fmaf(x, x, ss[e]);
                o[e] = g1v[e] * x;
              }
              st8(yr, o, live);
              yr += 4 * opitch;
              tr_ += 4 * GS;
            }
          }
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            ss[e] += __shfl_xor_sync(0xffffffffu, ss[e], 8);
            ss[e] += __shfl_xor_sync(0xffffffffu, ss[e], 16);
          }
          if (warp == 8 && lane == 0 && t == q) PSIF_TRACE2(15);
          {
            float xg[8], vl[8], o[8];
            ld8(gsc + (rc == 0 ? 0 : 128) + c8, xg);           // g (class 0) or g'' (class 1)
            ld8(trow + (C - 1) * GS + c8, vl);
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = rc == 0 ? xg[e] : fmaf(g1v[e], vl[e], xg[e] * ss[e]);
            st8(yp + (rc == 0 ? 0ll : (long long)(C - 1) * opitch), o, rc == 0 || (rc == 1 && C > 1));   // C == 1: value rows only
          }
          __syncwarp();                                        // the scratch row is rewritten by the next token
          if (warp == 8 && lane == 0 && t == q) PSIF_TRACE2(16);
        }
      } else if (tma_out) {
        // Plain epilogue through TMA.  A thread holds one row x 64 columns; stored straight from registers, every store
        // instruction of a warp touches 32 rows x 32 bytes = 32 cache lines, ~4000 LSU wavefronts per tile (timeline:
        // 4600 cycles of stores per tile against 3400 cycles of MMA work), and a residual doubles that.  Instead each
        // warp writes its 32 x 32 slab to a private 4 KiB staging tile in TMA's SWIZZLE_128B layout and lane 0 hands
        // it to the TMA unit: a plain tensor store, or (residual added in place, res == Y) a tensor REDUCE-ADD, which
        // performs Y += tile in L2 so the residual never travels to the SM at all.
        uint8_t* wbuf = base + RING_BYTES + 1024 + (warp - 8) * 4096;
        const uint32_t wbuf_s = smem_u32(wbuf);
        const bool with_bias = bias != nullptr && (C == 1 || (r % C) == 0);
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          const int c0 = n0 + ch * 32;
          if (c0 >= N) continue;            // ragged last column tile (N is a multiple of 32)
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // staging tile free again
          __syncwarp();
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            float4 o = make_float4(__uint_as_float(v[ch][4 * g]), __uint_as_float(v[ch][4 * g + 1]), __uint_as_float(v[ch][4 * g + 2]),
                                   __uint_as_float(v[ch][4 * g + 3]));
            if (with_bias) {
              const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + c0 + 4 * g));
              o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
            }
            if (act) { o.x = gelu_tanh(o.x); o.y = gelu_tanh(o.y); o.z = gelu_tanh(o.z); o.w = gelu_tanh(o.w); }
            *reinterpret_cast<float4*>(wbuf + lane * 128 + ((g ^ (lane & 7)) << 4)) = o;
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) {
            const int r0 = (int)(m0 + q * 32);
            if (res != nullptr)
              asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
                           ::"l"(&tmY), "r"(wbuf_s), "r"(c0), "r"(r0) : "memory");
            else
              asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];"
                           ::"l"(&tmY), "r"(wbuf_s), "r"(c0), "r"(r0) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
      } else if (r < M && q * 32 + lane < rpt) {
        const bool with_bias = bias != nullptr && (C == 1 || (r % C) == 0);
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          const int c0 = n0 + ch * 32;
          if (c0 >= N) continue;            // ragged last column tile (N is a multiple of 32)
          float* yp = Y + r * (long long)N + c0;
          const float* rp = res ? res + r * (long long)N + c0 : nullptr;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = __uint_as_float(v[ch][8 * g + e]);
            if (with_bias) {
              const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + c0 + 8 * g));
              const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + c0 + 8 * g + 4));
              o[0] += b0.x; o[1] += b0.y; o[2] += b0.z; o[3] += b0.w; o[4] += b1.x; o[5] += b1.y; o[6] += b1.z; o[7] += b1.w;
            }
            if (act) {
#pragma unroll
              for (int e = 0; e < 8; ++e) o[e] = gelu_tanh(o[e]);
            }
            if (rp) {
              float rr[8];
              ld_global_v8(rp + 8 * g, rr);
#pragma unroll
              for (int e = 0; e < 8; ++e) o[e] += rr[e];
            }
            st_global_v8(yp + 8 * g, o);
          }
        }
      }
      if (warp == 8 && lane == 0) PSIF_TRACE2(10);
      ++tcount;
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");      // this warp's TMA stores have landed
    if constexpr (PK) raise_range_flag(ovf, eamax);
  }
#undef PSIF_TRACE2
  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

// split weights once: hi = tf32(w), lo = w - hi
__global__ void tc_split_weights_kernel(const float* __restrict__ w, float* __restrict__ hi, float* __restrict__ lo, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = w[i];
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
  const float h = __uint_as_float(u);
  hi[i] = h;
  lo[i] = v - h;
}

// fp16 split of the weights: h0 = fp16(w), h1 = fp16(2^11 (w - h0))
__global__ void tc_split_weights_h_kernel(const float* __restrict__ w, __half* __restrict__ h0, __half* __restrict__ h1, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = w[i];
  const __half a = __float2half_rn(v);
  h0[i] = a;
  h1[i] = __float2half_rn((v - __half2float(a)) * H_LO_SCALE);
}


// ---- host side ---------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// Everything the GEMM launcher remembers between calls lives here, one instance per PsifHandle (= per device and host
// thread): no process-wide mutable state.  The PSIF_TC_* environment knobs are read once, by tc_ctx_init.
struct TcCtx {
  int device = 0, sms = 0;
  bool configured = false;       // cudaFuncSetAttribute done on `device`
  bool ss_configured = false;    // ... for the packed-operand kernel (gemm_ss.cuh)
  bool use_ss = true;            // PSIF_TC_SS=0: packed operands through tc_gemm_2cta_kernel's copy warps (A/B timing)
  bool ss_kp2 = true;            // PSIF_TC_KP2=0: two-pass reductions as two launches of the packed-operand kernel
  PFN_encodeTiled encode = nullptr;
  int dbg = 0;                   // PSIF_TC_EXPERIMENT: A/B switches for the tile-boundary handshakes (results stay correct)
  int kpass = 512;               // PSIF_TC_KPASS: K pass length in columns (0 = never split)
  bool kpass_env = false;
  bool fuse_gelu = true;         // PSIF_TC_FUSE_GELU=0 keeps the payload GELU a separate kernel
  long long* trace = nullptr;    // tools only: device buffer [2][18][512] for a clock64 timeline of cluster 0
  // weight tensor maps per (pointer, N, K pass, row pitch): they only change with psif_set_params' buffers, which are
  // allocated once per handle
  std::map<std::tuple<const void*, int, int, int>, CUtensorMap> wmaps;
};

inline int32_t tc_ctx_init(TcCtx& cx) {
  PSIF_CUDA_CHECK(cudaGetDevice(&cx.device));
  PSIF_CUDA_CHECK(cudaDeviceGetAttribute(&cx.sms, cudaDevAttrMultiProcessorCount, cx.device));
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
    cx.encode = reinterpret_cast<PFN_encodeTiled>(p);
  if (const char* e = getenv("PSIF_TC_EXPERIMENT")) cx.dbg = atoi(e);
  if (const char* e = getenv("PSIF_TC_KPASS")) {
    cx.kpass_env = true;
    cx.kpass = atoi(e);
    if (cx.kpass % TC_BK) cx.kpass = 512;
  }
  if (const char* e = getenv("PSIF_TC_FUSE_GELU")) cx.fuse_gelu = e[0] != '0';
  if (const char* e = getenv("PSIF_TC_SS")) cx.use_ss = e[0] != '0';
  if (const char* e = getenv("PSIF_TC_KP2")) cx.ss_kp2 = e[0] != '0';
  return PSIF_OK;
}

// 2-D fp32 row-major [rows][K] tensor, box = 32 columns x box_rows rows, 128-byte swizzle
inline int32_t tc_make_map(const TcCtx& cx, CUtensorMap* map, const float* ptr, long long rows, int K, int box_rows, int ld = 0) {
  if (!cx.encode) return fail(PSIF_E_CUDA, "cuTensorMapEncodeTiled entry point not available%s");
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)(ld ? ld : K) * 4};     // ld: row pitch in elements when [rows x K] is a column slice
  cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = cx.encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(PSIF_E_CUDA, "cuTensorMapEncodeTiled failed (%s%lld)", "", (long long)r);
  return PSIF_OK;
}

// 2-D fp16 row-major [rows][K] weight tensor (row pitch ld elements), box = 64 columns (128 bytes) x box_rows rows
inline int32_t tc_make_map_h(const TcCtx& cx, CUtensorMap* map, const __half* ptr, long long rows, int K, int box_rows, int ld) {
  if (!cx.encode) return fail(PSIF_E_CUDA, "cuTensorMapEncodeTiled entry point not available%s");
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)H_BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = cx.encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(ptr), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(PSIF_E_CUDA, "cuTensorMapEncodeTiled (fp16) failed (%s%lld)", "", (long long)r);
  return PSIF_OK;
}

// shapes the kernel takes: at least four row tiles, N a multiple of 32 (a ragged last column tile is fine: the orbital
// head has N = K_det (n_up + n_dn)), K a multiple of the 32-column K block
inline bool tc_gemm_supported(long long M, int N, int K) {
  return M >= 4 * TC_BM && N % 32 == 0 && N >= 64 && K % TC_BK == 0 && K >= TC_BK;
}

// act == 2 (payload GELU fused into the epilogue) needs whole tokens per 128-row tile; with fewer than two tokens per
// tile (C > 64) more than a third of each tile would be wasted
inline bool tc_gelu_fusable(const TcCtx& cx, long long M, int N, int K, int C) {
  if (!cx.fuse_gelu) return false;
  return tc_gemm_supported(M, N, K) && N % TS_BN == 0 && (C == 1 || (C >= 5 && (TC_BM / C) * C >= 85));
}

// the packed-operand kernel (gemm_ss.cuh)
inline bool ss_gemm_takes(bool a_packed, bool f16, int act, const float* res, const float* Y, int C);
inline int32_t ss_gemm_launch(TcCtx& cx, const CUtensorMap& mx, const CUtensorMap& mh, const CUtensorMap& ml, const float* bias,
                              bool reduce_add, float* Y, long long M, int N, int kk, int C, int act, int rpt, unsigned grid,
                              unsigned* ovf, int a_h1_col, cudaStream_t st, bool kp2);

// Whi / Wlo: tf32 split of the weights; Wh0 / Wh1: their fp16 split (nullptr: tf32 split only); f16_mode selects the
// latter; ovf: device flag raised when an activation does not fit fp16 (see the kernel comment)
// a_packed: X holds the packed fp16 pair written by the producer (common.cuh) instead of fp32 values; fp16 mode only.
// With a_packed the payload-GELU epilogue (act == 2) writes Y packed as well.
inline int32_t tc_gemm(TcCtx& cx, const float* X, const float* Whi, const float* Wlo, const float* bias, const float* res,
                       float* Y, long long M, int N, int K, int C, int act, cudaStream_t st, const __half* Wh0,
                       const __half* Wh1, unsigned* ovf, bool f16_mode, bool a_packed = false) {
  if ((reinterpret_cast<uintptr_t>(X) & 15) || (reinterpret_cast<uintptr_t>(Whi) & 15) || (reinterpret_cast<uintptr_t>(Wlo) & 15) ||
      (reinterpret_cast<uintptr_t>(Y) & 31) || (res && (reinterpret_cast<uintptr_t>(res) & 31)) ||
      (bias && (reinterpret_cast<uintptr_t>(bias) & 15)))
    return fail(PSIF_E_INVALID, "tc_gemm: operands must be 16-byte (outputs 32-byte) aligned%s");
  if (!tc_gemm_supported(M, N, K)) return fail(PSIF_E_INVALID, "tc_gemm: shape not supported%s");
  if (act == 2 && C == 1) act = 1;       // plain rows: the ordinary GELU epilogue
  if (act == 2 && !tc_gelu_fusable(cx, M, N, K, C)) return fail(PSIF_E_INVALID, "tc_gemm: payload GELU not fusable here%s");
  int rpt = TC_BM;                    // rows per tile: whole tokens when the payload GELU runs in the epilogue
  if (act == 2) rpt = (TC_BM / C) * C;
  if (!cx.configured) {
    PSIF_CUDA_CHECK(cudaFuncSetAttribute(tc_gemm_2cta_kernel<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, T2_SMEM_BYTES_GELU));
    PSIF_CUDA_CHECK(cudaFuncSetAttribute(tc_gemm_2cta_kernel<2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, T2_SMEM_BYTES_GELU));
    PSIF_CUDA_CHECK(cudaFuncSetAttribute(tc_gemm_2cta_kernel<1, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, h_smem_bytes(4, false)));
    PSIF_CUDA_CHECK(cudaFuncSetAttribute(tc_gemm_2cta_kernel<1, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, h_smem_bytes(3, true)));
    PSIF_CUDA_CHECK(cudaFuncSetAttribute((tc_gemm_2cta_kernel<1, 4, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, h_smem_bytes(4, false)));
    PSIF_CUDA_CHECK(cudaFuncSetAttribute((tc_gemm_2cta_kernel<1, 3, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, h_smem_bytes(3, true)));
    cx.configured = true;
  }
  const long long groups = (((M + rpt - 1) / rpt + 1) / 2) * ((N + TS_BN - 1) / TS_BN);
  const int smem2 = T2_SMEM_BYTES_GELU;      // ring + epilogue staging (plain: 32 KiB of it, payload GELU: 64 KiB)
  long long nclusters = cx.sms / 2;
  if (groups < nclusters) nclusters = groups;
  const unsigned grid = (unsigned)(nclusters * 2);
  // Long reductions run as passes of at most 512 columns of K, each accumulating onto the previous pass' output in
  // the epilogue (fp32 adds).  One main accumulator then sees at most 64 truncating tensor-core accumulations, the
  // same as the two alternating accumulators (NMAIN = 2) did for K = 1024, but TMEM keeps four operand slots instead
  // of two, which that variant's splitter <-> MMA hand-off could not hide (163 vs 205 TFLOP/s, tools/gemm_bench.py).
  // With fp16 operands an MMA covers 16 columns of K, so a 1024-column pass is the same 64 accumulations: plain rows
  // (C == 1: the Metropolis forward, where only log|psi| at 1e-5 relative is at stake) take K <= 1024 in ONE pass
  // (FC2 420 -> 365 us: one tile boundary and one read-modify-write of Y less).  Payload rows keep 512-column passes:
  // a single pass is as accurate as the tf32 split was (1.2e-6 vs 6e-7 on the GEMM), but it doubles the 90th
  // percentile of |E_L - E_L(fp64)| on random Be walkers (2e-5 -> 4e-5 Ha), and parity comes first.
  int kp = (cx.kpass > 0 && act == 0 && K > cx.kpass) ? cx.kpass : K;
  const bool have_h = f16_mode && Wh0 && Wh1;
  if (!cx.kpass_env && have_h && C == 1 && K <= 1024) kp = K;
  // epilogue through TMA (plain store, or reduce-add when the residual is added in place): 32 x 32 fp32 boxes
  const int tma_out = (act != 2 && (res == nullptr || res == Y) && !(reinterpret_cast<uintptr_t>(Y) & 127)) ? 1 : 0;
  CUtensorMap my;
  PSIF_TRY(tc_make_map(cx, &my, Y, M, N, 32));
  // fp16-split operands: every pass a multiple of 64 columns and at most 1024 (one main accumulator), 16-byte aligned rows
  const bool f16 = have_h && K % H_BK == 0 && kp % H_BK == 0 && kp <= 1024 &&
                   !(reinterpret_cast<uintptr_t>(Wh0) & 15) && !(reinterpret_cast<uintptr_t>(Wh1) & 15);
  if (a_packed && !f16) return fail(PSIF_E_INVALID, "tc_gemm: a packed A operand needs the fp16-split mode%s");
  // Packed-operand kernel, K = two passes: both halves go into the two accumulator pairs of ONE launch and are summed in
  // its epilogue (gemm_ss.cuh, KP2), so the second pass' read-modify-write of the whole output disappears.
  const bool kp2 = cx.use_ss && cx.ss_kp2 && f16 && K == 2 * kp && ss_gemm_takes(a_packed, f16, act, res, Y, C);
  if (kp2) kp = K;
  for (int k0 = 0; k0 < K; k0 += kp) {
    const int kk = K - k0 < kp ? K - k0 : kp;
    CUtensorMap mx, mh, ml;
    if (a_packed)      // fp16 view [M][2 K] of the packed rows, shifted to this pass: h0 at column 0, h1 at column K
      PSIF_TRY(tc_make_map_h(cx, &mx, reinterpret_cast<const __half*>(X) + k0, M, K + kk, TC_BM, 2 * K));
    else
      PSIF_TRY(tc_make_map(cx, &mx, X + k0, M, kk, TC_BM, K));
    for (int which = 0; which < 2; ++which) {
      const void* wp = f16 ? (const void*)((which ? Wh1 : Wh0) + k0) : (const void*)((which ? Wlo : Whi) + k0);
      auto key = std::make_tuple(wp, N, kk, K);
      auto it = cx.wmaps.find(key);
      if (it == cx.wmaps.end()) {
        CUtensorMap m;
        if (f16) PSIF_TRY(tc_make_map_h(cx, &m, static_cast<const __half*>(wp), N, kk, TS_BN / 2, K));
        else PSIF_TRY(tc_make_map(cx, &m, static_cast<const float*>(wp), N, kk, TS_BN / 2, K));
        it = cx.wmaps.emplace(key, m).first;
      }
      (which ? ml : mh) = it->second;
    }
    const float* bias_p = k0 == 0 ? bias : nullptr;
    const float* res_p = k0 == 0 ? res : Y;
    if (cx.use_ss && ss_gemm_takes(a_packed, f16, act, res_p, Y, C)) {
      PSIF_TRY(ss_gemm_launch(cx, mx, mh, ml, bias_p, res_p != nullptr, Y, M, N, kk, C, act, rpt, grid, ovf, K, st, kp2));
      continue;
    }
    if (f16 && a_packed) {
      if (act == 1) return fail(PSIF_E_INVALID, "tc_gemm: packed operand + plain GELU needs the packed-operand kernel (PSIF_TC_SS)%s");
      if (act == 2)
        PSIF_LAUNCH((tc_gemm_2cta_kernel<1, 3, true>), grid, T2_THREADS, h_smem_bytes(3, true), st, mx, mh, ml, bias_p, res_p, Y, M, N, kk, C, act, cx.trace, rpt, cx.dbg, ovf, my, tma_out, K);
      else
        PSIF_LAUNCH((tc_gemm_2cta_kernel<1, 4, true>), grid, T2_THREADS, h_smem_bytes(4, false), st, mx, mh, ml, bias_p, res_p, Y, M, N, kk, C, act, cx.trace, rpt, cx.dbg, ovf, my, tma_out, K);
    } else if (f16) {
      if (act == 2)
        PSIF_LAUNCH((tc_gemm_2cta_kernel<1, 3>), grid, T2_THREADS, h_smem_bytes(3, true), st, mx, mh, ml, bias_p, res_p, Y, M, N, kk, C, act, cx.trace, rpt, cx.dbg, ovf, my, tma_out, 0);
      else
        PSIF_LAUNCH((tc_gemm_2cta_kernel<1, 4>), grid, T2_THREADS, h_smem_bytes(4, false), st, mx, mh, ml, bias_p, res_p, Y, M, N, kk, C, act, cx.trace, rpt, cx.dbg, ovf, my, tma_out, 0);
    } else if (kk > 512) {
      PSIF_LAUNCH((tc_gemm_2cta_kernel<2, 0>), grid, T2_THREADS, smem2, st, mx, mh, ml, bias_p, res_p, Y, M, N, kk, C, act, cx.trace, rpt, cx.dbg, (unsigned*)nullptr, my, tma_out, 0);
    } else {
      PSIF_LAUNCH((tc_gemm_2cta_kernel<1, 0>), grid, T2_THREADS, smem2, st, mx, mh, ml, bias_p, res_p, Y, M, N, kk, C, act, cx.trace, rpt, cx.dbg, (unsigned*)nullptr, my, tma_out, 0);
    }
  }
  return PSIF_OK;
}

}  // namespace psif
