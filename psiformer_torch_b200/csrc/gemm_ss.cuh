// Split-precision GEMM, packed-operand form:   Y[M][N] (+)= X[M][K] * W[N][K]^T  (+ bias on rows r % C == 0)
//
// tc_gemm_ss_kernel: the A operand arrives as the packed fp16 pair its PRODUCER wrote (common.cuh: LayerNorm, attention,
// the payload-GELU epilogue), so nothing has to be converted and both operands of every tcgen05.mma are read straight
// from shared memory through descriptors (SS form).  tools/cp_rate.cu, B200: 12 SS MMAs of a K block take 880 cycles
// against 869 with A in tensor memory -- the tensor core's shared-memory read path is not the limit -- while
// tcgen05.cp of the same operands into TMEM costs 620 cycles on top.  Without TMEM operand slots all 512 TMEM columns
// hold accumulators: TWO {main, correction} pairs, so the MMA stream of tile t + 1 starts the moment its operands are
// there while tile t is drained and stored.  (tc_gemm_2cta_kernel, one accumulator pair + four TMEM operand slots fed by
// eight splitter warps, stalled ~2200 of every 5650 cycles at the tile boundary: tensor pipe 56 % busy, ncu round 2.)
//
//   cluster of 2 CTAs = one 256 x 128 output tile per step (cta_group::2, M = 256: 128 rows per CTA), persistent
//   warp 0   TMA producer (each CTA: its 128 rows of the h0 and h1 planes of X, its 64-row halves of W_h0 and W_h1;
//            all four boxes signal the LEADER's FULL barrier)
//   warp 1   MMA issuer (leader CTA, one elected lane): per 64-column K block
//                corr += X_h1 W_h0 (4 MMAs), corr += X_h0 W_h1 (4), main += X_h0 W_h0 (4)
//            the next stage's FULL barrier is tested in FRONT of the last four and the answer read behind them
//   warp 2   TMEM allocation (512 columns: accumulator pair b at columns 256 b: main | corr)
//   warps 4-11, 12-19   two epilogue groups; group g takes the tiles with (tile counter & 1) == g, i.e. accumulator pair g,
//            and so has two tile periods for: TMEM -> registers, Y = main + 2^-11 corr (+ bias), 2 KiB staging tile,
//            TMA tensor store -- or TMA reduce-add when the residual is added in place (it then never visits the SM)
//
// Payload GELU (GELU = true, the MLP up-projection in energy mode): row tiles start rpt = floor(128 / C) C rows apart so
// that a tile holds whole tokens; see the epilogue.
#pragma once
#include "gemm_tcgen05.cuh"

#ifndef PSIF_SS_STATS
#define PSIF_SS_STATS 0      // 1: build with the wait-cycle counters (python -m psiformer_torch_b200.build --stats)
#endif

namespace psif {

constexpr int SS_THREADS = 640;
constexpr int SS_STAGE_BYTES = 2 * TC_A_BYTES + 2 * T2_BH_BYTES;         // 48 KiB: X_h0 | X_h1 | W_h0 half | W_h1 half
constexpr int SS_OUT_BYTES = 2048;                                       // per epilogue warp: 32 rows x 16 fp32, SWIZZLE_64B
constexpr int SS_BAR_BYTES = 512;                                        // mbarriers, TMEM slot
// payload GELU, per epilogue group: a [128 rows][68 floats] staging tile of one 64-column half (the padding makes the
// row-per-lane 16-byte writes conflict free) + the token table {g, g', g'' sum t^2} x 64 columns of up to 9 tokens
constexpr int SS_GELU_STRIDE = 68;
constexpr int SS_GELU_MAX_TOKENS = 9;
constexpr int SS_GELU_GROUP_BYTES = TC_BM * SS_GELU_STRIDE * 4 + SS_GELU_MAX_TOKENS * 3 * 64 * 4;
constexpr int ss_smem_bytes(int nst, bool gelu) {
  return nst * SS_STAGE_BYTES + 1024 + SS_BAR_BYTES + (gelu ? 2 * SS_GELU_GROUP_BYTES : 16 * SS_OUT_BYTES);
}
static_assert(ss_smem_bytes(3, true) <= 232448 && ss_smem_bytes(4, false) <= 232448, "227 KiB of shared memory per CTA");

__device__ __forceinline__ void tc_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}

__device__ __forceinline__ void tc_ld8_nowait(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}

// Four SS MMAs over the four 16-column K slices of a 64-column fp16 tile (descriptor start addresses advance by 32 bytes).
// TEST: a non-blocking mbarrier test is issued in FRONT of them and its predicate read BEHIND them, in one asm block --
// a tcgen05.mma issue blocks until the tensor core's shallow queue has room, so the ~300-cycle round trip of the test
// overlaps the issues instead of draining the pipe (gemm_tcgen05.cuh, tc_mma4_test2_2sm).
template <bool TEST>
__device__ __forceinline__ uint32_t ss_mma4(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc0, uint32_t bar = 0,
                                            uint32_t par = 0) {
  uint32_t r = 0;
  if constexpr (TEST)
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t.reg .b64 a1, a2, a3, b1, b2, b3;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 q, [%6], %7;\n\t"
        "add.u64 a1, %2, 2;\n\tadd.u64 a2, %2, 4;\n\tadd.u64 a3, %2, 6;\n\t"
        "add.u64 b1, %3, 2;\n\tadd.u64 b2, %3, 4;\n\tadd.u64 b3, %3, 6;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%1], %2, %3, %4, p;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%1], a1, b1, %4, 1;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%1], a2, b2, %4, 1;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%1], a3, b3, %4, 1;\n\t"
        "selp.u32 %0, 1, 0, q;\n\t}"
        : "=r"(r)
        : "r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc0), "r"(bar), "r"(par)
        : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 a1, a2, a3, b1, b2, b3;\n\t"
        "add.u64 a1, %1, 2;\n\tadd.u64 a2, %1, 4;\n\tadd.u64 a3, %1, 6;\n\t"
        "add.u64 b1, %2, 2;\n\tadd.u64 b2, %2, 4;\n\tadd.u64 b3, %2, 6;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], a1, b1, %3, 1;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], a2, b2, %3, 1;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], a3, b3, %3, 1;\n\t}"
        ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc0)
        : "memory");
  return r;
}

// payload-GELU epilogue, phase 2: thread = (token, V adjacent columns): g, g', g'' of the value row (+ bias) and the sum over
// the tangent rows of t^2 -> token table {g | g' | g'' sum t^2} x 64 columns.  V is chosen so that the (token, column
// group) items fill the group's 256 threads in as few rounds as possible (Be: 9 tokens x 16 groups of 4 = one round
// instead of three with single columns).
template <int V>
__device__ __forceinline__ void ss_gelu_table(const float* stg, float* table, const float (&bv)[4], int tid8, int tpt, int C) {
  constexpr int GS = SS_GELU_STRIDE, CPT = 64 / V;
  for (int idx = tid8; idx < tpt * CPT; idx += 256) {
    const int t = idx / CPT, col = (idx % CPT) * V;
    const float* tp = stg + t * C * GS + col;
    float u[V], ss[V];
    if constexpr (V == 4) { const float4 x = *reinterpret_cast<const float4*>(tp); u[0] = x.x; u[1] = x.y; u[2] = x.z; u[3] = x.w; }
    else if constexpr (V == 2) { const float2 x = *reinterpret_cast<const float2*>(tp); u[0] = x.x; u[1] = x.y; }
    else u[0] = tp[0];
#pragma unroll
    for (int v = 0; v < V; ++v) ss[v] = 0.f;
#pragma unroll 4
    for (int c = 1; c < C - 1; ++c) {
      float x[V];
      if constexpr (V == 4) { const float4 y = *reinterpret_cast<const float4*>(tp + c * GS); x[0] = y.x; x[1] = y.y; x[2] = y.z; x[3] = y.w; }
      else if constexpr (V == 2) { const float2 y = *reinterpret_cast<const float2*>(tp + c * GS); x[0] = y.x; x[1] = y.y; }
      else x[0] = tp[c * GS];
#pragma unroll
      for (int v = 0; v < V; ++v) ss[v] = fmaf(x[v], x[v], ss[v]);
    }
    float g[V], g1[V], g2s[V];
#pragma unroll
    for (int v = 0; v < V; ++v) {
      float g2;
      gelu_tanh_d2(u[v] + bv[v], g[v], g1[v], g2);
      g2s[v] = g2 * ss[v];
    }
    float* te = table + t * 192 + col;
    if constexpr (V == 4) {
      *reinterpret_cast<float4*>(te) = make_float4(g[0], g[1], g[2], g[3]);
      *reinterpret_cast<float4*>(te + 64) = make_float4(g1[0], g1[1], g1[2], g1[3]);
      *reinterpret_cast<float4*>(te + 128) = make_float4(g2s[0], g2s[1], g2s[2], g2s[3]);
    } else if constexpr (V == 2) {
      *reinterpret_cast<float2*>(te) = make_float2(g[0], g[1]);
      *reinterpret_cast<float2*>(te + 64) = make_float2(g1[0], g1[1]);
      *reinterpret_cast<float2*>(te + 128) = make_float2(g2s[0], g2s[1]);
    } else {
      te[0] = g[0]; te[64] = g1[0]; te[128] = g2s[0];
    }
  }
}

// ... phase 3 operands of one (row, 8 columns) item: the row's values and its (a, m) from the token table
struct SsGeluItem { float4 x0, x1, m0, m1, a0, a1; };
__device__ __forceinline__ void ss_gelu_item_load(const float* stg, const float* table, int row, int c8, int C, uint32_t inv_c,
                                                  SsGeluItem& it) {
  constexpr int GS = SS_GELU_STRIDE;
  const int t = (int)(((uint32_t)row * inv_c) >> 16), c = row - t * C;       // row / C for row < 128, 5 <= C <= 64
  const float* xp = stg + row * GS + c8;
  const float* te = table + t * 192 + c8;
  it.x0 = *reinterpret_cast<const float4*>(xp); it.x1 = *reinterpret_cast<const float4*>(xp + 4);
  it.m0 = make_float4(0.f, 0.f, 0.f, 0.f); it.m1 = it.m0; it.a0 = it.m0; it.a1 = it.m0;
  if (c != 0) { it.m0 = *reinterpret_cast<const float4*>(te + 64); it.m1 = *reinterpret_cast<const float4*>(te + 68); }
  if (c == 0 || c == C - 1) {
    const float* ap = te + (c == 0 ? 0 : 128);
    it.a0 = *reinterpret_cast<const float4*>(ap); it.a1 = *reinterpret_cast<const float4*>(ap + 4);
  }
}
__device__ __forceinline__ void ss_gelu_item_store(const SsGeluItem& it, uint8_t* yp, long long plane_bytes, bool live, float& eamax) {
  const float4 o0 = make_float4(fmaf(it.m0.x, it.x0.x, it.a0.x), fmaf(it.m0.y, it.x0.y, it.a0.y), fmaf(it.m0.z, it.x0.z, it.a0.z),
                                fmaf(it.m0.w, it.x0.w, it.a0.w));
  const float4 o1 = make_float4(fmaf(it.m1.x, it.x1.x, it.a1.x), fmaf(it.m1.y, it.x1.y, it.a1.y), fmaf(it.m1.z, it.x1.z, it.a1.z),
                                fmaf(it.m1.w, it.x1.w, it.a1.w));
  uint2 p0, p1, q0, q1;
  pack_split4(o0, p0, p1, eamax);
  pack_split4(o1, q0, q1, eamax);
  st_global_u4_if(yp, make_uint4(p0.x, p0.y, q0.x, q0.y), live);                   // h0 plane
  st_global_u4_if(yp + plane_bytes, make_uint4(p1.x, p1.y, q1.x, q1.y), live);     // h1 plane: N halves further
}

template <int NST, bool GELU, bool KP2 = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(SS_THREADS, 1)
tc_gemm_ss_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmWhi,
                  const __grid_constant__ CUtensorMap tmWlo, const float* __restrict__ bias, float* Y, long long M, int N, int K,
                  int C, int act, int rpt, unsigned* ovf, const __grid_constant__ CUtensorMap tmY, int reduce_add, int a_h1_col, int dbg, long long* stats, int out_packed) {
  // out_packed (plain epilogue, no residual): Y is written as the packed fp16 pair (the value path's GELU output, which
  // the down-projection consumes); tmY is then an fp16 map over [M][2 N] with 16 x 32 boxes
  // stats (tools only, may be NULL): cluster 0 adds the cycles its roles spend blocked -- [0] MMA loop total, [1] on FULL,
  // [2] on ACC_EMPTY, [3] blocking FULL waits, [4] K blocks, [5] producer on EMPTY, [6] epilogue warp 4 on ACC_FULL,
  // [7] ... on its staging tile (wait_group.read), [8] ... busy from ACC_FULL to its last store, [9] its tiles
  // dbg (PSIF_TC_EXPERIMENT, tools only; results are WRONG with any bit set): 1 no output stores, 2 no MMAs, 4 no X loads,
  // 8 no epilogue work at all -- what each part of the pipeline costs when the others are taken away
#if PSIF_SS_STATS
  const bool st_on = stats != nullptr && blockIdx.x == 0;
#else
  constexpr bool st_on = false;      // the counters cost registers in the single-lane roles: compiled in for tools/ss_stats.py only
  (void)stats;
#endif
  static_assert(!(GELU && KP2), "the payload-GELU GEMM has K = d: one pass");
  constexpr int RING_BYTES = NST * SS_STAGE_BYTES;
  extern __shared__ uint8_t tc_smem_raw[];
  uint8_t* base = tc_smem_raw + ((1024u - (smem_u32(tc_smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + RING_BYTES);
  const uint32_t bar0 = smem_u32(bars);
  auto FULL = [&](int s) { return bar0 + 8u * s; };              // used in the leader: both CTAs' boxes of stage s have landed
  auto EMPTY = [&](int s) { return bar0 + 8u * (8 + s); };       // per CTA (multicast commit): stage s consumed
  auto ACC_FULL = [&](int b) { return bar0 + 8u * (16 + b); };   // per CTA (multicast commit): accumulator pair b complete
  auto ACC_EMPTY = [&](int b) { return bar0 + 8u * (18 + b); };  // used in the leader: pair b drained by 8 warps of each CTA
  // KP2: a pair is filled once per TILE but drained by alternating groups, so each (pair, group) has its own FULL barrier --
  // a group waiting on a shared one would skip a phase and could mistake the other group's tile for its own
  auto ACC_FULL2 = [&](int b, int g) { return bar0 + 8u * (20 + 2 * b + g); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  const bool leader = crank == 0;
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmWhi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmWlo) : "memory");
    if (!GELU) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmY) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < NST; ++s) { mbar_init(FULL(s), 1); mbar_init(EMPTY(s), 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(ACC_FULL(b), 1); mbar_init(ACC_EMPTY(b), 16);
      mbar_init(ACC_FULL2(b, 0), 1); mbar_init(ACC_FULL2(b, 1), 1);
    }
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

  const int tiles_n = (N + TS_BN - 1) / TS_BN;
  const long long tiles_m = (M + rpt - 1) / rpt;
  const long long groups = ((tiles_m + 1) / 2) * tiles_n;        // a group = 2 row tiles x 1 column tile
  const int nkb = K / H_BK;
  const uint32_t smem_base = smem_u32(base);
  const long long g0 = (long long)cluster_id_x(), gstep = (long long)cluster_count_x();

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 0) {
      if (elect_one()) {
        int stage = 0;
        uint32_t phase = 0;
        const uint32_t full_leader0 = mapa_rank(FULL(0), 0);
        for (long long grp = g0; grp < groups; grp += gstep) {
          const int m0 = (int)(((grp / tiles_n) * 2 + crank) * rpt), n0 = (int)(grp % tiles_n) * TS_BN + (int)crank * (TS_BN / 2);
          // (an L2 prefetch of the next tile's X rows from here made every shape SLOWER: QKV 256 -> 281 us, FC2 314 -> 405 us,
          // profiles/gemm_ss_prefetch_r02m.txt -- the ring's loads queue behind the prefetch burst)
          for (int kb = 0; kb < nkb; ++kb) {
            const long long tw0 = st_on ? clock64() : 0;
            mbar_wait(EMPTY(stage), phase ^ 1);
            if (st_on) stats[5] += clock64() - tw0;
            const uint32_t sa = smem_base + stage * SS_STAGE_BYTES;
            if (leader) mbar_arrive_expect_tx(FULL(stage), (dbg & 4) ? 4 * T2_BH_BYTES : 2 * SS_STAGE_BYTES);   // the boxes of both CTAs
            const uint32_t lb = full_leader0 + 8u * stage;
            if (!(dbg & 4)) {
              tma_load_2d_2sm(sa, &tmX, kb * H_BK, m0, lb);
              tma_load_2d_2sm(sa + TC_A_BYTES, &tmX, a_h1_col + kb * H_BK, m0, lb);
            }
            tma_load_2d_2sm(sa + 2 * TC_A_BYTES, &tmWhi, kb * H_BK, n0, lb);
            tma_load_2d_2sm(sa + 2 * TC_A_BYTES + T2_BH_BYTES, &tmWlo, kb * H_BK, n0, lb);
            if (++stage == NST) { stage = 0; phase ^= 1; }
          }
        }
      }
    } else if (warp == 1) {
      constexpr uint32_t idesc = tc_idesc_f16(2 * TC_BM, TS_BN);
      if (leader && elect_one()) {
        // A segment = the K blocks that go into one accumulator pair: a whole tile, or (KP2) one half of the tile's K
        // range.  Segment sg uses pair sg & 1 for the (sg >> 1)-th time.
        constexpr int NP = KP2 ? 2 : 1;
        const int nkb_seg = nkb / NP;
        int stage = 0;
        uint32_t phase = 0, sg = 0;
        bool ready = false, acc_ready = false;
        const long long tl0 = st_on ? clock64() : 0;
        long long w_full = 0, w_acc = 0, n_block = 0, n_kb = 0;
        for (long long grp = g0; grp < groups; grp += gstep) {
          for (int pass = 0; pass < NP; ++pass, ++sg) {
            const uint32_t b = sg & 1u;
            if (!acc_ready) {
              const long long tw0 = st_on ? clock64() : 0;
              mbar_wait_cluster(ACC_EMPTY(b), ((sg >> 1) & 1u) ^ 1u);
              if (st_on) w_acc += clock64() - tw0;
            }
            acc_ready = false;
            tc_fence_after();
            const uint32_t d_main = tmem_base + b * 256u, d_corr = d_main + TS_BN;
            const bool more = grp + gstep < groups || pass + 1 < NP;
            for (int kb = 0; kb < nkb_seg; ++kb) {
              if (!ready) {
                const long long tw0 = st_on ? clock64() : 0;
                mbar_wait_cluster(FULL(stage), phase);
                if (st_on) { w_full += clock64() - tw0; ++n_block; }
              }
              ++n_kb;
              tc_fence_after();
              const uint32_t sa = smem_base + stage * SS_STAGE_BYTES;
              const uint64_t a_h0 = tc_smem_desc(sa), a_h1 = tc_smem_desc(sa + TC_A_BYTES);
              const uint64_t b_hi = tc_smem_desc(sa + 2 * TC_A_BYTES), b_lo = tc_smem_desc(sa + 2 * TC_A_BYTES + T2_BH_BYTES);
              const uint32_t acc = kb != 0 ? 1u : 0u;
              int nstage = stage + 1;
              uint32_t nphase = phase;
              if (nstage == NST) { nstage = 0; nphase ^= 1; }
              const bool last = kb == nkb_seg - 1;
              if (dbg & 2) {
                ready = false; acc_ready = false;
              } else if (last && more) {
                // the next segment's accumulator pair (drained long ago unless the epilogues are the bottleneck)
                const uint32_t nsg = sg + 1;
                acc_ready = ss_mma4<true>(d_corr, a_h1, b_hi, idesc, acc, ACC_EMPTY(nsg & 1u), ((nsg >> 1) & 1u) ^ 1u) != 0;
              } else {
                ss_mma4<false>(d_corr, a_h1, b_hi, idesc, acc);
              }
              if (!(dbg & 2)) ss_mma4<false>(d_corr, a_h0, b_lo, idesc, 1u);
              if (dbg & 2) {
              } else if (!last || more) {
                ready = ss_mma4<true>(d_main, a_h0, b_hi, idesc, acc, FULL(nstage), nphase) != 0;
              } else {
                ss_mma4<false>(d_main, a_h0, b_hi, idesc, acc);
                ready = false;
              }
              tc_commit_2sm(EMPTY(stage));
              if (last) tc_commit_2sm(KP2 ? ACC_FULL2(b, (sg >> 1) & 1u) : ACC_FULL(b));
              stage = nstage; phase = nphase;
            }
          }
        }
        if (st_on) { stats[0] += clock64() - tl0; stats[1] += w_full; stats[2] += w_acc; stats[3] += n_block; stats[4] += n_kb; }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;   // 4 x 128 x 104 + 128 x 56 <= the 640 x 96 registers the CTA was launched with");
    const int e = warp - 4, grpid = e >> 3, w8 = e & 7;
    const int q = warp & 3;                                    // TMEM lane quadrant this warp may touch
    const uint32_t acc_empty_leader = mapa_rank(ACC_EMPTY(KP2 ? 0 : grpid), 0);
    float eamax = 0.f;
    uint32_t it = 0;
    for (long long grp = g0; grp < groups; grp += gstep, ++it) {
      if ((int)(it & 1u) != grpid) continue;
      const long long m0 = ((grp / tiles_n) * 2 + crank) * rpt;
      const int nt0 = (int)(grp % tiles_n) * TS_BN;
      const bool est = st_on && warp == 4 && lane == 0;
      if constexpr (!GELU) {
        // warp = (quadrant q, column half): 32 rows x 64 columns.  The whole slab goes to registers FIRST and the pair is
        // handed back at once (stats: as long as it was held through the four store rounds the MMA issuer spent 35 % of its
        // time waiting for an empty pair).  KP2: the tile's second K half sits in pair 1 and is added here -- fp32, one
        // rounding more than a single accumulator, but each main accumulator sees only K / 32 truncating tensor-core
        // additions (gemm_tcgen05.cuh) -- so the tile leaves in ONE store or reduce-add; as separate passes the second one
        // re-read and re-wrote the whole output (FC2: 470 of 1880 MB).
        // Then four 16-column rounds through a 2 KiB staging tile in TMA's SWIZZLE_64B layout (16-byte chunk c of row r at
        // c ^ ((r >> 1) & 3): the row-per-lane writes are conflict free) and a TMA tensor store / reduce-add each.
        const int half = w8 >> 2;
        const int n0 = nt0 + half * 64;
        const long long r = m0 + q * 32 + lane;
        const bool with_bias = bias != nullptr && (C == 1 || (r % C) == 0);
        // this warp's 64 bias values, two per lane, fetched BEFORE the wait for the accumulators (a __ldg per store round sat
        // on the critical path with a full L2 latency each); a value row picks its columns up by shuffle below
        float2 bv = make_float2(0.f, 0.f);
        if (bias != nullptr && n0 + 2 * lane < N) bv = __ldg(reinterpret_cast<const float2*>(bias + n0 + 2 * lane));
        uint8_t* wbuf = base + RING_BYTES + SS_BAR_BYTES + e * SS_OUT_BYTES;
        const uint32_t wbuf_s = smem_u32(wbuf);
        const long long te0 = est ? clock64() : 0;
        mbar_wait_warp(KP2 ? ACC_FULL2(0, grpid) : ACC_FULL(grpid), (it >> 1) & 1u);
        const long long te1 = est ? clock64() : 0;
        if (est) { stats[6] += te1 - te0; stats[9] += 1; }
        tc_fence_after();
        const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(KP2 ? 0 : grpid * 256) + half * 64;
        float h[64];
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          uint32_t vm[16], vc[16];
          tc_ld16_nowait(ta + ch * 16, vm);
          tc_ld16_nowait(ta + TS_BN + ch * 16, vc);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int i = 0; i < 16; ++i) h[16 * ch + i] = fmaf(__uint_as_float(vc[i]), 1.f / H_LO_SCALE, __uint_as_float(vm[i]));
        }
        tc_fence_before();
        __syncwarp();
        if (est) stats[10] += clock64() - te1;                               // TMEM -> registers
        if (lane == 0) mbar_arrive_cluster(acc_empty_leader);
        if constexpr (KP2) {
          mbar_wait_warp(ACC_FULL2(1, grpid), (it >> 1) & 1u);
          tc_fence_after();
#pragma unroll
          for (int ch = 0; ch < 8; ++ch) {
            uint32_t vm[8], vc[8];
            tc_ld8_nowait(ta + 256 + ch * 8, vm);
            tc_ld8_nowait(ta + 256 + TS_BN + ch * 8, vc);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int i = 0; i < 8; ++i) h[8 * ch + i] += fmaf(__uint_as_float(vc[i]), 1.f / H_LO_SCALE, __uint_as_float(vm[i]));
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(acc_empty_leader + 8u);          // ACC_EMPTY(1) sits right behind ACC_EMPTY(0)
        }
        const bool any_bias = __any_sync(0xffffffffu, with_bias);
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const int c0 = n0 + ch * 16;
          if (c0 >= N || (dbg & 8)) continue;   // ragged last column tile (N is a multiple of 32; W rows >= N are TMA zero fill)
          const long long tg0 = est ? clock64() : 0;
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // staging tile free again
          if (est) stats[7] += clock64() - tg0;
          __syncwarp();
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float4 o = make_float4(h[16 * ch + 4 * g], h[16 * ch + 4 * g + 1], h[16 * ch + 4 * g + 2], h[16 * ch + 4 * g + 3]);
            if (any_bias) {
              const int sl0 = 8 * ch + 2 * g;            // lane holding columns 16 ch + 4 g, + 1; the next lane holds + 2, + 3
              const float b0 = __shfl_sync(0xffffffffu, bv.x, sl0), b1 = __shfl_sync(0xffffffffu, bv.y, sl0);
              const float b2 = __shfl_sync(0xffffffffu, bv.x, sl0 + 1), b3 = __shfl_sync(0xffffffffu, bv.y, sl0 + 1);
              if (with_bias) { o.x += b0; o.y += b1; o.z += b2; o.w += b3; }
            }
            if (act) { o.x = gelu_tanh(o.x); o.y = gelu_tanh(o.y); o.z = gelu_tanh(o.z); o.w = gelu_tanh(o.w); }
            if (out_packed) {           // two [32 rows][16 halves] tiles: h0 plane | h1 plane
              uint2 q0, q1;
              pack_split4(o, q0, q1, eamax);
              *reinterpret_cast<uint2*>(wbuf + lane * 32 + g * 8) = q0;
              *reinterpret_cast<uint2*>(wbuf + 1024 + lane * 32 + g * 8) = q1;
            } else {
              *reinterpret_cast<float4*>(wbuf + lane * 64 + ((g ^ ((lane >> 1) & 3)) << 4)) = o;
            }
          }
          const long long tf0 = est ? clock64() : 0;
          if (est) stats[12] += tf0 - tg0;                                   // wait + bias + staging writes
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (est) stats[11] += clock64() - tf0;                             // proxy fence
          if (lane == 0 && !(dbg & 1) && out_packed) {
            const int r0 = (int)(m0 + q * 32);
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];"
                         ::"l"(&tmY), "r"(wbuf_s), "r"(c0), "r"(r0) : "memory");
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];"
                         ::"l"(&tmY), "r"(wbuf_s + 1024u), "r"(N + c0), "r"(r0) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          } else if (lane == 0 && !(dbg & 1)) {
            const int r0 = (int)(m0 + q * 32);
            if (reduce_add)
              asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
                           ::"l"(&tmY), "r"(wbuf_s), "r"(c0), "r"(r0) : "memory");
            else
              asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];"
                           ::"l"(&tmY), "r"(wbuf_s), "r"(c0), "r"(r0) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
        if (est) stats[8] += clock64() - te1;
      } else {
        // Payload GELU: a token's value row gives g, g', g''; tangent rows are scaled by g', the Laplacian
        // row becomes g' lap + g'' sum_t t^2.  The rows of a token sit in different threads, so each 64-column half of the
        // tile goes through shared memory, and the whole epilogue is organised to cost few ISSUE slots (the first version
        // spent 33 instructions per output element and was issue bound at 1.4x the tile's MMA time):
        //   1  warp = (quadrant q, 32-column slab): 32 x 32 piece of main + 2^-11 corr, row-per-lane, into the staging tile
        //   2  thread = (token, column): g, g', g'' of the value row (+ bias) and the sum over the tangent rows of t^2
        //      -> token table {g, g', g'' sum t^2}
        //   3  thread = (row, 8 columns), rows of the tile flat over the group's 256 threads (no idle row classes whatever
        //      C is): out = a + m x with (a, m) = (g, 0) on a value row, (0, g') on a tangent row, (g'' sum t^2, g') on the
        //      Laplacian row; packed as the fp16 pair the down-projection consumes (common.cuh), 16 bytes of the h0 plane +
        //      16 bytes of the h1 plane per thread.
        constexpr int GS = SS_GELU_STRIDE;
        const long long te0 = est ? clock64() : 0;
        mbar_wait_warp(ACC_FULL(grpid), (it >> 1) & 1u);
        if (est) { stats[6] += clock64() - te0; stats[9] += 1; }
        tc_fence_after();
        const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(grpid * 256);
        const int sl = w8 >> 2;
        float* stg = reinterpret_cast<float*>(base + RING_BYTES + SS_BAR_BYTES + grpid * SS_GELU_GROUP_BYTES);
        float* table = stg + TC_BM * GS;                                  // [token][g | g' | g'' sum t^2][64]
        const int tid8 = w8 * 32 + lane;
        const int tpt = rpt / C;
        const int V = (dbg & 32) ? 1 : (tpt * 64 <= 256 ? 1 : (tpt * 32 <= 256 ? 2 : 4));   // columns per phase-2 thread: fewest rounds of 256 items
        const uint32_t inv_c = (65536u + (uint32_t)C - 1u) / (uint32_t)C;   // row / C == (row * inv_c) >> 16 for row < 128, C >= 5
        const int bar_id = 1 + grpid;
        const long long opitch = (long long)N * 4;                       // bytes between payload rows (packed rows = fp32 rows)
#pragma unroll 1
        for (int hf = 0; hf < 2; ++hf) {
          const long long tp0 = est ? clock64() : 0;
          // phase 2 gives every thread the same columns in all its rounds (the column groups per token divide 256): their
          // bias values are fetched here, behind the TMEM loads, not on phase 2's critical path
          float bv[4] = {0.f, 0.f, 0.f, 0.f};
          if (bias != nullptr) {
            const int col = (tid8 % (64 / V)) * V;
#pragma unroll
            for (int v = 0; v < 4; ++v)
              if (v < V) bv[v] = __ldg(bias + nt0 + hf * 64 + col + v);
          }
          {
            uint32_t vm[32], vc[32];
            tc_ld32_nowait(ta + hf * 64 + sl * 32, vm);
            tc_ld32_nowait(ta + TS_BN + hf * 64 + sl * 32, vc);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (hf == 1) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive_cluster(acc_empty_leader);
            }
            float* srow = stg + (q * 32 + lane) * GS + sl * 32;
#pragma unroll
            for (int g = 0; g < 8; ++g)
              *reinterpret_cast<float4*>(srow + 4 * g) =
                  make_float4(fmaf(__uint_as_float(vc[4 * g]), 1.f / H_LO_SCALE, __uint_as_float(vm[4 * g])),
                              fmaf(__uint_as_float(vc[4 * g + 1]), 1.f / H_LO_SCALE, __uint_as_float(vm[4 * g + 1])),
                              fmaf(__uint_as_float(vc[4 * g + 2]), 1.f / H_LO_SCALE, __uint_as_float(vm[4 * g + 2])),
                              fmaf(__uint_as_float(vc[4 * g + 3]), 1.f / H_LO_SCALE, __uint_as_float(vm[4 * g + 3])));
          }
          asm volatile("bar.sync %0, 256;" ::"r"(bar_id) : "memory");
          const long long tp1 = est ? clock64() : 0;
          const int n0 = nt0 + hf * 64;
          if (V == 4) ss_gelu_table<4>(stg, table, bv, tid8, tpt, C);
          else if (V == 2) ss_gelu_table<2>(stg, table, bv, tid8, tpt, C);
          else ss_gelu_table<1>(stg, table, bv, tid8, tpt, C);
          asm volatile("bar.sync %0, 256;" ::"r"(bar_id) : "memory");
          const long long tp2 = est ? clock64() : 0;
          // a quarter-warp = 2 adjacent rows x 4 column groups of one 32-column half row: with 68-float rows its 16-byte
          // loads hit eight different bank groups, and its stores are 64 contiguous bytes per row and plane.  (Two items per
          // trip with all loads in front changed nothing: the epilogue is bound by shared-memory bandwidth -- staging
          // write + two reads + table, 53 KiB per K block on top of the 120 KiB of the MMA operands and TMA -- not by the
          // latency of a thread's trips.)
          const long long plane = (long long)N * 2;
          for (int idx = tid8; idx < rpt * 8; idx += 256) {
            const int row = ((idx >> 4) << 1) | ((idx >> 2) & 1), c8 = ((idx >> 3) & 1) * 32 + (idx & 3) * 8;
            SsGeluItem item;
            ss_gelu_item_load(stg, table, row, c8, C, inv_c, item);
            ss_gelu_item_store(item, reinterpret_cast<uint8_t*>(Y) + (m0 + row) * opitch + (long long)(n0 + c8) * 2, plane, m0 + row < M, eamax);
          }
          if (est) { const long long tp3 = clock64(); stats[13] += tp1 - tp0; stats[14] += tp2 - tp1; stats[15] += tp3 - tp2; }
          if (hf == 0) asm volatile("bar.sync %0, 256;" ::"r"(bar_id) : "memory");   // staging tile and table free for the second half
        }
        asm volatile("bar.sync %0, 256;" ::"r"(bar_id) : "memory");                  // ... and for this group's next tile
      }
    }
    if constexpr (!GELU) {
      if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");        // this warp's TMA stores have landed
      if (out_packed) raise_range_flag(ovf, eamax);
    } else {
      raise_range_flag(ovf, eamax);
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

// fp32 row-major [rows][N] output, box = 16 columns x 32 rows, 64-byte swizzle (the epilogue's staging tiles)
inline int32_t ss_make_out_map(const TcCtx& cx, CUtensorMap* map, const float* ptr, long long rows, int N) {
  if (!cx.encode) return fail(PSIF_E_CUDA, "cuTensorMapEncodeTiled entry point not available%s");
  cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)N * 4};
  cuuint32_t box[2] = {16u, 32u};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = cx.encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(PSIF_E_CUDA, "cuTensorMapEncodeTiled (output) failed (%s%lld)", "", (long long)r);
  return PSIF_OK;
}

// packed output: Y as fp16 [rows][2 N] (h0 plane | h1 plane), box = 16 halves x 32 rows, no swizzle (32-byte rows)
inline int32_t ss_make_out_map_packed(const TcCtx& cx, CUtensorMap* map, const float* ptr, long long rows, int N) {
  if (!cx.encode) return fail(PSIF_E_CUDA, "cuTensorMapEncodeTiled entry point not available%s");
  cuuint64_t dims[2] = {(cuuint64_t)(2 * N), (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)N * 4};
  cuuint32_t box[2] = {16u, 32u};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = cx.encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(PSIF_E_CUDA, "cuTensorMapEncodeTiled (packed output) failed (%s%lld)", "", (long long)r);
  return PSIF_OK;
}

constexpr int SS_NST_PLAIN = 4, SS_NST_GELU = 3;

// Which calls the packed-operand kernel takes (the rest stays with tc_gemm_2cta_kernel): a packed A operand in fp16 mode,
// the output stored through TMA (no residual, or the residual added in place) or the fused payload GELU.
inline bool ss_gemm_takes(bool a_packed, bool f16, int act, const float* res, const float* Y, int C) {
  if (!a_packed || !f16) return false;
  if (act == 2) return res == nullptr && TC_BM / C <= SS_GELU_MAX_TOKENS;
  if (act == 1) return res == nullptr && !(reinterpret_cast<uintptr_t>(Y) & 127);     // plain-row GELU: packed output
  return (res == nullptr || res == Y) && !(reinterpret_cast<uintptr_t>(Y) & 127);
}

// one K pass; mx / mh / ml as tc_gemm builds them
inline int32_t ss_gemm_launch(TcCtx& cx, const CUtensorMap& mx, const CUtensorMap& mh, const CUtensorMap& ml, const float* bias,
                              bool reduce_add, float* Y, long long M, int N, int kk, int C, int act, int rpt, unsigned grid,
                              unsigned* ovf, int a_h1_col, cudaStream_t st, bool kp2) {
  const int dbg = cx.dbg;
  if (!cx.ss_configured) {
    PSIF_CUDA_CHECK(cudaFuncSetAttribute((tc_gemm_ss_kernel<SS_NST_PLAIN, false>), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         ss_smem_bytes(SS_NST_PLAIN, false)));
    PSIF_CUDA_CHECK(cudaFuncSetAttribute((tc_gemm_ss_kernel<SS_NST_GELU, true>), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         ss_smem_bytes(SS_NST_GELU, true)));
    PSIF_CUDA_CHECK(cudaFuncSetAttribute((tc_gemm_ss_kernel<SS_NST_PLAIN, false, true>), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         ss_smem_bytes(SS_NST_PLAIN, false)));
    cx.ss_configured = true;
  }
  // with a packed A operand the GELU epilogues (act != 0: the MLP up-projection) write Y packed as well
  const int out_packed = act == 1 ? 1 : 0;
  CUtensorMap my;
  if (out_packed) PSIF_TRY(ss_make_out_map_packed(cx, &my, Y, M, N));
  else PSIF_TRY(ss_make_out_map(cx, &my, Y, M, N));
  if (act == 2)
    PSIF_LAUNCH((tc_gemm_ss_kernel<SS_NST_GELU, true>), grid, SS_THREADS, ss_smem_bytes(SS_NST_GELU, true), st, mx, mh, ml, bias, Y, M,
                N, kk, C, act, rpt, ovf, my, 0, a_h1_col, dbg, cx.trace, 0);
  else if (kp2)
    PSIF_LAUNCH((tc_gemm_ss_kernel<SS_NST_PLAIN, false, true>), grid, SS_THREADS, ss_smem_bytes(SS_NST_PLAIN, false), st, mx, mh, ml,
                bias, Y, M, N, kk, C, act, rpt, ovf, my, reduce_add ? 1 : 0, a_h1_col, dbg, cx.trace, 0);
  else
    PSIF_LAUNCH((tc_gemm_ss_kernel<SS_NST_PLAIN, false>), grid, SS_THREADS, ss_smem_bytes(SS_NST_PLAIN, false), st, mx, mh, ml, bias,
                Y, M, N, kk, C, act, rpt, ovf, my, reduce_add ? 1 : 0, a_h1_col, dbg, cx.trace, out_packed);
  return PSIF_OK;
}

}  // namespace psif
