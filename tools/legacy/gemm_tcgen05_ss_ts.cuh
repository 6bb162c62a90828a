// Superseded one-CTA-per-tile kernels of round 1 (tf32 split, operands in shared memory / in TMEM).
// Not compiled into libpsiformer_b200.so any more; kept for tools/gemm_bench.py archaeology only.
// They need the PTX wrappers of psiformer_torch_b200/csrc/gemm_tcgen05.cuh in front of them.
template <int BN, int NMAIN>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmWhi,
               const __grid_constant__ CUtensorMap tmWlo, const float* __restrict__ bias, const float* res, float* Y,
               long long M, int N, int K, int C, int act) {
  using Cfg = TcCfg<BN, NMAIN>;
  constexpr int TC_BN = BN, TC_STAGES = Cfg::STAGES, TC_STAGE_BYTES = Cfg::STAGE_BYTES;
  constexpr int TC_B_BYTES = Cfg::B_BYTES, NACC = Cfg::NACC, ACC_COLS = Cfg::ACC_COLS;
  extern __shared__ uint8_t tc_smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + TC_STAGES * TC_STAGE_BYTES);
  // barrier slots: full[S], split[S], empty[S], tfull[2], tempty[2], then the TMEM base address
  const uint32_t bar0 = smem_u32(bars);
  auto FULL = [&](int s) { return bar0 + 8u * s; };
  auto SPLIT = [&](int s) { return bar0 + 8u * (TC_STAGES + s); };
  auto EMPTY = [&](int s) { return bar0 + 8u * (2 * TC_STAGES + s); };
  auto TFULL = [&](int a) { return bar0 + 8u * (3 * TC_STAGES + a); };
  auto TEMPTY = [&](int a) { return bar0 + 8u * (3 * TC_STAGES + 2 + a); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * TC_STAGES + 4);  // (NACC <= 2 barrier pairs)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmWhi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmWlo) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(FULL(s), 1);
      mbar_init(SPLIT(s), 4);
      mbar_init(EMPTY(s), 1);
    }
    for (int a = 0; a < NACC; ++a) {
      mbar_init(TFULL(a), 1);
      mbar_init(TEMPTY(a), 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int tiles_n = N / TC_BN;
  const long long tiles_m = (M + TC_BM - 1) / TC_BM;
  const long long total = tiles_m * tiles_n;
  const int nkb = K / TC_BK;
  const uint32_t smem_base = smem_u32(base);

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (long long tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int m0 = (int)((tile / tiles_n) * TC_BM), n0 = (int)(tile % tiles_n) * TC_BN;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(EMPTY(stage), phase ^ 1);
          const uint32_t sa = smem_base + stage * TC_STAGE_BYTES;
          mbar_arrive_expect_tx(FULL(stage), TC_A_BYTES + 2 * TC_B_BYTES);
          tma_load_2d(sa, &tmX, kb * TC_BK, m0, FULL(stage));
          tma_load_2d(sa + 2 * TC_A_BYTES, &tmWhi, kb * TC_BK, n0, FULL(stage));
          tma_load_2d(sa + 2 * TC_A_BYTES + TC_B_BYTES, &tmWlo, kb * TC_BK, n0, FULL(stage));
          if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = tc_idesc(TC_BM, TC_BN);
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
    for (long long tile = blockIdx.x; tile < total; tile += gridDim.x) {
      mbar_wait(TEMPTY(acc), acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_set = tmem_base + (uint32_t)(acc * ACC_COLS), d_corr = d_set + NMAIN * TC_BN;
      for (int kb = 0; kb < nkb; ++kb) {
        const uint32_t d_main = d_set + (uint32_t)((kb % NMAIN) * TC_BN);
        mbar_wait(FULL(stage), phase);
        mbar_wait(SPLIT(stage), phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = smem_base + stage * TC_STAGE_BYTES;
          const uint32_t a_hi = sa, a_lo = sa + TC_A_BYTES, b_hi = sa + 2 * TC_A_BYTES, b_lo = b_hi + TC_B_BYTES;
#pragma unroll
          for (int k = 0; k < TC_BK / 8; ++k)
            tc_mma_tf32(d_corr, tc_smem_desc(a_lo + k * 32), tc_smem_desc(b_hi + k * 32), idesc, (kb | k) != 0);
#pragma unroll
          for (int k = 0; k < TC_BK / 8; ++k) tc_mma_tf32(d_corr, tc_smem_desc(a_hi + k * 32), tc_smem_desc(b_lo + k * 32), idesc, 1);
#pragma unroll
          for (int k = 0; k < TC_BK / 8; ++k)
            tc_mma_tf32(d_main, tc_smem_desc(a_hi + k * 32), tc_smem_desc(b_hi + k * 32), idesc, (kb >= NMAIN) || (k != 0));
          tc_commit(EMPTY(stage));
          if (kb == nkb - 1) tc_commit(TFULL(acc));
        }
        __syncwarp();
        if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
      }
      if (++acc == NACC) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 4 && warp < 8) {
    int stage = 0;
    uint32_t phase = 0;
    const int t = threadIdx.x - 128;
    for (long long tile = blockIdx.x; tile < total; tile += gridDim.x) {
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(FULL(stage), phase);
        float4* hi = reinterpret_cast<float4*>(base + stage * TC_STAGE_BYTES);
        float4* lo = reinterpret_cast<float4*>(base + stage * TC_STAGE_BYTES + TC_A_BYTES);
#pragma unroll
        for (int j = 0; j < TC_A_BYTES / 16 / 128; ++j) {
          const int idx = t + 128 * j;
          const float4 v = hi[idx];
          float4 h, l;
          uint32_t u;
          asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v.x)); h.x = __uint_as_float(u); l.x = v.x - h.x;
          asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v.y)); h.y = __uint_as_float(u); l.y = v.y - h.y;
          asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v.z)); h.z = __uint_as_float(u); l.z = v.z - h.z;
          asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v.w)); h.w = __uint_as_float(u); l.w = v.w - h.w;
          hi[idx] = h;
          lo[idx] = l;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(SPLIT(stage));
        if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 8) {
    int acc = 0;
    uint32_t acc_phase = 0;
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    for (long long tile = blockIdx.x; tile < total; tile += gridDim.x) {
      const long long m0 = (tile / tiles_n) * TC_BM;
      const int n0 = (int)(tile % tiles_n) * TC_BN;
      mbar_wait(TFULL(acc), acc_phase);
      tc_fence_after();
      const long long r = m0 + q * 32 + lane;
      const bool row_ok = r < M;
      const bool with_bias = bias != nullptr && (C == 1 || (r % C) == 0);
#pragma unroll 1
      for (int ch = 0; ch < TC_BN / 32; ++ch) {
        uint32_t v[32], vc[32];
        const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * ACC_COLS + ch * 32);
        tc_ld32(ta + NMAIN * TC_BN, v);          // correction first, then the main partial sums
#pragma unroll
        for (int mj = 0; mj < NMAIN; ++mj) {
          tc_ld32(ta + mj * TC_BN, vc);
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) + __uint_as_float(vc[e]));
        }
        if (row_ok) {
          const int c0 = n0 + ch * 32;
          float* yp = Y + r * (long long)N + c0;
          const float* rp = res ? res + r * (long long)N + c0 : nullptr;
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            float4 o = make_float4(__uint_as_float(v[4 * g]), __uint_as_float(v[4 * g + 1]), __uint_as_float(v[4 * g + 2]),
                                   __uint_as_float(v[4 * g + 3]));
            if (with_bias) {
              const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + c0 + 4 * g));
              o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
            }
            if (act) { o.x = gelu_tanh(o.x); o.y = gelu_tanh(o.y); o.z = gelu_tanh(o.z); o.w = gelu_tanh(o.w); }
            if (rp) {
              const float4 rr = *reinterpret_cast<const float4*>(rp + 4 * g);
              o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w;
            }
            *reinterpret_cast<float4*>(yp + 4 * g) = o;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(TEMPTY(acc));
      if (++acc == NACC) { acc = 0; acc_phase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

// CL = thread-block-cluster size along M: the CL CTAs of a cluster work on CL consecutive row tiles of the SAME
// column tile; each loads 1/CL of the W_hi / W_lo tiles and TMA-multicasts it to all of them, dividing the
// L2 -> SM weight traffic by CL (ncu, round 1: the un-clustered kernel moved 48 KiB per K block per SM and sat at
// ~55 % of L2 throughput with the tensor pipe 54 % busy).
template <int NMAIN, int CL>
__global__ void __launch_bounds__(TS_THREADS, 1)
tc_gemm_ts_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmWhi,
                  const __grid_constant__ CUtensorMap tmWlo, const float* __restrict__ bias, const float* res, float* Y,
                  long long M, int N, int K, int C, int act, int dbg, int pf, long long* trace, int rpt) {
  // rpt = rows per tile (<= 128): consecutive row tiles start rpt rows apart, so that with rpt a multiple of the
  // payload channel count C every tile holds whole tokens (act == 2); rows rpt..127 of a tile are computed and dropped
  constexpr int ACC_COLS = (NMAIN + 1) * TS_BN;
  constexpr int TA_STAGES = (512 - ACC_COLS) / 64;           // 4 (NMAIN = 1) or 2 (NMAIN = 2)
  static_assert(TA_STAGES >= 2, "need at least two TMEM operand stages");
  extern __shared__ uint8_t tc_smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + TS_SM_STAGES * TS_STAGE_BYTES);
  const uint32_t bar0 = smem_u32(bars);
  // barrier slots: FULL[4] EMPTY_S[4] SPLIT[4] EMPTY_A[4] TFULL TEMPTY, then the TMEM base address
  auto FULL = [&](int s) { return bar0 + 8u * s; };
  auto EMPTY_S = [&](int s) { return bar0 + 8u * (4 + s); };
  auto SPLIT = [&](int a) { return bar0 + 8u * (8 + a); };
  auto EMPTY_A = [&](int a) { return bar0 + 8u * (12 + a); };
  // TFULL: all MMAs of the tile done.  CEMPTY / TEMPTY: correction / main accumulators drained (the epilogue drains
  // the correction accumulator first, so the next tile's 8 leading correction MMAs start after half of the drain and
  // the main accumulator is usually free by the time its first MMA is issued)
  const uint32_t TFULL = bar0 + 8u * 16, TEMPTY = bar0 + 8u * 17, CEMPTY = bar0 + 8u * 18;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 19);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // optional timeline of CTA 0 (tests/tools only): trace[role * 512 + i] = clock64 at event i of that role
  const bool tracing = trace != nullptr && blockIdx.x == 0;
  int tcount = 0;
#define PSIF_TRACE(role) do { if (tracing && lane == 0 && tcount < 512) trace[(role) * 512 + tcount] = clock64(); } while (0)
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmWhi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmWlo) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < TS_SM_STAGES; ++s) { mbar_init(FULL(s), 1); mbar_init(EMPTY_S(s), CL); }
    for (int a = 0; a < TA_STAGES; ++a) { mbar_init(SPLIT(a), 4); mbar_init(EMPTY_A(a), 1); }
    mbar_init(TFULL, 1);
    mbar_init(TEMPTY, 8);
    mbar_init(CEMPTY, 8);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  if (CL > 1) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int tiles_n = N / TS_BN;
  const long long tiles_m = (M + rpt - 1) / rpt;
  const long long groups = ((tiles_m + CL - 1) / CL) * tiles_n;   // a group = CL row tiles x 1 column tile
  const int nkb = K / TC_BK;
  const uint32_t smem_base = smem_u32(base);
  const uint32_t crank = CL > 1 ? cluster_ctarank() : 0u;
  const long long g0 = CL > 1 ? (long long)cluster_id_x() : (long long)blockIdx.x;
  const long long gstep = CL > 1 ? (long long)cluster_count_x() : (long long)gridDim.x;
  constexpr uint16_t kMask = (uint16_t)((1u << CL) - 1u);
  constexpr int BROWS = TS_BN / CL;                      // weight-tile rows this CTA fetches

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      // The X tiles stream from HBM (each is read once), the weight tiles from L2.  With 4 smem stages the HBM
      // latency is not covered (timing experiments, round 1), so X is prefetched into L2 `pf` K blocks ahead.
      long long pgrp = g0;
      int pkb = 0;
      for (int i = 0; i < pf && pgrp < groups; ++i) {
        tma_prefetch_2d(&tmX, pkb * TC_BK, (int)(((pgrp / tiles_n) * CL + crank) * rpt));
        if (++pkb == nkb) { pkb = 0; pgrp += gstep; }
      }
      for (long long grp = g0; grp < groups; grp += gstep) {
        const int m0 = (int)(((grp / tiles_n) * CL + crank) * rpt), n0 = (int)(grp % tiles_n) * TS_BN;
        for (int kb = 0; kb < nkb; ++kb) {
          if (pf > 0 && pgrp < groups) {
            tma_prefetch_2d(&tmX, pkb * TC_BK, (int)(((pgrp / tiles_n) * CL + crank) * rpt));
            if (++pkb == nkb) { pkb = 0; pgrp += gstep; }
          }
          mbar_wait(EMPTY_S(stage), phase ^ 1);
          if (tracing && tcount < 512) trace[0 * 512 + tcount] = clock64();
          const uint32_t sa = smem_base + stage * TS_STAGE_BYTES;
          mbar_arrive_expect_tx(FULL(stage), TC_A_BYTES + ((dbg & 8) ? 1 : 2) * TS_B_BYTES);
          tma_load_2d(sa, &tmX, kb * TC_BK, m0, FULL(stage));
          if (CL > 1) {
            const uint32_t off = crank * (BROWS * 128);
            tma_load_2d_mc(sa + TC_A_BYTES + off, &tmWhi, kb * TC_BK, n0 + crank * BROWS, FULL(stage), kMask);
            tma_load_2d_mc(sa + TC_A_BYTES + TS_B_BYTES + off, &tmWlo, kb * TC_BK, n0 + crank * BROWS, FULL(stage), kMask);
          } else {
            tma_load_2d(sa + TC_A_BYTES, &tmWhi, kb * TC_BK, n0, FULL(stage));
            if (!(dbg & 8)) tma_load_2d(sa + TC_A_BYTES + TS_B_BYTES, &tmWlo, kb * TC_BK, n0, FULL(stage));
          }
          if (tracing && tcount < 512) trace[1 * 512 + tcount] = clock64();
          ++tcount;
          if (++stage == TS_SM_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // MMA issuer: a single elected thread runs the whole loop (no warp-level operation inside).  The barriers of
    // K block kb+1 are awaited after the 8 correction MMAs of K block kb have been queued and before its 4 main
    // MMAs, so the tensor pipe never drains while this thread sits in a try_wait.
    constexpr uint32_t idesc = tc_idesc(TC_BM, TS_BN);
    if (elect_one()) {
      int stage = 0, ta = 0;
      uint32_t phase = 0, ta_phase = 0, acc_phase = 0;
      const uint32_t d_corr = tmem_base + NMAIN * TS_BN;
      bool ready = false;   // barriers of the K block about to be issued already awaited?
      for (long long grp = g0; grp < groups; grp += gstep) {
        mbar_wait(CEMPTY, acc_phase ^ 1);
        tc_fence_after();
        for (int kb = 0; kb < nkb; ++kb) {
          // SPLIT implies FULL: the splitter warps wait for the stage's TMA transaction (X and both weight tiles)
          // before they convert and arrive, so one barrier test per K block is enough here
          if (!ready) {
            mbar_wait(SPLIT(ta), ta_phase);
            tc_fence_after();
          }
          if (tracing && tcount < 512) trace[5 * 512 + tcount] = clock64();
          const uint32_t sa = smem_base + stage * TS_STAGE_BYTES;
          const uint32_t b_hi = sa + TC_A_BYTES, b_lo = b_hi + TS_B_BYTES;
          const uint32_t a_hi = tmem_base + ACC_COLS + ta * 64, a_lo = a_hi + 32;
          const uint32_t d_main = tmem_base + (uint32_t)((kb % NMAIN) * TS_BN);
          if (!(dbg & 1)) {
#pragma unroll
            for (int k = 0; k < TC_BK / 8; ++k)
              tc_mma_tf32_ts(d_corr, a_lo + 8 * k, tc_smem_desc(b_hi + k * 32), idesc, (kb | k) != 0);
#pragma unroll
            for (int k = 0; k < TC_BK / 8; ++k) tc_mma_tf32_ts(d_corr, a_hi + 8 * k, tc_smem_desc(b_lo + k * 32), idesc, 1);
          }
          // look ahead: next K block of this tile (the first K block of the next tile also needs TEMPTY, so it is
          // awaited at the top of the tile loop instead)
          int nstage = stage + 1, nta = ta + 1;
          uint32_t nphase = phase, nta_phase = ta_phase;
          if (nstage == TS_SM_STAGES) { nstage = 0; nphase ^= 1; }
          if (nta == TA_STAGES) { nta = 0; nta_phase ^= 1; }
          ready = false;
          if (kb + 1 < nkb) {
            mbar_wait(SPLIT(nta), nta_phase);
            tc_fence_after();
            ready = true;
          }
          if (kb == 0) {
            mbar_wait(TEMPTY, acc_phase ^ 1);
            tc_fence_after();
          }
#pragma unroll
          for (int k = 0; k < TC_BK / 8; ++k)
            tc_mma_tf32_ts(d_main, a_hi + 8 * k, tc_smem_desc(b_hi + k * 32), idesc, (kb >= NMAIN) || (k != 0));
          if (CL > 1) tc_commit_mc(EMPTY_S(stage), kMask); else tc_commit(EMPTY_S(stage));
          tc_commit(EMPTY_A(ta));
          if (kb == nkb - 1) tc_commit(TFULL);
          if (tracing && tcount < 512) { trace[6 * 512 + tcount] = clock64(); ++tcount; }
          stage = nstage; phase = nphase; ta = nta; ta_phase = nta_phase;
        }
        acc_phase ^= 1;
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // splitter: thread = tile row; raw X row (128 B, 128B-swizzled) -> hi / lo -> TMEM lanes of this warp's quarter
    int stage = 0, ta = 0;
    uint32_t phase = 0, ta_phase = 0;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    for (long long grp = g0; grp < groups; grp += gstep) {
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait_warp(FULL(stage), phase);
        if (warp == 4) PSIF_TRACE(2);
        mbar_wait_warp(EMPTY_A(ta), ta_phase ^ 1);
        if (warp == 4) PSIF_TRACE(7);
        tc_fence_after();
        if (dbg & 4) {
          __syncwarp();
          if (lane == 0) mbar_arrive(SPLIT(ta));
          if (++stage == TS_SM_STAGES) { stage = 0; phase ^= 1; }
          if (++ta == TA_STAGES) { ta = 0; ta_phase ^= 1; }
          continue;
        }
        const uint8_t* rp = base + stage * TS_STAGE_BYTES + row * 128;
        uint32_t hi[32], lo[32];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 v = *reinterpret_cast<const float4*>(rp + ((c ^ (row & 7)) << 4));
          const float vv[4] = {v.x, v.y, v.z, v.w};
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
        const uint32_t ta_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ACC_COLS + ta * 64);
        tc_st32(ta_addr, hi);
        tc_st32(ta_addr + 32, lo);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(SPLIT(ta));
        if (warp == 4) PSIF_TRACE(3);
        ++tcount;
        if (++stage == TS_SM_STAGES) { stage = 0; phase ^= 1; }
        if (++ta == TA_STAGES) { ta = 0; ta_phase ^= 1; }
      }
    }
  } else if (warp >= 8) {
    uint32_t acc_phase = 0;
    const int q = warp & 3, half = (warp - 8) >> 2;   // lane quarter, 64-column half of the 128-wide tile
    for (long long grp = g0; grp < groups; grp += gstep) {
      const long long m0 = ((grp / tiles_n) * CL + crank) * rpt;
      const int n0 = (int)(grp % tiles_n) * TS_BN + half * 64;
      const long long r = m0 + q * 32 + lane;
      const bool row_ok = r < M && q * 32 + lane < rpt;
      const bool with_bias = bias != nullptr && (C == 1 || (r % C) == 0);
      mbar_wait_warp(TFULL, acc_phase);
      if (warp == 8) PSIF_TRACE(8);
      tc_fence_after();
      // drain: correction first, then the main partial sums, 2 x 32 columns -> 64 registers, then free TMEM
      uint32_t v[2][32];
      const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * 64);
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) tc_ld32_nowait(ta + NMAIN * TS_BN + ch * 32, v[ch]);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(CEMPTY);
#pragma unroll
      for (int mj = 0; mj < NMAIN; ++mj) {
        uint32_t w[2][32];
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) tc_ld32_nowait(ta + mj * TS_BN + ch * 32, w[ch]);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int ch = 0; ch < 2; ++ch)
#pragma unroll
          for (int e = 0; e < 32; ++e) v[ch][e] = __float_as_uint(__uint_as_float(v[ch][e]) + __uint_as_float(w[ch][e]));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(TEMPTY);
      if (warp == 8) PSIF_TRACE(9);
      acc_phase ^= 1;
      if (act == 2) {
        // GELU on the payload: a token's value row gives g, g', g''; its tangent rows are scaled by
        // g' and its Laplacian row becomes g' lap + g'' sum_t t^2.  Rows of a token sit in different threads, so
        // the tile goes through shared memory 32 columns at a time and is re-read with thread = column.
        float* buf = reinterpret_cast<float*>(base + TS_SM_STAGES * TS_STAGE_BYTES + 256) + half * (TC_BM * TS_GELU_STRIDE);
        const int lr = q * 32 + lane;
        const int tpt = rpt / C;
        const int bar_id = 1 + half;
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
#pragma unroll
          for (int e = 0; e < 32; ++e) buf[lr * TS_GELU_STRIDE + e] = __uint_as_float(v[ch][e]);
          asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
          const int col = n0 + ch * 32 + lane;
          const float bcol = bias ? __ldg(bias + col) : 0.f;
          // warp = every 4th token of the tile, lane = column.  The kernel is bound by shared-memory bandwidth (TMA
          // writes + tensor-core operand reads + splitter), so the staging tile is read exactly once
          for (int t = q; t < tpt; t += 4) {
            const long long gr = m0 + (long long)t * C;
            if (gr >= M) break;
            const float* bp = buf + t * C * TS_GELU_STRIDE + lane;
            float* yp = Y + gr * (long long)N + col;
            const float v0 = bp[0];
            const float vl = bp[(C - 1) * TS_GELU_STRIDE];
            float g, g1, g2;
            gelu_tanh_d2(v0 + bcol, g, g1, g2);
            yp[0] = g;
            float ss = 0.f;
            for (int c = 1; c < C - 1; c += 8) {
              float tv[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) tv[j] = c + j < C - 1 ? bp[(c + j) * TS_GELU_STRIDE] : 0.f;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                ss = fmaf(tv[j], tv[j], ss);
                if (c + j < C - 1) yp[(long long)(c + j) * N] = g1 * tv[j];
              }
            }
            yp[(long long)(C - 1) * N] = fmaf(g1, vl, g2 * ss);
          }
        }
      } else if (row_ok && !(dbg & 2)) {
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          const int c0 = n0 + ch * 32;
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
      if (warp == 8) PSIF_TRACE(10);
      ++tcount;
    }
  }
  tc_fence_before();
  if (CL > 1) cluster_sync_all(); else __syncthreads();   // peers may still multicast into this CTA until they are done
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
#undef PSIF_TRACE
}

