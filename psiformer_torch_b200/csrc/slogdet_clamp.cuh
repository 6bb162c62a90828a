// Derivatives of log|psi| on walkers where the reference's singular-value clamp is ACTIVE (logdet_matmul.py:50-51).
//
// det_combine_kernel (slogdet.cuh) differentiates log|det A| = -log|det A^-1| through a Gauss-Jordan inverse, which is
// what torch.linalg.svd -> clamp(s, 1e-6) -> log reduces to as long as no singular value is below the clamp.  A block
// with sigma_min < 1e-6 gets its VALUE from the clamped singular values there, but its tangent / Laplacian terms would
// still be those of the unclamped function; the kernel raises PSIF_ST_CLAMP_SUSPECT on such walkers (2 % of raw N(0, I)
// walkers of N2, none of the pinned fixtures).  In the reference the clamped directions carry no gradient
// (torch.clamp) and the second derivative picks up the second-order perturbation of the remaining singular values.
//
// det_clamp_fixup_kernel runs behind det_combine_kernel in energy mode: one warp per walker, warps of unflagged walkers
// leave at once.  A flagged walker is recomputed from the closed forms of logdet_math.cuh (the ones behind the
// twice-differentiable LogDetMatmul op):  per block  f = sum_i log max(s_i, 1e-6),  per tangent channel  <G, dA> and
// <H[dA], dA>,  Laplacian channel  <G, lap A>  -- lanes over the 2 K blocks -- then the same combine over determinants
// and the same local-energy assembly as det_combine_kernel, and the walker's contribution to the energy statistics is
// replaced.  Blocks of the walker that are not clamped go through the same functions' Gauss-Jordan branch.
#pragma once
#include "logdet_math.cuh"
#include "slogdet.cuh"

namespace psif {

__host__ __device__ inline int detfix_smem_doubles(int K, int T) { return 6 * K + 2 * K * T + 2 * K + T + 8; }

__global__ void __launch_bounds__(32)
det_clamp_fixup_kernel(DetArgs a) {
  extern __shared__ double fsm[];
  const long long b = blockIdx.x;
  if (b >= a.B || a.status == nullptr) return;
  const uint32_t st_old = a.status[b];
  if (!(st_old & PSIF_ST_CLAMP_SUSPECT)) return;
  const int lane = threadIdx.x;
  const int K = a.K, T = a.C - 2;
  double* ell = fsm;                 // [2][K]
  double* sgn = ell + 2 * K;         // [2][K]
  double* lapt = sgn + 2 * K;        // [2][K]
  double* q = lapt + 2 * K;          // [K]
  double* ck = q + K;                // [K]
  double* G = ck + K;                // [T]
  double* misc = G + T;              // [8]
  double* gs = misc + 8;             // [2][K][T]

  for (int idx = lane; idx < 2 * K; idx += 32) {
    const int sg = idx / K, k = idx - sg * K;
    const int n = a.n[sg];
    if (n <= 0) {                      // an empty spin channel: det = 1, no derivatives
      ell[idx] = 0.0; sgn[idx] = 1.0; lapt[idx] = 0.0;
      for (int c = 0; c < T; ++c) gs[(size_t)idx * T + c] = 0.0;
      continue;
    }
    const float* base = a.phi[sg] + b * a.wstride[sg] + (long long)k * a.kstride[sg];
    const long long is = a.istride[sg];
    double A[LD_MAXN * LD_MAXN], E[LD_MAXN * LD_MAXN], H[LD_MAXN * LD_MAXN];
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) A[i * n + j] = (double)__ldg(base + i * is + j) + ((i == j) ? LD_DET_JITTER : 0.0);
    LdBlock blk;
    ld_factor(A, n, blk);
    ell[idx] = blk.logdet;
    sgn[idx] = blk.sign;
    double lapacc = 0.0;
    for (int c = 0; c < T; ++c) {
      const float* pc = base + (long long)(1 + c) * a.cstride;
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) E[i * n + j] = (double)__ldg(pc + i * is + j);
      gs[(size_t)idx * T + c] = ld_grad_dot(blk, E);
      ld_hess_apply(blk, E, H);
      double h = 0.0;
      for (int e = 0; e < n * n; ++e) h += H[e] * E[e];
      lapacc += h;
    }
    {
      const float* pc = base + (long long)(a.C - 1) * a.cstride;
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) E[i * n + j] = (double)__ldg(pc + i * is + j);
      lapacc += ld_grad_dot(blk, E);
    }
    lapt[idx] = lapacc;
  }
  __syncwarp();
  for (int k = lane; k < K; k += 32) {
    double s2 = 0.0;
    for (int c = 0; c < T; ++c) {
      const double v = gs[(size_t)k * T + c] + gs[(size_t)(K + k) * T + c];
      s2 += v * v;
    }
    q[k] = lapt[k] + lapt[K + k] + s2;
  }
  __syncwarp();
  if (lane == 0) {
    // the combine of det_combine_kernel, literally (logdet_matmul.py:58-69)
    double m0 = -INFINITY, m1 = -INFINITY;
    bool nan_in = false;
    for (int kk = 0; kk < K; ++kk) {
      m0 = fmax(m0, ell[kk]);
      m1 = fmax(m1, ell[K + kk]);
      nan_in = nan_in || (ell[kk] != ell[kk]) || (ell[K + kk] != ell[K + kk]);
    }
    double S = 0.0;
    for (int kk = 0; kk < K; ++kk) {
      const double D = (double)__ldg(a.w + kk) * sgn[kk] * sgn[K + kk] * exp(ell[kk] - m0 + ell[K + kk] - m1);
      ck[kk] = D;
      S += D;
    }
    const double invS = 1.0 / S;
    for (int kk = 0; kk < K; ++kk) ck[kk] *= invS;
    misc[0] = (nan_in || S != S) ? (double)NAN : log(fmax(fabs(S), kOutputFloor)) + m0 + m1;
    misc[1] = (S > 0.0) ? 1.0 : ((S < 0.0) ? -1.0 : 0.0);
    misc[2] = fabs(S);
  }
  __syncwarp();
  for (int c = lane; c < T; c += 32) {
    double g = 0.0;
    for (int kk = 0; kk < K; ++kk) g += ck[kk] * (gs[(size_t)kk * T + c] + gs[(size_t)(K + kk) * T + c]);
    G[c] = g;
    if (a.grad) a.grad[b * T + c] = (float)(g + (a.jgrad ? a.jgrad[b * T + c] : 0.0));
  }
  __syncwarp();
  if (lane == 0) {
    uint32_t st = st_old & ~(uint32_t)(PSIF_ST_NONFINITE_LOGDET | PSIF_ST_NONFINITE_ELOC | PSIF_ST_FLOOR);
    const double jv = a.jval ? a.jval[b] : 0.0;
    const double logdet = misc[0];
    if (!isfinite(logdet)) st |= PSIF_ST_NONFINITE_LOGDET;
    if (misc[2] < kOutputFloor) st |= PSIF_ST_FLOOR;
    a.logabs[b] = (float)(logdet + jv);
    if (a.sign) a.sign[b] = (float)misc[1];
    double tsum = 0.0;
    for (int kk = 0; kk < K; ++kk) tsum += ck[kk] * q[kk];
    double g2 = 0.0, gj2 = 0.0;
    for (int c = 0; c < T; ++c) {
      const double g = G[c];
      const double gt = g + (a.jgrad ? a.jgrad[b * T + c] : 0.0);
      g2 += g * g;
      gj2 += gt * gt;
    }
    const double lap = tsum - g2 + (a.jlap ? a.jlap[b] : 0.0);
    const double v = a.pot ? a.pot[b] : 0.0;
    const double e = -0.5 * (lap + gj2) + v;
    if (!isfinite(e)) st |= PSIF_ST_NONFINITE_ELOC;
    const double e_old = a.e_loc ? (double)a.e_loc[b] : 0.0;
    if (a.e_loc) a.e_loc[b] = (float)e;
    if (a.lap) a.lap[b] = (float)lap;
    if (a.pot_out) a.pot_out[b] = (float)v;
    a.status[b] = st;
    if (a.accum != nullptr && a.e_loc != nullptr) {
      // replace this walker's contribution to {sum E, sum E^2, n} (det_combine_kernel added its value before rounding it to
      // the float read back here: a 1e-7 relative difference on one walker of the statistics)
      const bool was = !(st_old & (PSIF_ST_NONFINITE_LOGDET | PSIF_ST_NONFINITE_ELOC));
      const bool is = !(st & (PSIF_ST_NONFINITE_LOGDET | PSIF_ST_NONFINITE_ELOC));
      const double d0 = (is ? e : 0.0) - (was ? e_old : 0.0);
      const double d1 = (is ? e * e : 0.0) - (was ? e_old * e_old : 0.0);
      const double d2 = (is ? 1.0 : 0.0) - (was ? 1.0 : 0.0);
      atomicAdd(a.accum + 0, d0); atomicAdd(a.accum + 1, d1);
      if (d2 != 0.0) atomicAdd(a.accum + 2, d2);
    }
  }
}

inline int32_t det_clamp_fixup_launch(const DetArgs& a, cudaStream_t st) {
  if (a.B <= 0 || a.status == nullptr) return PSIF_OK;
  if (a.B > 0x7fffffffLL) return fail(PSIF_E_INVALID, "slogdet: too many walkers for one launch%s");
  const size_t smem = (size_t)detfix_smem_doubles(a.K, a.C - 2) * sizeof(double);
  if (smem > 200 * 1024) return fail(PSIF_E_INVALID, "slogdet: shared memory budget exceeded%s");
  size_t& configured = dev_smem_cfg().detfix;
  if (smem > 48 * 1024 && smem > configured) {
    PSIF_CUDA_CHECK(cudaFuncSetAttribute(det_clamp_fixup_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  PSIF_LAUNCH(det_clamp_fixup_kernel, (unsigned)a.B, 32, smem, st, a);
  return PSIF_OK;
}

}  // namespace psif
