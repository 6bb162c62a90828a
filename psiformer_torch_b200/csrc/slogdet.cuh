// Multi-determinant slogdet with derivatives, Jastrow, Coulomb potential and the local-energy
// assembly.  Replaces logdet_matmul.py:35-70 (SVD based) + the autograd Laplacian through it
// (hamiltonian.py:56-95) + jastrow.py:67-87 + hamiltonian.py:15-35,52-54.
//
//   one thread per (walker, spin, determinant): A = Phi + 1e-4 I, Gauss-Jordan inverse in fp64 ->
//     l = log|det A|, sign, and per tangent channel c:  tr(A^-1 d_c A),  tr((A^-1 d_c A)^2),
//     Laplacian channel: tr(A^-1 lap A)                                   
//   then a per-walker combine over determinants:
//     D_k = w_k s_k exp(l_k - m_up - m_dn), S = sum D_k, c_k = D_k / S,
//     grad log|S| = sum c_k grad l_k,  lap log|S| = sum c_k (lap l_k + |grad l_k|^2) - |grad log|S||^2
//   and E_L = -1/2 (lap log psi + |grad log psi|^2) + V with log psi = log|S| + shifts + J.
//
// The reference clamps singular values at 1e-6 (logdet_matmul.py:50-51).  sigma_min >= 1/||A^-1||_F,
// so blocks with ||A^-1||_F <= 1e6 are provably unclamped; the rest get their singular values from a
// one-sided Jacobi sweep in the same thread and the clamp is applied exactly (value); walkers with an ACTIVE clamp are
// flagged PSIF_ST_CLAMP_SUSPECT and, in energy mode, recomputed by det_clamp_fixup_kernel (slogdet_clamp.cuh): the
// clamped directions have zero gradient in the reference, which the smooth formulas below cannot reproduce.
#pragma once
#include "common.cuh"
#include "smallmat.cuh"

namespace psif {

struct DetArgs {
  const float* phi[2];     // [0] spin-up blocks, [1] spin-down blocks
  long long wstride[2];    // floats between walkers
  long long kstride[2];    // floats between determinants
  long long istride[2];    // floats between matrix rows (electrons)
  long long cstride;       // floats between payload channels
  int C, K, n[2];
  const float* w;          // [K] determinant weights
  const double* jval;      // [B]      Jastrow value      (may be null -> 0)
  const double* jgrad;     // [B][3N]  Jastrow gradient   (DERIV only)
  const double* jlap;      // [B]
  const double* pot;       // [B]
  float* e_loc; float* logabs; float* sign; float* grad; float* lap; float* pot_out;
  uint32_t* status;
  double* accum;
  long long B;
  int tpw, wpb;            // threads per walker (2K), walkers per block
};

__host__ __device__ inline int det_smem_doubles_per_walker(int K, int T) {
  // ell[2K] sgn[2K] lapt[2K] q[K] ck[K] G[T] misc[8] gs[2K*T]
  return 2 * K * 3 + 2 * K + T + 8 + 2 * K * T;
}

template <int NM, bool DERIV>
__global__ void __launch_bounds__(128)
det_combine_kernel(DetArgs a) {
  extern __shared__ double dsm[];
  constexpr bool XSMEM = DERIV && (NM > 5);
  const int K = a.K;
  const int T = DERIV ? a.C - 2 : 0;
  const int tid = threadIdx.x;
  const int wl = tid / a.tpw, rem = tid - wl * a.tpw;
  const int sg = rem / K, k = rem - sg * K;
  const long long b = (long long)blockIdx.x * a.wpb + wl;
  const bool in_slot = wl < a.wpb;          // blockDim is rounded up to a warp multiple
  const bool active = in_slot && b < a.B;
  const int per_w = det_smem_doubles_per_walker(K, T);
  double* W = dsm + (size_t)(in_slot ? wl : 0) * per_w;
  double* ell = W;                 // [2][K]
  double* sgn = ell + 2 * K;       // [2][K]
  double* lapt = sgn + 2 * K;      // [2][K]
  double* q = lapt + 2 * K;        // [K]
  double* ck = q + K;              // [K]
  double* G = ck + K;              // [T]
  double* misc = G + T;            // [8]
  double* gs = misc + 8;           // [2][K][T]
  double* Xs = dsm + (size_t)a.wpb * per_w;  // [NM*NM][blockDim] (XSMEM only)

  uint32_t my_flags = 0;
  if (active) {
    const int n = a.n[sg];
    const float* base = a.phi[sg] + b * a.wstride[sg] + (long long)k * a.kstride[sg];
    const long long is = a.istride[sg];
    double X[NM * NM];
#pragma unroll
    for (int i = 0; i < NM; ++i)
#pragma unroll
      for (int j = 0; j < NM; ++j) {
        double v = (i == j) ? 1.0 : 0.0;
        if (i < n && j < n) v = (double)__ldg(base + i * is + j) + ((i == j) ? kDetJitter : 0.0);
        X[i * NM + j] = v;
      }
    double ld, sgv, minpiv;
    gj_inverse<NM>(X, ld, sgv, minpiv);
    double fro2 = 0.0;
#pragma unroll
    for (int e = 0; e < NM * NM; ++e) fro2 += X[e] * X[e];
    fro2 -= (double)(NM - n);
    if (!(fro2 <= 1e12)) {
      // possible singular value below 1e-6: get them exactly and apply the reference clamp
      double Acp[NM * NM], sv[NM];
#pragma unroll
      for (int i = 0; i < NM; ++i)
#pragma unroll
        for (int j = 0; j < NM; ++j) {
          double v = (i == j) ? 1.0 : 0.0;
          if (i < n && j < n) v = (double)__ldg(base + i * is + j) + ((i == j) ? kDetJitter : 0.0);
          Acp[i * NM + j] = v;
        }
      jacobi_singular_values<NM>(Acp, sv);
      double ldc = 0.0;
      bool clamped = false;
#pragma unroll
      for (int i = 0; i < NM; ++i) {
        if (sv[i] < kMinSingular) clamped = true;
        ldc += log(fmax(sv[i], kMinSingular));
      }
      if (clamped) {
        ld = ldc;
        my_flags |= PSIF_ST_CLAMP_SUSPECT;
        if (sgv == 0.0) sgv = 1.0;
      }
    }
    ell[sg * K + k] = ld;
    sgn[sg * K + k] = sgv;

    if (DERIV) {
      if (XSMEM) {
#pragma unroll
        for (int e = 0; e < NM * NM; ++e) Xs[(size_t)e * blockDim.x + tid] = X[e];
      }
#define PSIF_XE(r, i) (XSMEM ? Xs[(size_t)((r) * NM + (i)) * blockDim.x + tid] : X[(r) * NM + (i)])
      double lapacc = 0.0;
      double* gout = gs + (size_t)(sg * K + k) * T;
      for (int c = 0; c < T; ++c) {
        const float* pc = base + (long long)(1 + c) * a.cstride;
        double M[NM * NM];
#pragma unroll
        for (int e = 0; e < NM * NM; ++e) M[e] = 0.0;
#pragma unroll
        for (int i = 0; i < NM; ++i) {
          if (i < n) {
#pragma unroll
            for (int j = 0; j < NM; ++j) {
              if (j < n) {
                const double da = (double)__ldg(pc + i * is + j);
#pragma unroll
                for (int r = 0; r < NM; ++r) M[r * NM + j] += PSIF_XE(r, i) * da;
              }
            }
          }
        }
        double tr = 0.0, tr2 = 0.0;
#pragma unroll
        for (int r = 0; r < NM; ++r) {
          tr += M[r * NM + r];
#pragma unroll
          for (int j = 0; j < NM; ++j) tr2 += M[r * NM + j] * M[j * NM + r];
        }
        gout[c] = tr;
        lapacc -= tr2;
      }
      {
        const float* pc = base + (long long)(a.C - 1) * a.cstride;
#pragma unroll
        for (int i = 0; i < NM; ++i)
          if (i < n) {
#pragma unroll
            for (int r = 0; r < NM; ++r)
              if (r < n) lapacc += PSIF_XE(r, i) * (double)__ldg(pc + i * is + r);
          }
      }
#undef PSIF_XE
      lapt[sg * K + k] = lapacc;
    }
  }
  // clamp flags of all blocks of a walker are OR-ed through shared memory
  unsigned* flagw = reinterpret_cast<unsigned*>(misc + 6);
  if (in_slot && rem == 0) *flagw = 0u;
  __syncthreads();
  if (active && my_flags) atomicOr(flagw, my_flags);
  if (active && DERIV && sg == 0) {
    double s2 = 0.0;
    const double* gu = gs + (size_t)k * T;
    const double* gd = gs + (size_t)(K + k) * T;
    for (int c = 0; c < T; ++c) {
      const double v = gu[c] + gd[c];
      s2 += v * v;
    }
    q[k] = lapt[k] + lapt[K + k] + s2;
  }
  if (active && rem == a.tpw - 1) {
    // serial over K <= 64: shifts, weighted sum, coefficients  (logdet_matmul.py:58-69)
    double m0 = -INFINITY, m1 = -INFINITY;
    bool nan_in = false;        // fmax() drops NaN operands; torch.max / torch.clamp (logdet_matmul.py:58-68) propagate them
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
  __syncthreads();
  if (active && DERIV) {
    for (int c = rem; c < T; c += a.tpw) {
      double g = 0.0;
      for (int kk = 0; kk < K; ++kk) g += ck[kk] * (gs[(size_t)kk * T + c] + gs[(size_t)(K + kk) * T + c]);
      G[c] = g;
      if (a.grad) a.grad[b * T + c] = (float)(g + (a.jgrad ? a.jgrad[b * T + c] : 0.0));
    }
  }
  __syncthreads();
  double e_ok = 0.0, e2_ok = 0.0, n_ok = 0.0;
  if (active && rem == 0) {
    uint32_t st = *flagw;
    const double jv = a.jval ? a.jval[b] : 0.0;
    const double logdet = misc[0];
    if (!isfinite(logdet)) st |= PSIF_ST_NONFINITE_LOGDET;
    if (misc[2] < kOutputFloor) st |= PSIF_ST_FLOOR;
    a.logabs[b] = (float)(logdet + jv);
    if (a.sign) a.sign[b] = (float)misc[1];
    if (DERIV) {
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
      if (a.e_loc) a.e_loc[b] = (float)e;
      if (a.lap) a.lap[b] = (float)lap;
      if (a.pot_out) a.pot_out[b] = (float)v;
      // the reference keeps every walker with a finite log|psi| and E_L (train.py:79-90)
      if (!(st & (PSIF_ST_NONFINITE_LOGDET | PSIF_ST_NONFINITE_ELOC))) { e_ok = e; e2_ok = e * e; n_ok = 1.0; }
    }
    if (a.status) a.status[b] = st;
  }
  if (DERIV && a.accum != nullptr) {
    // block reduction of the energy statistics (threads with rem != 0 contribute zeros)
    __shared__ double red[3][4];
    e_ok = warp_sum(e_ok); e2_ok = warp_sum(e2_ok); n_ok = warp_sum(n_ok);
    const int warp = tid >> 5, lane = tid & 31;
    if (lane == 0) { red[0][warp] = e_ok; red[1][warp] = e2_ok; red[2][warp] = n_ok; }
    __syncthreads();
    if (tid == 0) {
      const int nw = (blockDim.x + 31) >> 5;
      double s0 = 0, s1 = 0, s2 = 0;
      for (int w2 = 0; w2 < nw; ++w2) { s0 += red[0][w2]; s1 += red[1][w2]; s2 += red[2][w2]; }
      if (s2 > 0.0) { atomicAdd(a.accum + 0, s0); atomicAdd(a.accum + 1, s1); atomicAdd(a.accum + 2, s2); }
    }
  }
}

template <int NM, bool DERIV>
inline int32_t det_launch_t(DetArgs& a, cudaStream_t st) {
  const int T = DERIV ? a.C - 2 : 0;
  a.tpw = 2 * a.K;
  a.wpb = a.tpw >= 128 ? 1 : 128 / a.tpw;
  const int threads = ((a.tpw * a.wpb + 31) / 32) * 32;
  constexpr bool XSMEM = DERIV && (NM > 5);
  size_t smem = (size_t)a.wpb * det_smem_doubles_per_walker(a.K, T) * sizeof(double);
  if (XSMEM) smem += (size_t)NM * NM * threads * sizeof(double);
  if (smem > 220 * 1024) return fail(PSIF_E_INVALID, "slogdet: shared memory budget exceeded%s");
  size_t& configured = dev_smem_cfg().det[NM][DERIV ? 1 : 0];
  if (smem > 48 * 1024 && smem > configured) {
    PSIF_CUDA_CHECK(cudaFuncSetAttribute(det_combine_kernel<NM, DERIV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  const unsigned grid = (unsigned)cdiv(a.B, a.wpb);
  auto kern = det_combine_kernel<NM, DERIV>;
  PSIF_LAUNCH(kern, grid, threads, smem, st, a);
  return PSIF_OK;
}

inline int32_t det_launch(DetArgs& a, bool deriv, cudaStream_t st) {
  if (a.B <= 0) return PSIF_OK;
  const int nm = a.n[0] > a.n[1] ? a.n[0] : a.n[1];
  if (nm > PSIF_MAX_SPIN || a.K > PSIF_MAX_DET || a.K < 1) return fail(PSIF_E_INVALID, "slogdet: n_spin > 8 or n_det > 64 unsupported%s");
#define PSIF_DET_CASE(NMV)                                          \
  case NMV:                                                         \
    return deriv ? det_launch_t<NMV, true>(a, st) : det_launch_t<NMV, false>(a, st);
  switch (nm <= 1 ? 1 : nm) {
    PSIF_DET_CASE(1) PSIF_DET_CASE(2) PSIF_DET_CASE(3) PSIF_DET_CASE(4)
    PSIF_DET_CASE(5) PSIF_DET_CASE(6) PSIF_DET_CASE(7) PSIF_DET_CASE(8)
  }
#undef PSIF_DET_CASE
  return fail(PSIF_E_INVALID, "slogdet: bad size%s");
}

// ------------------------------------------------------------------------------------------
// Jastrow (jastrow.py:67-87) with closed-form gradient/Laplacian, and the softened Coulomb
// potential (hamiltonian.py:15-35 + nuclear repulsion).  One thread per walker, fp64.
// ------------------------------------------------------------------------------------------
struct NucleiD {
  int natom;
  double R[PSIF_MAX_ATOMS][3];
  double Z[PSIF_MAX_ATOMS];
  double vnn;
};

__global__ void __launch_bounds__(128)
jastrow_potential_kernel(const float* __restrict__ x, long long B, int N, int n_up, double a_par, double a_anti,
                         const float* __restrict__ alpha_dev /* [anti, par] or null */, NucleiD nuc, int deriv, int want_pot, double* __restrict__ jval,
                         double* __restrict__ jgrad, double* __restrict__ jlap, double* __restrict__ pot) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  if (alpha_dev != nullptr) { a_anti = (double)alpha_dev[0]; a_par = (double)alpha_dev[1]; }
  double px[PSIF_MAX_ELEC], py[PSIF_MAX_ELEC], pz[PSIF_MAX_ELEC];
  double gx[PSIF_MAX_ELEC], gy[PSIF_MAX_ELEC], gz[PSIF_MAX_ELEC];
#pragma unroll
  for (int i = 0; i < PSIF_MAX_ELEC; ++i) {
    if (i < N) {
      px[i] = x[(b * N + i) * 3 + 0]; py[i] = x[(b * N + i) * 3 + 1]; pz[i] = x[(b * N + i) * 3 + 2];
    }
    gx[i] = gy[i] = gz[i] = 0.0;
  }
  double val = 0.0, lap = 0.0, vee = 0.0, ven = 0.0;
#pragma unroll
  for (int i = 0; i < PSIF_MAX_ELEC; ++i) {
    if (i >= N) continue;
    if (want_pot) {
      for (int at = 0; at < nuc.natom; ++at) {
        const double dx = px[i] - nuc.R[at][0], dy = py[i] - nuc.R[at][1], dz = pz[i] - nuc.R[at][2];
        ven -= nuc.Z[at] / (sqrt(dx * dx + dy * dy + dz * dz + kCoulombEps) + kCoulombEps);
      }
    }
#pragma unroll
    for (int j = i + 1; j < PSIF_MAX_ELEC; ++j) {
      if (j >= N) continue;
      const double dx = px[i] - px[j], dy = py[i] - py[j], dz = pz[i] - pz[j];
      const double d2 = dx * dx + dy * dy + dz * dz;
      if (want_pot) vee += 1.0 / (sqrt(d2 + kCoulombEps) + kCoulombEps);
      const bool same = (i < n_up) == (j < n_up);
      const double c = same ? -0.25 : -0.5;
      const double al = same ? a_par : a_anti;
      const double rt = sqrt(d2 + kJastrowEps);
      const double den = al + rt;
      val += c * al * al / den;
      if (deriv) {
        const double f1 = -c * al * al / (den * den);
        const double f2 = 2.0 * c * al * al / (den * den * den);
        const double s = f1 / rt;
        gx[i] += s * dx; gy[i] += s * dy; gz[i] += s * dz;
        gx[j] -= s * dx; gy[j] -= s * dy; gz[j] -= s * dz;
        lap += 2.0 * (f2 * d2 / (rt * rt) + f1 * (3.0 / rt - d2 / (rt * rt * rt)));
      }
    }
  }
  if (jval) jval[b] = val;
  if (want_pot && pot) pot[b] = ven + vee + nuc.vnn;
  if (deriv) {
    jlap[b] = lap;
#pragma unroll
    for (int i = 0; i < PSIF_MAX_ELEC; ++i)
      if (i < N) {
        jgrad[(b * N + i) * 3 + 0] = gx[i];
        jgrad[(b * N + i) * 3 + 1] = gy[i];
        jgrad[(b * N + i) * 3 + 2] = gz[i];
      }
  }
}

// The same, one LANE per electron (half a warp per walker, N <= 16): electron i sums its N - 1 pair terms -- value and
// potential count every pair from both ends (hence the halves), the gradient is electron i's own, the Laplacian's pair
// term 2 (f'' ...) is one (f'' ...) per ordered pair -- and the half-warp adds up in fp64.  The thread-per-walker kernel
// above walks through all N (N - 1) / 2 pairs of a walker serially in fp64: 16 - 32 CTAs and 27 - 65 us per launch
// whatever the batch, 5 % of a Metropolis step.
__global__ void __launch_bounds__(128)
jastrow_potential_lane_kernel(const float* __restrict__ x, long long B, int N, int n_up, double a_par, double a_anti,
                              const float* __restrict__ alpha_dev /* [anti, par] or null */, NucleiD nuc, int deriv, int want_pot,
                              double* __restrict__ jval, double* __restrict__ jgrad, double* __restrict__ jlap, double* __restrict__ pot) {
  const int i = threadIdx.x & 15;
  const long long b = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
  const bool live = b < B && i < N;
  if (alpha_dev != nullptr) { a_anti = (double)alpha_dev[0]; a_par = (double)alpha_dev[1]; }
  const long long bb = b < B ? b : B - 1;                       // idle half-warps shadow the last walker (no stores)
  const int ii = i < N ? i : 0;
  const double xi = x[(bb * N + ii) * 3 + 0], yi = x[(bb * N + ii) * 3 + 1], zi = x[(bb * N + ii) * 3 + 2];
  double val = 0.0, lap = 0.0, vee = 0.0, ven = 0.0, gx = 0.0, gy = 0.0, gz = 0.0;
  if (live) {
    if (want_pot)
      for (int at = 0; at < nuc.natom; ++at) {
        const double dx = xi - nuc.R[at][0], dy = yi - nuc.R[at][1], dz = zi - nuc.R[at][2];
        ven -= nuc.Z[at] / (sqrt(dx * dx + dy * dy + dz * dz + kCoulombEps) + kCoulombEps);
      }
    for (int j = 0; j < N; ++j) {
      if (j == i) continue;
      const double dx = xi - (double)x[(bb * N + j) * 3 + 0], dy = yi - (double)x[(bb * N + j) * 3 + 1],
                   dz = zi - (double)x[(bb * N + j) * 3 + 2];
      const double d2 = dx * dx + dy * dy + dz * dz;
      if (want_pot) vee += 0.5 / (sqrt(d2 + kCoulombEps) + kCoulombEps);
      const bool same = (i < n_up) == (j < n_up);
      const double c = same ? -0.25 : -0.5;
      const double al = same ? a_par : a_anti;
      const double rt = sqrt(d2 + kJastrowEps);
      const double den = al + rt;
      val += 0.5 * c * al * al / den;
      if (deriv) {
        const double f1 = -c * al * al / (den * den);
        const double f2 = 2.0 * c * al * al / (den * den * den);
        const double sgr = f1 / rt;
        gx += sgr * dx; gy += sgr * dy; gz += sgr * dz;
        lap += f2 * d2 / (rt * rt) + f1 * (3.0 / rt - d2 / (rt * rt * rt));
      }
    }
  }
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) {
    val += __shfl_xor_sync(0xffffffffu, val, o);
    lap += __shfl_xor_sync(0xffffffffu, lap, o);
    vee += __shfl_xor_sync(0xffffffffu, vee, o);
    ven += __shfl_xor_sync(0xffffffffu, ven, o);
  }
  if (b < B && i == 0) {
    if (jval) jval[b] = val;
    if (want_pot && pot) pot[b] = ven + vee + nuc.vnn;
    if (deriv) jlap[b] = lap;
  }
  if (live && deriv) {
    jgrad[(b * N + i) * 3 + 0] = gx;
    jgrad[(b * N + i) * 3 + 1] = gy;
    jgrad[(b * N + i) * 3 + 2] = gz;
  }
}

inline int32_t jastrow_potential_launch(const float* x, long long B, int N, int n_up, double a_par, double a_anti,
                                        const float* alpha_dev, const NucleiD& nuc, int deriv, int want_pot, double* jval,
                                        double* jgrad, double* jlap, double* pot, cudaStream_t st) {
  if (B <= 0) return PSIF_OK;
  if (N <= 16) {
    PSIF_LAUNCH(jastrow_potential_lane_kernel, (unsigned)cdiv(B * 16, 128), 128, 0, st, x, B, N, n_up, a_par, a_anti, alpha_dev,
                nuc, deriv, want_pot, jval, jgrad, jlap, pot);
  } else {
    PSIF_LAUNCH(jastrow_potential_kernel, (unsigned)cdiv(B, 128), 128, 0, st, x, B, N, n_up, a_par, a_anti, alpha_dev, nuc,
                deriv, want_pot, jval, jgrad, jlap, pot);
  }
  return PSIF_OK;
}

}  // namespace psif
