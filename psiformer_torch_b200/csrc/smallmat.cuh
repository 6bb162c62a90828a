// Small dense matrices held by ONE thread: in-place Gauss-Jordan inverse with partial pivoting,
// log|det| and sign, written with compile-time indices only so everything stays in registers.
// fp64: the orbital blocks are ill-conditioned and this stage is O(K n^3) per
// walker, i.e. negligible next to the GEMMs.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define PSIF_HD __host__ __device__ __forceinline__
#else
#define PSIF_HD inline
#endif

namespace psif {

// a: row-major NM x NM, overwritten with its inverse.  Returns log|det a| and sign(det a) in {-1,0,1}.
// A singular matrix gives logdet = -inf, sign = 0 and a non-finite inverse.
template <int NM>
PSIF_HD void gj_inverse(double (&a)[NM * NM], double& logdet, double& sign, double& min_pivot) {
  int piv[NM];
  logdet = 0.0;
  sign = 1.0;
  min_pivot = INFINITY;
#pragma unroll
  for (int k = 0; k < NM; ++k) {
    // pivot search in column k, rows k..NM-1 (first maximum wins)
    int p = k;
    double best = fabs(a[k * NM + k]);
#pragma unroll
    for (int r = k + 1; r < NM; ++r) {
      const double v = fabs(a[r * NM + k]);
      if (v > best) { best = v; p = r; }
    }
    piv[k] = p;
#pragma unroll
    for (int r = k + 1; r < NM; ++r) {
      if (p == r) {
#pragma unroll
        for (int j = 0; j < NM; ++j) {
          const double t = a[k * NM + j];
          a[k * NM + j] = a[r * NM + j];
          a[r * NM + j] = t;
        }
        sign = -sign;
      }
    }
    const double pv = a[k * NM + k];
    if (best < min_pivot) min_pivot = best;
    logdet += log(fabs(pv));
    if (pv < 0.0) sign = -sign;
    if (pv == 0.0) sign = 0.0;
    const double inv = 1.0 / pv;
    a[k * NM + k] = 1.0;
#pragma unroll
    for (int j = 0; j < NM; ++j) a[k * NM + j] *= inv;
#pragma unroll
    for (int i = 0; i < NM; ++i) {
      if (i != k) {
        const double f = a[i * NM + k];
        a[i * NM + k] = 0.0;
#pragma unroll
        for (int j = 0; j < NM; ++j) a[i * NM + j] -= f * a[k * NM + j];
      }
    }
  }
  // undo the row interchanges as column interchanges, in reverse order
#pragma unroll
  for (int k = NM - 1; k >= 0; --k) {
#pragma unroll
    for (int c = k + 1; c < NM; ++c) {
      if (piv[k] == c) {
#pragma unroll
        for (int i = 0; i < NM; ++i) {
          const double t = a[i * NM + k];
          a[i * NM + k] = a[i * NM + c];
          a[i * NM + c] = t;
        }
      }
    }
  }
}

// Singular values of a (row-major NM x NM, destroyed) by one-sided Jacobi; s[] unsorted.
// Used only on the rare blocks whose smallest singular value may fall under the 1e-6 clamp of
// logdet_matmul.py:50-51, to reproduce sum_i log(max(s_i, 1e-6)) exactly.
template <int NM>
PSIF_HD void jacobi_singular_values(double (&a)[NM * NM], double (&s)[NM]) {
  for (int sweep = 0; sweep < 30; ++sweep) {
    double off = 0.0;
#pragma unroll
    for (int p = 0; p < NM - 1; ++p) {
#pragma unroll
      for (int q = p + 1; q < NM; ++q) {
        double alpha = 0.0, beta = 0.0, gamma = 0.0;
#pragma unroll
        for (int i = 0; i < NM; ++i) {
          alpha += a[i * NM + p] * a[i * NM + p];
          beta += a[i * NM + q] * a[i * NM + q];
          gamma += a[i * NM + p] * a[i * NM + q];
        }
        if (gamma != 0.0) {
          const double lim = fabs(gamma) / sqrt(alpha * beta + 1e-300);
          if (lim > off) off = lim;
          const double zeta = (beta - alpha) / (2.0 * gamma);
          const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
          const double c = 1.0 / sqrt(1.0 + t * t), sn = c * t;
#pragma unroll
          for (int i = 0; i < NM; ++i) {
            const double ap = a[i * NM + p], aq = a[i * NM + q];
            a[i * NM + p] = c * ap - sn * aq;
            a[i * NM + q] = sn * ap + c * aq;
          }
        }
      }
    }
    if (off < 1e-15) break;
  }
#pragma unroll
  for (int j = 0; j < NM; ++j) {
    double n2 = 0.0;
#pragma unroll
    for (int i = 0; i < NM; ++i) n2 += a[i * NM + j] * a[i * NM + j];
    s[j] = sqrt(n2);
  }
}

// One-sided Jacobi with the right singular vectors accumulated: on return the columns of w are u_j s_j and V holds the
// v_j (columns); the same sweep order and stopping rule as ld_jacobi (logdet_math.cuh), compile-time sized.
template <int NM>
PSIF_HD void jacobi_svd(double (&w)[NM * NM], double (&V)[NM * NM]) {
#pragma unroll
  for (int i = 0; i < NM; ++i)
#pragma unroll
    for (int j = 0; j < NM; ++j) V[i * NM + j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 40; ++sweep) {
    double off = 0.0;
#pragma unroll
    for (int p = 0; p < NM - 1; ++p) {
#pragma unroll
      for (int q = p + 1; q < NM; ++q) {
        double alpha = 0.0, beta = 0.0, gamma = 0.0;
#pragma unroll
        for (int i = 0; i < NM; ++i) {
          alpha += w[i * NM + p] * w[i * NM + p];
          beta += w[i * NM + q] * w[i * NM + q];
          gamma += w[i * NM + p] * w[i * NM + q];
        }
        if (gamma != 0.0) {
          const double lim = fabs(gamma) / sqrt(alpha * beta + 1e-300);
          if (lim > off) off = lim;
          const double zeta = (beta - alpha) / (2.0 * gamma);
          const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
          const double c = 1.0 / sqrt(1.0 + t * t), sn = c * t;
#pragma unroll
          for (int i = 0; i < NM; ++i) {
            const double wp = w[i * NM + p], wq = w[i * NM + q];
            w[i * NM + p] = c * wp - sn * wq;
            w[i * NM + q] = sn * wp + c * wq;
            const double vp = V[i * NM + p], vq = V[i * NM + q];
            V[i * NM + p] = c * vp - sn * vq;
            V[i * NM + q] = sn * vp + c * vq;
          }
        }
      }
    }
    if (off < 1e-15) break;
  }
}

}  // namespace psif
