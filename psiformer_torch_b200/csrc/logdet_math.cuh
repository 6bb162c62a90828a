// Per-block mathematics of the standalone logdet_matmul derivatives (logdet_matmul.py:35-70, 94-120), runtime size
// n <= 8, fp64, one thread per block.  Host/device so that tests/ can exercise exactly this code on the CPU.
//
//   f(A) = sum_i log max(s_i(A), 1e-6)           s_i = singular values        (logdet_matmul.py:45-55)
//   G    = df/dA    = sum_{i in U} u_i v_i^T / s_i                             U = {i : s_i >= 1e-6}, Cl = the rest
//   H[E] = d<G,E>/dA = U M V^T,  P = U^T E V,
//            M_ab = -P_ba / (s_a s_b)                                   a, b in U
//            M_ci = (s_i P_ci + s_c P_ic) / (s_i (s_i^2 - s_c^2))       c in Cl, i in U      (second-order perturbation
//            M_ic = (s_i P_ic + s_c P_ci) / (s_i (s_i^2 - s_c^2))                              of singular values)
//            M_cc' = 0
// With no clamped singular value this is G = A^-T, H[E] = -(A^-1 E A^-1)^T, which is what the fast path uses (pivoted
// Gauss-Jordan); a block takes the SVD path only when ||A^-1||_F > 1e6 (then sigma_min may be below the clamp).
#pragma once
#include <math.h>

#ifndef PSIF_HD
#ifdef __CUDACC__
#define PSIF_HD __host__ __device__ __forceinline__
#else
#define PSIF_HD inline
#endif
#endif

namespace psif {

constexpr int LD_MAXN = 8, LD_MAX_DET = 64;
constexpr double LD_MIN_SINGULAR = 1e-6;   // logdet_matmul.py:16
constexpr double LD_OUTPUT_FLOOR = 1e-12;  // logdet_matmul.py:17
constexpr double LD_DET_JITTER = 1e-4;     // logdet_matmul.py:18

struct LdBlock {
  int n;
  bool svd;                         // SVD path (some singular value may be clamped)
  double logdet, sign;              // f(A) and sign(det A)
  double inv[LD_MAXN * LD_MAXN];    // fast path: A^-1 (row major)
  double U[LD_MAXN * LD_MAXN], V[LD_MAXN * LD_MAXN], s[LD_MAXN];   // SVD path: A = U diag(s) V^T (columns)
};

// in-place inverse of a (row-major n x n) by Gauss-Jordan with partial pivoting; returns log|det|, sign
PSIF_HD void ld_gj_inverse(double* a, int n, double& logdet, double& sign) {
  int piv[LD_MAXN];
  logdet = 0.0;
  sign = 1.0;
  for (int k = 0; k < n; ++k) {
    int p = k;
    double best = fabs(a[k * n + k]);
    for (int r = k + 1; r < n; ++r) {
      const double v = fabs(a[r * n + k]);
      if (v > best) { best = v; p = r; }
    }
    piv[k] = p;
    if (p != k) {
      for (int j = 0; j < n; ++j) { const double t = a[k * n + j]; a[k * n + j] = a[p * n + j]; a[p * n + j] = t; }
      sign = -sign;
    }
    const double pv = a[k * n + k];
    logdet += log(fabs(pv));
    if (pv < 0.0) sign = -sign;
    if (pv == 0.0) sign = 0.0;
    const double inv = 1.0 / pv;
    a[k * n + k] = 1.0;
    for (int j = 0; j < n; ++j) a[k * n + j] *= inv;
    for (int i = 0; i < n; ++i) {
      if (i == k) continue;
      const double f = a[i * n + k];
      a[i * n + k] = 0.0;
      for (int j = 0; j < n; ++j) a[i * n + j] -= f * a[k * n + j];
    }
  }
  for (int k = n - 1; k >= 0; --k) {
    const int c = piv[k];
    if (c != k)
      for (int i = 0; i < n; ++i) { const double t = a[i * n + k]; a[i * n + k] = a[i * n + c]; a[i * n + c] = t; }
  }
}

// one-sided Jacobi: on return the columns of w are u_j s_j, V holds the right singular vectors (columns)
PSIF_HD void ld_jacobi(double* w, double* V, int n) {
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) V[i * n + j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 40; ++sweep) {
    double off = 0.0;
    for (int p = 0; p < n - 1; ++p)
      for (int q = p + 1; q < n; ++q) {
        double alpha = 0.0, beta = 0.0, gamma = 0.0;
        for (int i = 0; i < n; ++i) {
          alpha += w[i * n + p] * w[i * n + p];
          beta += w[i * n + q] * w[i * n + q];
          gamma += w[i * n + p] * w[i * n + q];
        }
        if (gamma == 0.0) continue;
        const double lim = fabs(gamma) / sqrt(alpha * beta + 1e-300);
        if (lim > off) off = lim;
        const double zeta = (beta - alpha) / (2.0 * gamma);
        const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / sqrt(1.0 + t * t), sn = c * t;
        for (int i = 0; i < n; ++i) {
          const double wp = w[i * n + p], wq = w[i * n + q];
          w[i * n + p] = c * wp - sn * wq;
          w[i * n + q] = sn * wp + c * wq;
          const double vp = V[i * n + p], vq = V[i * n + q];
          V[i * n + p] = c * vp - sn * vq;
          V[i * n + q] = sn * vp + c * vq;
        }
      }
    if (off < 1e-15) break;
  }
}

// factor one block: A = x + 1e-4 I already formed by the caller in `a` (row major, n x n; destroyed)
PSIF_HD void ld_factor(const double* a, int n, LdBlock& blk) {
  blk.n = n;
  blk.svd = false;
  for (int e = 0; e < n * n; ++e) blk.inv[e] = a[e];
  ld_gj_inverse(blk.inv, n, blk.logdet, blk.sign);
  double fro2 = 0.0;
  for (int e = 0; e < n * n; ++e) fro2 += blk.inv[e] * blk.inv[e];
  if (fro2 <= 1e12) return;                 // sigma_min >= 1 / ||A^-1||_F >= 1e-6: the clamp is provably inactive
  double w[LD_MAXN * LD_MAXN];
  for (int e = 0; e < n * n; ++e) w[e] = a[e];
  ld_jacobi(w, blk.V, n);
  bool clamped = false;
  double ldc = 0.0;
  for (int j = 0; j < n; ++j) {
    double n2 = 0.0;
    for (int i = 0; i < n; ++i) n2 += w[i * n + j] * w[i * n + j];
    blk.s[j] = sqrt(n2);
    if (blk.s[j] < LD_MIN_SINGULAR) clamped = true;
    ldc += log(fmax(blk.s[j], LD_MIN_SINGULAR));
  }
  if (!clamped) return;
  blk.svd = true;
  blk.logdet = ldc;
  if (blk.sign == 0.0) blk.sign = 1.0;      // torch.sign(det(u) det(v)) of an exactly singular block is +-1; take +1
  // left singular vectors; columns with a clamped (tiny) s_j are re-orthogonalised against all others (modified
  // Gram-Schmidt), starting from A v_j or, if that vanished, from the unit vector least represented so far
  for (int j = 0; j < n; ++j)
    if (blk.s[j] >= LD_MIN_SINGULAR)
      for (int i = 0; i < n; ++i) blk.U[i * n + j] = w[i * n + j] / blk.s[j];
  for (int j = 0; j < n; ++j) {
    if (blk.s[j] >= LD_MIN_SINGULAR) continue;
    double u[LD_MAXN];
    for (int attempt = 0; attempt < 1 + LD_MAXN; ++attempt) {
      if (attempt == 0) { for (int i = 0; i < n; ++i) u[i] = w[i * n + j]; }
      else { for (int i = 0; i < n; ++i) u[i] = (i == attempt - 1) ? 1.0 : 0.0; }
      double nrm0 = 0.0;
      for (int i = 0; i < n; ++i) nrm0 += u[i] * u[i];
      if (nrm0 == 0.0) continue;
      for (int i = 0; i < n; ++i) u[i] /= sqrt(nrm0);
      for (int pass = 0; pass < 2; ++pass)
        for (int o = 0; o < n; ++o) {
          const bool done = (blk.s[o] >= LD_MIN_SINGULAR) || (o < j);   // columns already fixed
          if (!done || o == j) continue;
          double d = 0.0;
          for (int i = 0; i < n; ++i) d += u[i] * blk.U[i * n + o];
          for (int i = 0; i < n; ++i) u[i] -= d * blk.U[i * n + o];
        }
      double nrm = 0.0;
      for (int i = 0; i < n; ++i) nrm += u[i] * u[i];
      if (nrm > 1e-6) {
        for (int i = 0; i < n; ++i) blk.U[i * n + j] = u[i] / sqrt(nrm);
        break;
      }
    }
  }
}

// G = df/dA (row major n x n)
PSIF_HD void ld_grad(const LdBlock& blk, double* G) {
  const int n = blk.n;
  if (!blk.svd) {
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) G[i * n + j] = blk.inv[j * n + i];
    return;
  }
  for (int a = 0; a < n; ++a)
    for (int b = 0; b < n; ++b) {
      double g = 0.0;
      for (int i = 0; i < n; ++i)
        if (blk.s[i] >= LD_MIN_SINGULAR) g += blk.U[a * n + i] * blk.V[b * n + i] / blk.s[i];
      G[a * n + b] = g;
    }
}

// <G, E>
PSIF_HD double ld_grad_dot(const LdBlock& blk, const double* E) {
  const int n = blk.n;
  double acc = 0.0;
  if (!blk.svd) {
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) acc += blk.inv[j * n + i] * E[i * n + j];
    return acc;
  }
  double G[LD_MAXN * LD_MAXN];
  ld_grad(blk, G);
  for (int e = 0; e < n * n; ++e) acc += G[e] * E[e];
  return acc;
}

// H[E] = d<G, E>/dA (row major n x n)
PSIF_HD void ld_hess_apply(const LdBlock& blk, const double* E, double* H) {
  const int n = blk.n;
  double T[LD_MAXN * LD_MAXN];
  if (!blk.svd) {
    // -(A^-1 E A^-1)^T
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) {
        double t = 0.0;
        for (int k = 0; k < n; ++k) t += blk.inv[i * n + k] * E[k * n + j];
        T[i * n + j] = t;
      }
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) {
        double t = 0.0;
        for (int k = 0; k < n; ++k) t += T[i * n + k] * blk.inv[k * n + j];
        H[j * n + i] = -t;
      }
    return;
  }
  double P[LD_MAXN * LD_MAXN], M[LD_MAXN * LD_MAXN];
  // P = U^T E V
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      double t = 0.0;
      for (int k = 0; k < n; ++k) t += E[i * n + k] * blk.V[k * n + j];
      T[i * n + j] = t;
    }
  for (int a = 0; a < n; ++a)
    for (int b = 0; b < n; ++b) {
      double t = 0.0;
      for (int k = 0; k < n; ++k) t += blk.U[k * n + a] * T[k * n + b];
      P[a * n + b] = t;
    }
  for (int a = 0; a < n; ++a)
    for (int b = 0; b < n; ++b) {
      const bool ua = blk.s[a] >= LD_MIN_SINGULAR, ub = blk.s[b] >= LD_MIN_SINGULAR;
      double m = 0.0;
      if (ua && ub) m = -P[b * n + a] / (blk.s[a] * blk.s[b]);
      else if (!ua && ub) {          // a = c clamped, b = i unclamped
        const double si = blk.s[b], sc = blk.s[a];
        m = (si * P[a * n + b] + sc * P[b * n + a]) / (si * (si * si - sc * sc));
      } else if (ua && !ub) {        // a = i unclamped, b = c clamped
        const double si = blk.s[a], sc = blk.s[b];
        m = (si * P[a * n + b] + sc * P[b * n + a]) / (si * (si * si - sc * sc));
      }
      M[a * n + b] = m;
    }
  // H = U M V^T
  for (int a = 0; a < n; ++a)
    for (int b = 0; b < n; ++b) {
      double t = 0.0;
      for (int k = 0; k < n; ++k) t += M[a * n + k] * blk.V[b * n + k];
      T[a * n + b] = t;
    }
  for (int a = 0; a < n; ++a)
    for (int b = 0; b < n; ++b) {
      double t = 0.0;
      for (int k = 0; k < n; ++k) t += blk.U[a * n + k] * T[k * n + b];
      H[a * n + b] = t;
    }
}

// ---- the whole op for one walker (all K determinants, both spins); see logdet_grad.cuh for the formulas ------------
struct LdGradArgs {
  const float *x1, *x2, *w, *gbar;      // [B][K][nu][nu], [B][K][nd][nd], [K], [B]
  const float *v1, *v2, *vw;            // second order only: cotangents of dx1, dx2 (same shapes) and dw [K]
  float *o1, *o2, *ow, *ogbar;          // outputs: [B][K][nu][nu], [B][K][nd][nd], [B][K] (per walker, caller sums), [B]
  long long B;
  int K, nu, nd;
};

PSIF_HD void ld_load_block(const float* p, int n, double* a) {
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) a[i * n + j] = (double)p[i * n + j] + (i == j ? LD_DET_JITTER : 0.0);
}

template <bool SECOND>
PSIF_HD void ld_walker(const LdGradArgs& a, long long b) {
  const int K = a.K;
  const int nn[2] = {a.nu * a.nu, a.nd * a.nd};
  const float* xs[2] = {a.x1 + b * K * nn[0], a.x2 + b * K * nn[1]};
  const float* vs[2] = {SECOND ? a.v1 + b * K * nn[0] : nullptr, SECOND ? a.v2 + b * K * nn[1] : nullptr};
  float* os[2] = {a.o1 + b * K * nn[0], a.o2 + b * K * nn[1]};
  const int ns[2] = {a.nu, a.nd};
  double ell[2][LD_MAX_DET], sgn[2][LD_MAX_DET], ab[LD_MAX_DET];   // ab = alpha_k + beta_k
  LdBlock blk;
  double A[LD_MAXN * LD_MAXN], E[LD_MAXN * LD_MAXN];
  for (int k = 0; k < K; ++k) {
    ab[k] = 0.0;
    for (int sp = 0; sp < 2; ++sp) {
      ld_load_block(xs[sp] + k * nn[sp], ns[sp], A);
      ld_factor(A, ns[sp], blk);
      ell[sp][k] = blk.logdet;
      sgn[sp][k] = blk.sign;
      if (SECOND) {
        for (int e = 0; e < nn[sp]; ++e) E[e] = (double)vs[sp][k * nn[sp] + e];
        ab[k] += ld_grad_dot(blk, E);
      }
    }
  }
  double m0 = -INFINITY, m1 = -INFINITY;
  for (int k = 0; k < K; ++k) { m0 = fmax(m0, ell[0][k]); m1 = fmax(m1, ell[1][k]); }
  double S = 0.0;
  for (int k = 0; k < K; ++k) {
    ell[0][k] = sgn[0][k] * sgn[1][k] * exp(ell[0][k] - m0 + ell[1][k] - m1);   // e_k
    S += (double)a.w[k] * ell[0][k];
  }
  const double invS = (fabs(S) < LD_OUTPUT_FLOOR || !(S == S)) ? 0.0 : 1.0 / S;
  const double gb = (double)a.gbar[b];
  double s = 0.0;
  if (SECOND) {
    for (int k = 0; k < K; ++k) s += invS * ell[0][k] * ((double)a.w[k] * ab[k] + (double)a.vw[k]);
    a.ogbar[b] = (float)s;
  }
  for (int k = 0; k < K; ++k) {
    const double ek = invS * ell[0][k], wk = (double)a.w[k], ck = wk * ek;
    if (!SECOND) {
      a.ow[b * K + k] = (float)(gb * ek);
    } else {
      a.ow[b * K + k] = (float)(gb * ek * (ab[k] - s));
    }
    // tau_k - s with the Vw_k / w_k term folded in as c_k (Vw_k / w_k) = e_k Vw_k (safe for w_k = 0)
    const double c_tau = SECOND ? ck * (ab[k] - s) + ek * (double)a.vw[k] : 0.0;
    for (int sp = 0; sp < 2; ++sp) {
      const int n = ns[sp];
      ld_load_block(xs[sp] + k * nn[sp], n, A);
      ld_factor(A, n, blk);
      double G[LD_MAXN * LD_MAXN];
      ld_grad(blk, G);
      float* o = os[sp] + k * nn[sp];
      if (!SECOND) {
        for (int e = 0; e < nn[sp]; ++e) o[e] = (float)(gb * ck * G[e]);
      } else {
        double H[LD_MAXN * LD_MAXN];
        for (int e = 0; e < nn[sp]; ++e) E[e] = (double)vs[sp][k * nn[sp] + e];
        ld_hess_apply(blk, E, H);
        for (int e = 0; e < nn[sp]; ++e) o[e] = (float)(gb * (c_tau * G[e] + ck * H[e]));
      }
    }
  }
}

}  // namespace psif
