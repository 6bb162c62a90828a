// CPU harness for psiformer_torch_b200/csrc/logdet_math.cuh (test infrastructure): the per-block and per-walker
// mathematics of the logdet_matmul derivative kernels are host/device functions; this file exposes them to ctypes so that
// tests/test_logdet_math.py can check them against torch autograd through the reference's SVD formulation without a GPU.
#include "../../psiformer_torch_b200/csrc/logdet_math.cuh"
#include "../../psiformer_torch_b200/csrc/smallmat.cuh"

using namespace psif;

// the compile-time sized routines of smallmat.cuh (registers on the device): what det_combine_kernel and the clamp fix-up use
template <int NM>
static void sm_svd_t(const double* A, double* W, double* V) {
  double w[NM * NM], v[NM * NM];
  for (int e = 0; e < NM * NM; ++e) w[e] = A[e];
  jacobi_svd<NM>(w, v);
  for (int e = 0; e < NM * NM; ++e) { W[e] = w[e]; V[e] = v[e]; }
}
template <int NM>
static void sm_inv_t(const double* A, double* X, double* logdet, double* sign) {
  double x[NM * NM], mp;
  for (int e = 0; e < NM * NM; ++e) x[e] = A[e];
  gj_inverse<NM>(x, *logdet, *sign, mp);
  for (int e = 0; e < NM * NM; ++e) X[e] = x[e];
}

extern "C" {

#define SM_CASES(F, ...) switch (n) { case 1: F<1>(__VA_ARGS__); break; case 2: F<2>(__VA_ARGS__); break; case 3: F<3>(__VA_ARGS__); break; \
  case 4: F<4>(__VA_ARGS__); break; case 5: F<5>(__VA_ARGS__); break; case 6: F<6>(__VA_ARGS__); break; case 7: F<7>(__VA_ARGS__); break; \
  case 8: F<8>(__VA_ARGS__); break; default: return -1; } return 0;
// A = (columns u_j s_j of W) V^T
int sm_host_jacobi_svd(const double* A, int n, double* W, double* V) { SM_CASES(sm_svd_t, A, W, V) }
int sm_host_gj_inverse(const double* A, int n, double* X, double* logdet, double* sign) { SM_CASES(sm_inv_t, A, X, logdet, sign) }

// one block: A (already jittered), direction E -> f, sign, G, H[E], whether the SVD path was taken
void ld_host_block(const double* A, int n, const double* E, double* f, double* sign, double* G, double* H, int* svd) {
  LdBlock blk;
  ld_factor(A, n, blk);
  *f = blk.logdet;
  *sign = blk.sign;
  *svd = blk.svd ? 1 : 0;
  ld_grad(blk, G);
  ld_hess_apply(blk, E, H);
}

// <G, E> and <H[E], E> of a clamped block without forming H: the algebra of det_clamp_fixup_kernel (slogdet_clamp.cuh),
// P = U^T E V, <G, E> = sum_{i unclamped} P_ii / s_i, <H[E], E> = <M(P), P>; returns 0 if the block is not on the SVD path
int ld_host_clamped_forms(const double* A, int n, const double* E, double* gdot, double* quad) {
  LdBlock blk;
  ld_factor(A, n, blk);
  if (!blk.svd) return 0;
  double M[LD_MAXN * LD_MAXN], P[LD_MAXN * LD_MAXN];
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      double t = 0.0;
      for (int k = 0; k < n; ++k) t += E[i * n + k] * blk.V[k * n + j];
      M[i * n + j] = t;
    }
  for (int a = 0; a < n; ++a)
    for (int b = 0; b < n; ++b) {
      double t = 0.0;
      for (int k = 0; k < n; ++k) t += blk.U[k * n + a] * M[k * n + b];
      P[a * n + b] = t;
    }
  double g = 0.0, q = 0.0;
  for (int p = 0; p < n; ++p) {
    const double sa = blk.s[p];
    const bool ua = sa >= LD_MIN_SINGULAR;
    for (int r = p; r < n; ++r) {
      const double pab = P[p * n + r], pba = P[r * n + p];
      if (r == p) {
        if (ua) { g += pab / sa; q -= pab * pab / (sa * sa); }
      } else {
        const double sb = blk.s[r];
        const bool ub = sb >= LD_MIN_SINGULAR;
        if (ua && ub) q -= 2.0 * pab * pba / (sa * sb);
        else if (ua != ub) {
          const double si = ua ? sa : sb, sc = ua ? sb : sa;
          const double den = 1.0 / (si * (si * si - sc * sc));
          q += ((si * pab + sc * pba) * pab + (si * pba + sc * pab) * pba) * den;
        }
      }
    }
  }
  *gdot = g;
  *quad = q;
  return 1;
}

void ld_host_grad(const float* x1, const float* x2, const float* w, const float* gbar, long long B, int K, int nu, int nd,
                  float* dx1, float* dx2, float* dw_per_walker) {
  LdGradArgs a{};
  a.x1 = x1; a.x2 = x2; a.w = w; a.gbar = gbar; a.o1 = dx1; a.o2 = dx2; a.ow = dw_per_walker;
  a.B = B; a.K = K; a.nu = nu; a.nd = nd;
  for (long long b = 0; b < B; ++b) ld_walker<false>(a, b);
}

void ld_host_grad_grad(const float* x1, const float* x2, const float* w, const float* gbar, const float* v1, const float* v2,
                       const float* vw, long long B, int K, int nu, int nd, float* d_gbar, float* d1, float* d2,
                       float* dw_per_walker) {
  LdGradArgs a{};
  a.x1 = x1; a.x2 = x2; a.w = w; a.gbar = gbar; a.v1 = v1; a.v2 = v2; a.vw = vw;
  a.o1 = d1; a.o2 = d2; a.ow = dw_per_walker; a.ogbar = d_gbar;
  a.B = B; a.K = K; a.nu = nu; a.nd = nd;
  for (long long b = 0; b < B; ++b) ld_walker<true>(a, b);
}

}  // extern "C"
