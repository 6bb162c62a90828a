// First and second derivatives of the standalone logdet_matmul op (logdet_matmul.py:94-120: the reference's backward
// re-runs the value function under autograd with create_graph, so the op is twice differentiable).
//
//   f(x1, x2, w) = log|S|,  S = sum_k w_k e_k,  e_k = sgn_k exp(f1_k - m1 + f2_k - m2),  f1_k = f(x1_k + 1e-4 I) of
//   logdet_math.cuh (singular values clamped at 1e-6), c_k = w_k e_k / S.
//   backward   : dx1_k = gbar c_k G1_k,  dx2_k = gbar c_k G2_k,  dw_k = sum_b gbar e_k / S
//   double bwd : with cotangents (V1, V2, Vw) of (dx1, dx2, dw):  alpha_k = <G1_k, V1_k>, beta_k = <G2_k, V2_k>,
//                s = sum_k c_k (alpha_k + beta_k) + (e_k / S) Vw_k                      -> d/d gbar
//                d s / d x1_j = c_j (tau_j - s) G1_j + c_j H1_j[V1_j],  tau_j = alpha_j + beta_j + Vw_j / w_j
//                d s / d w_j  = (e_j / S)(alpha_j + beta_j - s)
//   Below the 1e-12 floor of logdet_matmul.py:68 the clamp has zero gradient: everything is zero there.
// One thread per walker, fp64, blocks factored twice (once for the weights c_k, once for the outputs): this op is the
// API-parity path, the fused pipeline (slogdet.cuh) is the fast one.
#pragma once
#include "common.cuh"
#include "logdet_math.cuh"

namespace psif {

template <bool SECOND>
__global__ void __launch_bounds__(64)
logdet_grad_kernel(LdGradArgs a) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b < a.B) ld_walker<SECOND>(a, b);
}

inline int32_t logdet_grad_launch(const LdGradArgs& a, bool second, cudaStream_t st) {
  if (a.B <= 0) return PSIF_OK;
  if (a.K < 1 || a.K > LD_MAX_DET || a.nu < 1 || a.nd < 1 || a.nu > LD_MAXN || a.nd > LD_MAXN)
    return fail(PSIF_E_INVALID, "logdet_matmul derivatives: need 1 <= n_det <= 64 and 1 <= n <= 8 per spin%s");
  const unsigned grid = (unsigned)cdiv(a.B, 64);
  if (second) PSIF_LAUNCH(logdet_grad_kernel<true>, grid, 64, 0, st, a);
  else PSIF_LAUNCH(logdet_grad_kernel<false>, grid, 64, 0, st, a);
  return PSIF_OK;
}

}  // namespace psif
