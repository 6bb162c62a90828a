// Derivatives of log|psi| on walkers where the reference's singular-value clamp is ACTIVE (logdet_matmul.py:50-51).
//
// det_combine_kernel (slogdet.cuh) differentiates log|det A| = -log|det A^-1| through a Gauss-Jordan inverse, which is
// what torch.linalg.svd -> clamp(s, 1e-6) -> log reduces to as long as no singular value is below the clamp.  A block
// with sigma_min < 1e-6 gets its VALUE from the clamped singular values there, but its tangent / Laplacian terms would
// still be those of the unclamped function; the kernel raises PSIF_ST_CLAMP_SUSPECT on such walkers (2 % of raw N(0, I)
// walkers of N2, none of the pinned fixtures).  In the reference the clamped directions carry no gradient
// (torch.clamp) and the second derivative picks up the second-order perturbation of the remaining singular values.
//
// det_clamp_fixup_kernel runs behind det_combine_kernel in energy mode: CTAs of 256 threads look at one or a few walkers
// each and skip the unflagged ones.  A flagged walker is recomputed from the closed forms of logdet_math.cuh (the ones
// behind the twice-differentiable LogDetMatmul op):
//   phase A  one thread per block: Gauss-Jordan inverse in registers (as det_combine_kernel); a block whose inverse is
//            large goes through a one-sided Jacobi SVD (registers; the steps of ld_factor) and, if a singular value is
//            below the clamp, leaves U, V, s in shared memory; the others leave A^-1;
//   phase B  threads over (block, slice of the channels): per tangent channel  <G, dA> and <H[dA], dA>,  Laplacian
//            channel <G, lap A>.  Unclamped blocks: M = A^-1 dA, tr M and -tr M^2.  Clamped blocks: P = U^T dA V, then
//            <G, dA> = sum_{i in U} P_ii / s_i  and  <H[dA], dA> = <M, P>  with M of logdet_math.cuh -- H is never formed;
//   phase C  the same combine over determinants and the same local-energy assembly as det_combine_kernel, and the
//            walker's contribution to the energy statistics is replaced.
// Everything per block is compile-time sized (NM = max(n_up, n_dn), identity padding) so it lives in registers; the first
// version of this kernel ran the runtime-size functions of logdet_math.cuh for every (block, channel) out of local memory,
// one warp per walker, and a single flagged walker of N2 held its chunk for 20 ms.
#pragma once
#include "logdet_math.cuh"
#include "slogdet.cuh"

namespace psif {

constexpr int kDetFixThreads = 256;

__host__ __device__ inline int detfix_slices(int K) { return kDetFixThreads / (2 * K) > 0 ? kDetFixThreads / (2 * K) : 1; }
__host__ __device__ inline int detfix_blk_doubles(int nm) { return 2 * nm * nm + nm; }
__host__ __device__ inline int detfix_smem_doubles(int K, int T, int nm) {
  return 6 * K + 2 * K + T + 8 + 2 * K * T + 2 * K * detfix_slices(K) + 2 * K * detfix_blk_doubles(nm) + K;
}

template <int NM>
__global__ void __launch_bounds__(kDetFixThreads)
det_clamp_fixup_kernel(DetArgs a, int wpc) {
  extern __shared__ double fsm[];
  if (a.status == nullptr) return;
  const int tid = threadIdx.x;
  const int K = a.K, T = a.C - 2;
  const int S = detfix_slices(K);
  constexpr int BS = 2 * NM * NM + NM;
  double* ell = fsm;                 // [2][K]
  double* sgn = ell + 2 * K;         // [2][K]
  double* lapt = sgn + 2 * K;        // [2][K]
  double* q = lapt + 2 * K;          // [K]
  double* ck = q + K;                // [K]
  double* G = ck + K;                // [T]
  double* misc = G + T;              // [8]
  double* gs = misc + 8;             // [2][K][T]
  double* lp = gs + (size_t)2 * K * T;       // [2 K][S]   Laplacian partial of one slice of the channels
  double* blkd = lp + (size_t)2 * K * S;     // [2 K][BS]  A^-1  or  U | V | s
  int* kind = reinterpret_cast<int*>(blkd + (size_t)2 * K * BS);   // [2 K]  0 inverse, 1 clamped (SVD), 2 empty spin channel

  for (int wi = 0; wi < wpc; ++wi) {
  const long long b = (long long)blockIdx.x * wpc + wi;
  if (b >= a.B) break;
  const uint32_t st_old = a.status[b];
  if (!(st_old & PSIF_ST_CLAMP_SUSPECT)) continue;      // uniform over the CTA

  // ---- phase A: factor every block ----
  for (int idx = tid; idx < 2 * K; idx += kDetFixThreads) {
    const int sg = idx / K, k = idx - sg * K;
    const int n = a.n[sg];
    if (n <= 0) { ell[idx] = 0.0; sgn[idx] = 1.0; kind[idx] = 2; continue; }
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
    double* bd = blkd + (size_t)idx * BS;
    int kd = 0;
    if (!(fro2 <= 1e12)) {
      // the steps of ld_factor (logdet_math.cuh) on the identity-padded block, in registers
      double Wm[NM * NM], Vm[NM * NM];
#pragma unroll
      for (int i = 0; i < NM; ++i)
#pragma unroll
        for (int j = 0; j < NM; ++j) {
          double v = (i == j) ? 1.0 : 0.0;
          if (i < n && j < n) v = (double)__ldg(base + i * is + j) + ((i == j) ? kDetJitter : 0.0);
          Wm[i * NM + j] = v;
        }
      jacobi_svd<NM>(Wm, Vm);
      double sv[NM], ldc = 0.0;
      bool clamped = false;
#pragma unroll
      for (int j = 0; j < NM; ++j) {
        double n2 = 0.0;
#pragma unroll
        for (int i = 0; i < NM; ++i) n2 += Wm[i * NM + j] * Wm[i * NM + j];
        sv[j] = sqrt(n2);
        if (sv[j] < LD_MIN_SINGULAR) clamped = true;
        ldc += log(fmax(sv[j], LD_MIN_SINGULAR));
      }
      if (clamped) {
        kd = 1;
        ld = ldc;
        if (sgv == 0.0) sgv = 1.0;
        double* U = bd;
#pragma unroll
        for (int j = 0; j < NM; ++j) {
          const bool big = sv[j] >= LD_MIN_SINGULAR;
#pragma unroll
          for (int i = 0; i < NM; ++i) {
            U[i * NM + j] = big ? Wm[i * NM + j] / sv[j] : Wm[i * NM + j];     // clamped columns: raw A v_j, fixed below
            bd[NM * NM + i * NM + j] = Vm[i * NM + j];
          }
          bd[2 * NM * NM + j] = sv[j];
        }
        // left vectors of the clamped columns: modified Gram-Schmidt against all fixed columns, starting from A v_j or,
        // if that vanished, from the unit vectors in turn
        const double* svs = bd + 2 * NM * NM;
        for (int j = 0; j < NM; ++j) {
          if (svs[j] >= LD_MIN_SINGULAR) continue;
          double u[NM], w0[NM];
#pragma unroll
          for (int i = 0; i < NM; ++i) w0[i] = U[i * NM + j];
          for (int attempt = 0; attempt < 1 + NM; ++attempt) {
#pragma unroll
            for (int i = 0; i < NM; ++i) u[i] = attempt == 0 ? w0[i] : ((i == attempt - 1) ? 1.0 : 0.0);
            double nrm0 = 0.0;
#pragma unroll
            for (int i = 0; i < NM; ++i) nrm0 += u[i] * u[i];
            if (nrm0 == 0.0) continue;
            const double r0 = sqrt(nrm0);
#pragma unroll
            for (int i = 0; i < NM; ++i) u[i] /= r0;
            for (int pass = 0; pass < 2; ++pass)
              for (int o = 0; o < NM; ++o) {
                const bool done = (svs[o] >= LD_MIN_SINGULAR) || (o < j);
                if (!done || o == j) continue;
                double d = 0.0;
#pragma unroll
                for (int i = 0; i < NM; ++i) d += u[i] * U[i * NM + o];
#pragma unroll
                for (int i = 0; i < NM; ++i) u[i] -= d * U[i * NM + o];
              }
            double nrm = 0.0;
#pragma unroll
            for (int i = 0; i < NM; ++i) nrm += u[i] * u[i];
            if (nrm > 1e-6) {
              const double r1 = sqrt(nrm);
#pragma unroll
              for (int i = 0; i < NM; ++i) U[i * NM + j] = u[i] / r1;
              break;
            }
          }
        }
      }
    }
    if (kd == 0) {
#pragma unroll
      for (int e = 0; e < NM * NM; ++e) bd[e] = X[e];
    }
    kind[idx] = kd;
    ell[idx] = ld;
    sgn[idx] = sgv;
  }
  __syncthreads();

  // ---- phase B: (block, slice of channels) ----
  for (int item = tid; item < 2 * K * S; item += kDetFixThreads) {
    const int idx = item % (2 * K), slice = item / (2 * K);
    const int sg = idx / K, k = idx - sg * K;
    const int n = a.n[sg];
    const int kd = kind[idx];
    double lapacc = 0.0;
    if (kd == 2) {
      for (int c = slice; c < T; c += S) gs[(size_t)idx * T + c] = 0.0;
      lp[(size_t)idx * S + slice] = 0.0;
      continue;
    }
    const float* base = a.phi[sg] + b * a.wstride[sg] + (long long)k * a.kstride[sg];
    const long long is = a.istride[sg];
    const double* bd = blkd + (size_t)idx * BS;
    for (int c = slice; c <= T; c += S) {
      const bool lapch = c == T;
      const float* pc = base + (long long)(lapch ? a.C - 1 : 1 + c) * a.cstride;
      // the whole channel block first: a lone warp cannot hide 49 dependent trips to HBM
      float dA[NM * NM];
#pragma unroll
      for (int i = 0; i < NM; ++i)
#pragma unroll
        for (int j = 0; j < NM; ++j) dA[i * NM + j] = (i < n && j < n) ? __ldg(pc + i * is + j) : 0.0f;
      double M[NM * NM];
#pragma unroll
      for (int e = 0; e < NM * NM; ++e) M[e] = 0.0;
      double gdot, quad;
      if (kd == 0) {
        // M = A^-1 dA
#pragma unroll
        for (int i = 0; i < NM; ++i) {
          if (i < n) {
#pragma unroll
            for (int j = 0; j < NM; ++j) {
              if (j < n) {
                const double da = (double)dA[i * NM + j];
#pragma unroll
                for (int r = 0; r < NM; ++r) M[r * NM + j] += bd[r * NM + i] * da;
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
        gdot = tr;
        quad = -tr2;
      } else {
        const double* U = bd;
        const double* V = bd + NM * NM;
        const double* sv = bd + 2 * NM * NM;
        // M = dA V   (rows of dA beyond n are zero)
#pragma unroll
        for (int i = 0; i < NM; ++i) {
          if (i < n) {
#pragma unroll
            for (int kk = 0; kk < NM; ++kk) {
              if (kk < n) {
                const double da = (double)dA[i * NM + kk];
#pragma unroll
                for (int j = 0; j < NM; ++j) M[i * NM + j] += da * V[kk * NM + j];
              }
            }
          }
        }
        // P = U^T M, one (a, b) / (b, a) pair at a time
        gdot = 0.0;
        quad = 0.0;
#pragma unroll
        for (int p = 0; p < NM; ++p) {
          const double sa = sv[p];
          const bool ua = sa >= LD_MIN_SINGULAR;
#pragma unroll
          for (int r = p; r < NM; ++r) {
            double pab = 0.0, pba = 0.0;
#pragma unroll
            for (int kk = 0; kk < NM; ++kk) {
              pab += U[kk * NM + p] * M[kk * NM + r];
              if (r != p) pba += U[kk * NM + r] * M[kk * NM + p];
            }
            if (r == p) {
              if (ua) { gdot += pab / sa; quad -= pab * pab / (sa * sa); }
            } else {
              const double sb = sv[r];
              const bool ub = sb >= LD_MIN_SINGULAR;
              if (ua && ub) {
                quad -= 2.0 * pab * pba / (sa * sb);
              } else if (ua != ub) {
                const double si = ua ? sa : sb, sc = ua ? sb : sa;
                const double den = 1.0 / (si * (si * si - sc * sc));
                quad += ((si * pab + sc * pba) * pab + (si * pba + sc * pab) * pba) * den;
              }
            }
          }
        }
      }
      if (lapch) {
        lapacc += gdot;
      } else {
        gs[(size_t)idx * T + c] = gdot;
        lapacc += quad;
      }
    }
    lp[(size_t)idx * S + slice] = lapacc;
  }
  __syncthreads();

  // ---- phase C: combine (first warp) ----
  if (tid < 32) {
  const int lane = tid;
  for (int idx = lane; idx < 2 * K; idx += 32) {
    double l = 0.0;
    for (int s = 0; s < S; ++s) l += lp[(size_t)idx * S + s];
    lapt[idx] = l;
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
  }   // first warp
  __syncthreads();
  }   // walkers of this CTA
}

template <int NM>
inline int32_t det_clamp_fixup_launch_t(const DetArgs& a, cudaStream_t st) {
  const size_t smem = (size_t)detfix_smem_doubles(a.K, a.C - 2, NM) * sizeof(double);
  if (smem > 227 * 1024) return fail(PSIF_E_INVALID, "slogdet: shared memory budget exceeded%s");
  size_t& configured = dev_smem_cfg().detfix[NM];
  if (smem > 48 * 1024 && smem > configured) {
    PSIF_CUDA_CHECK(cudaFuncSetAttribute(det_clamp_fixup_kernel<NM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  auto kern = det_clamp_fixup_kernel<NM>;
  // walkers per CTA: flagged walkers of one CTA are worked through one after the other (a 1.4 % flagged N2 chunk is free
  // at 1 per CTA and costs 1.4 % of the step at 8), empty CTAs cost 8 ns each (Be: 4096 walkers) -- keep the grid <= 2048
  static const int wpc_env = [] {
    const char* e = getenv("PSIF_CLAMP_WPC");
    const int v = e ? atoi(e) : 0;
    return v >= 1 && v <= 64 ? v : 0;
  }();
  const int wpc = wpc_env ? wpc_env : (int)((a.B + 2047) / 2048);
  PSIF_LAUNCH(kern, (unsigned)cdiv(a.B, wpc), kDetFixThreads, smem, st, a, wpc);
  return PSIF_OK;
}

inline int32_t det_clamp_fixup_launch(const DetArgs& a, cudaStream_t st) {
  if (a.B <= 0 || a.status == nullptr) return PSIF_OK;
  if (a.B > 0x7fffffffLL) return fail(PSIF_E_INVALID, "slogdet: too many walkers for one launch%s");
  const int nm = a.n[0] > a.n[1] ? a.n[0] : a.n[1];
  switch (nm <= 1 ? 1 : nm) {
    case 1: return det_clamp_fixup_launch_t<1>(a, st);
    case 2: return det_clamp_fixup_launch_t<2>(a, st);
    case 3: return det_clamp_fixup_launch_t<3>(a, st);
    case 4: return det_clamp_fixup_launch_t<4>(a, st);
    case 5: return det_clamp_fixup_launch_t<5>(a, st);
    case 6: return det_clamp_fixup_launch_t<6>(a, st);
    case 7: return det_clamp_fixup_launch_t<7>(a, st);
    case 8: return det_clamp_fixup_launch_t<8>(a, st);
  }
  return fail(PSIF_E_INVALID, "slogdet: bad size%s");
}

}  // namespace psif
