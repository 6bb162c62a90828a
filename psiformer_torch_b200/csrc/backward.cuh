// Parameter backward of log|psi|: d(sum_b gbar_b log|psi|(x_b)) / d(params), the gradient that
// loss.backward() needs at train.py:141-148.  Reverse mode through the VALUE network (C = 1 payload); the batch is
// the sample set of one training step (tokens = walkers * N), so these kernels favour determinism and clarity over
// peak throughput: every reduction over tokens goes through fixed-order partial sums (no float atomics).
#pragma once
#include "common.cuh"
#include "elementwise.cuh"
#include "smallmat.cuh"

namespace psif {

// ---- C[N1][N2] (+)= sum_m A[m][N1] * B[m][N2]   (weight gradients dW = dY^T X) -------------------------------
// grid (ceil(N2/64), ceil(N1/64), S): block z handles rows m in [z*chunk, (z+1)*chunk) and writes partial[z].
__global__ void __launch_bounds__(256)
gemm_at_b_partial_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ partial,
                         long long M, int N1, int N2, long long chunk) {
  __shared__ float As[16][64 + 1], Bs[16][64 + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int n1_0 = blockIdx.y * 64, n2_0 = blockIdx.x * 64;
  const long long m_lo = (long long)blockIdx.z * chunk;
  long long m_hi = m_lo + chunk;
  if (m_hi > M) m_hi = M;
  float acc[4][4] = {};
  for (long long m0 = m_lo; m0 < m_hi; m0 += 16) {
    for (int idx = threadIdx.x; idx < 16 * 64; idx += 256) {
      const int r = idx >> 6, c = idx & 63;
      const long long m = m0 + r;
      As[r][c] = (m < m_hi && n1_0 + c < N1) ? A[m * N1 + n1_0 + c] : 0.f;
      Bs[r][c] = (m < m_hi && n2_0 + c < N2) ? B[m * N2 + n2_0 + c] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[r][ty * 4 + i]; b[i] = Bs[r][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* P = partial + (size_t)blockIdx.z * N1 * N2;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = n1_0 + ty * 4 + i, c = n2_0 + tx * 4 + j;
      if (r < N1 && c < N2) P[(size_t)r * N2 + c] = acc[i][j];
    }
}

// out[i] (+)= sum_s partial[s][i]
__global__ void reduce_partials_kernel(const float* __restrict__ partial, float* __restrict__ out, long long n, int S,
                                       int accumulate) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int z = 0; z < S; ++z) s += partial[(size_t)z * n + i];
  out[i] = accumulate ? out[i] + s : s;
}

// partial[z][w] = sum over rows of chunk z of in[r][w] * (scale ? scale[r] : 1)
__global__ void __launch_bounds__(256)
colsum_partial_kernel(const float* __restrict__ in, float* __restrict__ partial, long long R, int W, long long chunk) {
  __shared__ float red[8][32];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + cx;
  const long long r_lo = (long long)blockIdx.y * chunk;
  long long r_hi = r_lo + chunk;
  if (r_hi > R) r_hi = R;
  float s = 0.f;
  if (col < W)
    for (long long r = r_lo + ry; r < r_hi; r += 8) s += in[r * W + col];
  red[ry][cx] = s;
  __syncthreads();
  if (ry == 0 && col < W) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][cx];
    partial[(size_t)blockIdx.y * W + col] = t;
  }
}

__global__ void transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int R, int Ccols) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int r = r0 + j, c = c0 + threadIdx.x;
    tile[j][threadIdx.x] = (r < R && c < Ccols) ? in[(size_t)r * Ccols + c] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int c = c0 + j, r = r0 + threadIdx.x;
    if (r < R && c < Ccols) out[(size_t)c * R + r] = tile[threadIdx.x][j];
  }
}

// ---- LayerNorm backward: dx = dres + s (g - mean(g) - xhat mean(g xhat)), g = dy * gamma; prod = dy * xhat -------
// one warp per token, generic d (lane strides)
__global__ void __launch_bounds__(256)
layernorm_backward_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ gamma,
                          const float* dres, float* dx, float* __restrict__ prod, long long tokens, int d) {
  const long long tok = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (tok >= tokens) return;
  const int lane = threadIdx.x & 31;
  const float* xp = x + tok * d;
  const float* gp = dy + tok * d;
  const float inv_d = 1.0f / (float)d;
  float sum = 0.f;
  for (int e = lane; e < d; e += 32) sum += xp[e];
  const float mean = warp_sum(sum) * inv_d;
  float sq = 0.f;
  for (int e = lane; e < d; e += 32) { const float c = xp[e] - mean; sq += c * c; }
  const float s = rsqrtf(warp_sum(sq) * inv_d + kLnEps);
  float sg = 0.f, sgx = 0.f;
  for (int e = lane; e < d; e += 32) {
    const float xh = (xp[e] - mean) * s, g = gp[e] * gamma[e];
    sg += g; sgx += g * xh;
  }
  const float mg = warp_sum(sg) * inv_d, mgx = warp_sum(sgx) * inv_d;
  for (int e = lane; e < d; e += 32) {
    const float xh = (xp[e] - mean) * s, g = gp[e] * gamma[e];
    const float v = s * (g - mg - xh * mgx);
    dx[tok * d + e] = (dres ? dres[tok * d + e] : 0.f) + v;
    prod[tok * d + e] = gp[e] * xh;
  }
}

// ---- GELU backward: g = gelu(u); dg <- dg * gelu'(u) (in place) ------------------------------------------------------
__global__ void gelu_backward_kernel(const float* __restrict__ u, float* __restrict__ g, float* __restrict__ dg, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v, g1, g2;
  gelu_tanh_d2(u[i], v, g1, g2);
  g[i] = v;
  dg[i] *= g1;
}

// ---- attention backward, one CTA per (walker, head) -------------------------------------------------------------
__global__ void __launch_bounds__(128)
attention_backward_kernel(const float* __restrict__ qkv, const float* __restrict__ dy, float* __restrict__ dqkv, int N, int d,
                          int H) {
  extern __shared__ float asm_[];
  const int hd = d / H, row = hd + 1;
  const long long b = blockIdx.x / H;
  const int h = (int)(blockIdx.x % H);
  const long long tok0 = b * N;
  const int d3 = 3 * d, qcol = h * hd, kcol = d + h * hd, vcol = 2 * d + h * hd;
  const float scale = rsqrtf((float)hd);
  float* q = asm_;
  float* k = q + N * row;
  float* v = k + N * row;
  float* g = v + N * row;      // dy
  float* p = g + N * row;      // [N][N]
  float* ds = p + N * N;       // [N][N]
  for (int idx = threadIdx.x; idx < N * hd; idx += 128) {
    const int i = idx / hd, e = idx - i * hd;
    const float* base = qkv + (tok0 + i) * (long long)d3;
    q[i * row + e] = base[qcol + e];
    k[i * row + e] = base[kcol + e];
    v[i * row + e] = base[vcol + e];
    g[i * row + e] = dy[(tok0 + i) * (long long)d + qcol + e];
  }
  __syncthreads();
  for (int pidx = threadIdx.x; pidx < N * N; pidx += 128) {
    const int i = pidx / N, j = pidx - i * N;
    float a = 0.f, dp = 0.f;
    for (int e = 0; e < hd; ++e) {
      a = fmaf(q[i * row + e], k[j * row + e], a);
      dp = fmaf(g[i * row + e], v[j * row + e], dp);
    }
    p[pidx] = a * scale;
    ds[pidx] = dp;
  }
  __syncthreads();
  if (threadIdx.x < N) {
    const int i = threadIdx.x;
    float mx = -INFINITY;
    for (int j = 0; j < N; ++j) mx = fmaxf(mx, p[i * N + j]);
    float den = 0.f;
    for (int j = 0; j < N; ++j) { const float e = expf(p[i * N + j] - mx); p[i * N + j] = e; den += e; }
    const float inv = 1.0f / den;
    float dot = 0.f;
    for (int j = 0; j < N; ++j) { p[i * N + j] *= inv; dot = fmaf(p[i * N + j], ds[i * N + j], dot); }
    for (int j = 0; j < N; ++j) ds[i * N + j] = p[i * N + j] * (ds[i * N + j] - dot) * scale;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < N * hd; idx += 128) {
    const int i = idx / hd, e = idx - i * hd;
    float dq = 0.f, dk = 0.f, dv = 0.f;
    for (int j = 0; j < N; ++j) {
      dq = fmaf(ds[i * N + j], k[j * row + e], dq);
      dk = fmaf(ds[j * N + i], q[j * row + e], dk);
      dv = fmaf(p[j * N + i], g[j * row + e], dv);
    }
    float* o = dqkv + (tok0 + i) * (long long)d3;
    o[qcol + e] = dq;
    o[kcol + e] = dk;
    o[vcol + e] = dv;
  }
}

// ---- envelope forward that keeps its factors: env[t][col], phi = lin * env (own-spin columns, others 0) -----------
__global__ void __launch_bounds__(128)
orbital_envelope_save_kernel(const float* __restrict__ lin, const float* __restrict__ x, const float* __restrict__ sigma,
                             const float* __restrict__ pi, float* __restrict__ env, float* __restrict__ phi, int N, int n_up,
                             int Kup, int Korb, Nuclei nuc) {
  const long long tok = blockIdx.x;
  const int i = (int)(tok % N);
  const int col0 = i < n_up ? 0 : Kup;
  const int ncol = i < n_up ? Kup : Korb - Kup;
  const float px = x[tok * 3 + 0], py = x[tok * 3 + 1], pz = x[tok * 3 + 2];
  for (int c = threadIdx.x; c < Korb; c += blockDim.x) {
    float e0 = 0.f;
    const bool own = c >= col0 && c < col0 + ncol;
    if (own) {
      for (int a = 0; a < nuc.natom; ++a) {
        const float dx = px - nuc.R[a][0], dy = py - nuc.R[a][1], dz = pz - nuc.R[a][2];
        const float r = sqrtf(dx * dx + dy * dy + dz * dz);
        e0 += pi[a * Korb + c] * expf(-r * sigma[a * Korb + c]);
      }
    }
    env[tok * Korb + c] = e0;
    phi[tok * Korb + c] = own ? lin[tok * Korb + c] * e0 : 0.f;
  }
}

// ---- determinant backward: dPhi = gbar c_k A^-T; writes d(lin) = dPhi*env, d(env) = dPhi*lin, gbar*c_k -------------
struct DetBwdArgs {
  const float* phi;      // [T][Korb]
  const float* lin;      // [T][Korb]
  const float* env;      // [T][Korb]
  const float* w;        // [K]
  const float* gbar;     // [B]
  float* dlin;           // [T][Korb] (pre-zeroed)
  float* denv;           // [T][Korb] (pre-zeroed)
  float* ck;             // [B][K]   gbar_b * c_bk
  long long B;
  int N, K, nu, nd, Kup, Korb, tpw, wpb;
};

template <int NM>
__global__ void __launch_bounds__(128)
det_backward_kernel(DetBwdArgs a) {
  extern __shared__ double bsm[];
  const int K = a.K, tid = threadIdx.x;
  const int wl = tid / a.tpw, rem = tid - wl * a.tpw;
  const int sg = rem / K, k = rem - sg * K;
  const long long b = (long long)blockIdx.x * a.wpb + wl;
  const bool in_slot = wl < a.wpb, active = in_slot && b < a.B;
  double* W = bsm + (size_t)(in_slot ? wl : 0) * (5 * K);
  double *ell = W, *sgn = W + 2 * K, *ck = W + 4 * K;
  double X[NM * NM];
  const int n = sg ? a.nd : a.nu;
  const long long tok0 = b * a.N + (sg ? a.nu : 0);
  const int col0 = (sg ? a.Kup : 0) + k * n;
  if (active) {
#pragma unroll
    for (int i = 0; i < NM; ++i)
#pragma unroll
      for (int j = 0; j < NM; ++j) {
        double v = (i == j) ? 1.0 : 0.0;
        if (i < n && j < n) v = (double)a.phi[(tok0 + i) * a.Korb + col0 + j] + ((i == j) ? kDetJitter : 0.0);
        X[i * NM + j] = v;
      }
    double ld, sv, mp;
    gj_inverse<NM>(X, ld, sv, mp);
    ell[sg * K + k] = ld;
    sgn[sg * K + k] = sv;
  }
  __syncthreads();
  if (active && rem == 0) {
    double m0 = -INFINITY, m1 = -INFINITY;
    for (int kk = 0; kk < K; ++kk) { m0 = fmax(m0, ell[kk]); m1 = fmax(m1, ell[K + kk]); }
    double S = 0.0;
    for (int kk = 0; kk < K; ++kk) {
      const double D = (double)a.w[kk] * sgn[kk] * sgn[K + kk] * exp(ell[kk] - m0 + ell[K + kk] - m1);
      ck[kk] = D;
      S += D;
    }
    const double gb = (double)a.gbar[b];
    // below the 1e-12 floor of logdet_matmul.py:68 the reference's clamp has zero gradient
    const double inv = fabs(S) < kOutputFloor ? 0.0 : 1.0 / S;
    for (int kk = 0; kk < K; ++kk) {
      ck[kk] *= inv;
      a.ck[b * K + kk] = (float)(gb * ck[kk]);
    }
  }
  __syncthreads();
  if (active) {
    const double coef = (double)a.gbar[b] * ck[k];
#pragma unroll
    for (int i = 0; i < NM; ++i)
#pragma unroll
      for (int j = 0; j < NM; ++j)
        if (i < n && j < n) {
          const long long o = (tok0 + i) * a.Korb + col0 + j;
          const double dphi = coef * X[j * NM + i];
          a.dlin[o] = (float)(dphi * (double)a.env[o]);
          a.denv[o] = (float)(dphi * (double)a.lin[o]);
        }
  }
}

// ---- envelope parameter gradients: thread per (atom, column) and CHUNK of walkers (blockIdx.y), fp64 partial sums in a
// fixed order; env_param_reduce_kernel adds the chunks up, again in a fixed order.  (One thread walking over ALL walkers
// took 4.9 ms of a 51 ms Be training step.)
__global__ void env_param_grad_kernel(const float* __restrict__ denv, const float* __restrict__ x, const float* __restrict__ params,
                                      size_t off_up_pi, size_t off_up_rs, size_t off_dn_pi, size_t off_dn_rs, long long B, int N,
                                      int n_up, int Kup, int Korb, Nuclei nuc, double* __restrict__ partial, long long chunk) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = nuc.natom * Korb;
  if (idx >= n) return;
  const int at = idx / Korb, col = idx - at * Korb;
  const bool up = col < Kup;
  const int Kdn = Korb - Kup;
  const size_t o_pi = up ? off_up_pi + (size_t)at * Kup + col : off_dn_pi + (size_t)at * Kdn + (col - Kup);
  const size_t o_rs = up ? off_up_rs + (size_t)at * Kup + col : off_dn_rs + (size_t)at * Kdn + (col - Kup);
  const float pi_raw = params[o_pi], rs = params[o_rs];
  const float sp = rs > 20.0f ? rs : log1pf(expf(rs));
  const float sig_un = sp + 1e-6f;
  const float sigma = fminf(fmaxf(sig_un, 1e-3f), 1e3f);
  const float pic = fminf(fmaxf(pi_raw, 1e-3f), 1e3f);
  double gpi = 0.0, gsg = 0.0;
  const int i_lo = up ? 0 : n_up, i_hi = up ? n_up : N;
  const long long b_lo = (long long)blockIdx.y * chunk;
  long long b_hi = b_lo + chunk;
  if (b_hi > B) b_hi = B;
  for (long long b = b_lo; b < b_hi; ++b)
    for (int i = i_lo; i < i_hi; ++i) {
      const long long t = b * N + i;
      const float dx = x[t * 3] - nuc.R[at][0], dy = x[t * 3 + 1] - nuc.R[at][1], dz = x[t * 3 + 2] - nuc.R[at][2];
      const float r = sqrtf(dx * dx + dy * dy + dz * dz);
      const float ex = expf(-r * sigma);
      const float g = denv[t * Korb + col];
      gpi += (double)(g * ex);
      gsg += (double)(g * pic * (-r) * ex);
    }
  partial[((size_t)blockIdx.y * n + idx) * 2 + 0] = gpi;
  partial[((size_t)blockIdx.y * n + idx) * 2 + 1] = gsg;
}
__global__ void env_param_reduce_kernel(const double* __restrict__ partial, int S, const float* __restrict__ params, size_t off_up_pi,
                                        size_t off_up_rs, size_t off_dn_pi, size_t off_dn_rs, int Kup, int Korb, int natom,
                                        float* __restrict__ gparams) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = natom * Korb;
  if (idx >= n) return;
  const int at = idx / Korb, col = idx - at * Korb;
  const bool up = col < Kup;
  const int Kdn = Korb - Kup;
  const size_t o_pi = up ? off_up_pi + (size_t)at * Kup + col : off_dn_pi + (size_t)at * Kdn + (col - Kup);
  const size_t o_rs = up ? off_up_rs + (size_t)at * Kup + col : off_dn_rs + (size_t)at * Kdn + (col - Kup);
  const float pi_raw = params[o_pi], rs = params[o_rs];
  const float sp = rs > 20.0f ? rs : log1pf(expf(rs));
  const float sig_un = sp + 1e-6f;
  const bool pi_live = pi_raw >= 1e-3f && pi_raw <= 1e3f, sg_live = sig_un >= 1e-3f && sig_un <= 1e3f;
  double gpi = 0.0, gsg = 0.0;
  for (int z = 0; z < S; ++z) {
    gpi += partial[((size_t)z * n + idx) * 2 + 0];
    gsg += partial[((size_t)z * n + idx) * 2 + 1];
  }
  const float dsp = rs > 20.0f ? 1.0f : 1.0f / (1.0f + expf(-rs));   // softplus'
  gparams[o_pi] += pi_live ? (float)gpi : 0.f;
  gparams[o_rs] += sg_live ? (float)gsg * dsp : 0.f;
}

// ---- embedding gradients: dW0[e][f] = sum_t dH[t][e] feat_t[f] -----------------------------------------------------------
// thread per (e, f) and CHUNK of tokens (blockIdx.y): fp64 partial sums in a fixed order, then embed_grad_reduce_kernel.
// (One thread walking over ALL tokens took 10.9 ms of a 51 ms Be training step.)
__global__ void embed_grad_kernel(const float* __restrict__ dh, const float* __restrict__ x, long long T, int d, Nuclei nuc,
                                  double* __restrict__ partial, long long chunk) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int nf = 4 * nuc.natom;
  if (idx >= d * nf) return;
  const int e = idx / nf, f = idx - e * nf;
  const int at = f >> 2, comp = f & 3;
  const long long t_lo = (long long)blockIdx.y * chunk;
  long long t_hi = t_lo + chunk;
  if (t_hi > T) t_hi = T;
  double s = 0.0;
  for (long long t = t_lo; t < t_hi; ++t) {
    const float dx = x[t * 3] - nuc.R[at][0], dy = x[t * 3 + 1] - nuc.R[at][1], dz = x[t * 3 + 2] - nuc.R[at][2];
    const float feat = comp == 0 ? dx : comp == 1 ? dy : comp == 2 ? dz : sqrtf(dx * dx + dy * dy + dz * dz);
    s += (double)(dh[t * d + e] * feat);
  }
  partial[(size_t)blockIdx.y * d * nf + idx] = s;
}
__global__ void embed_grad_reduce_kernel(const double* __restrict__ partial, int S, int n, float* __restrict__ gW0) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  double s = 0.0;
  for (int z = 0; z < S; ++z) s += partial[(size_t)z * n + idx];
  gW0[idx] += (float)s;
}

// ---- Jastrow parameter gradients and the det_logits gradient (single block, fixed order) ------------------------------------
__global__ void jastrow_logits_grad_kernel(const float* __restrict__ x, const float* __restrict__ gbar, const float* __restrict__ ckbuf,
                                           const float* __restrict__ w, const float* __restrict__ alpha /*[anti, par]*/, long long B,
                                           int N, int n_up, int K, float* __restrict__ g_alpha, float* __restrict__ g_logits) {
  __shared__ double red[256];
  const int tid = threadIdx.x;
  const double a_anti = alpha[0], a_par = alpha[1];
  double s_par = 0.0, s_anti = 0.0, s_g = 0.0;
  for (long long b = tid; b < B; b += blockDim.x) {
    const double gb = gbar[b];
    s_g += gb;
    for (int i = 0; i < N; ++i)
      for (int j = i + 1; j < N; ++j) {
        const double dx = (double)x[(b * N + i) * 3] - x[(b * N + j) * 3], dy = (double)x[(b * N + i) * 3 + 1] - x[(b * N + j) * 3 + 1],
                     dz = (double)x[(b * N + i) * 3 + 2] - x[(b * N + j) * 3 + 2];
        const double r = sqrt(dx * dx + dy * dy + dz * dz + kJastrowEps);
        const bool same = (i < n_up) == (j < n_up);
        const double al = same ? a_par : a_anti, c = same ? -0.25 : -0.5;
        const double dv = c * al * (al + 2.0 * r) / ((al + r) * (al + r));
        if (same) s_par += gb * dv; else s_anti += gb * dv;
      }
  }
  auto block_sum = [&](double v) {
    red[tid] = v;
    __syncthreads();
    for (int o = blockDim.x >> 1; o > 0; o >>= 1) { if (tid < o) red[tid] += red[tid + o]; __syncthreads(); }
    const double r = red[0];
    __syncthreads();
    return r;
  };
  const double t_par = block_sum(s_par), t_anti = block_sum(s_anti), t_g = block_sum(s_g);
  if (tid == 0) { g_alpha[0] += (float)t_anti; g_alpha[1] += (float)t_par; }
  for (int kk = 0; kk < K; ++kk) {
    double s = 0.0;
    for (long long b = tid; b < B; b += blockDim.x) s += (double)ckbuf[b * K + kk];
    const double tot = block_sum(s);
    if (tid == 0) g_logits[kk] += (float)(tot - (double)w[kk] * t_g);
  }
}

// c[i] += a[i]
__global__ void axpy_add_kernel(float* __restrict__ c, const float* __restrict__ a, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) c[i] += a[i];
}

}  // namespace psif
