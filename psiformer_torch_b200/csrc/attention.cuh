// Per-walker, per-head attention on payloads (psiformer.py:42-62):
//   s = q k^T / sqrt(hd), p = softmax(s), y = p v, for all N electrons (no mask), together with the
//   3N tangent channels and the Laplacian channel (bilinear + softmax rules).
// The sequence is the electron list (N <= 16), so the N x N part lives in shared memory of one CTA
// per (walker, head); tensor cores have nothing to do here.
//
// qkv payload rows are (b, i, c) with 3d columns [q | k | v]; head h owns columns h*hd..(h+1)*hd of each.
#pragma once
#include "common.cuh"

namespace psif {

constexpr int ATT_THREADS = 128;

struct AttSmem {
  // offsets in floats into dynamic shared memory
  int q0, k0, v0, qc, kc, vc, s0, sL, quad, mb, sT, total;
};

__host__ __device__ inline AttSmem att_layout(int N, int hd, int C) {
  AttSmem L;
  const int row = hd + 1;  // +1 float padding: conflict-free dot products across rows
  int o = 0;
  L.q0 = o; o += N * row;
  L.k0 = o; o += N * row;
  L.v0 = o; o += N * row;
  L.qc = o; o += N * row;
  L.kc = o; o += N * row;
  L.vc = o; o += N * row;
  L.s0 = o; o += N * N;
  L.sL = o; o += N * N;
  L.quad = o; o += N * N;
  L.mb = o; o += (C > 1 ? (C - 2) : 0) * N;
  L.sT = o; o += (C > 1 ? (C - 2) : 0) * N * N;
  L.total = o;
  return L;
}

__device__ __forceinline__ void att_load_rows(float* dst, const float* __restrict__ qkv, long long tok0, int N,
                                              int C, int c, int d3, int col0, int hd) {
  const int row = hd + 1;
  for (int idx = threadIdx.x; idx < N * hd; idx += ATT_THREADS) {
    const int i = idx / hd, e = idx - i * hd;
    dst[i * row + e] = __ldg(qkv + ((tok0 + i) * C + c) * (long long)d3 + col0 + e);
  }
}

__global__ void __launch_bounds__(ATT_THREADS)
attention_payload_kernel(const float* __restrict__ qkv, float* __restrict__ out, int N, int C, int d, int H) {
  extern __shared__ float sm[];
  const int hd = d / H;
  const int row = hd + 1;
  const AttSmem L = att_layout(N, hd, C);
  const long long b = blockIdx.x / H;
  const int h = (int)(blockIdx.x % H);
  const long long tok0 = b * N;
  const int d3 = 3 * d;
  const int qcol = h * hd, kcol = d + h * hd, vcol = 2 * d + h * hd;
  const float scale = rsqrtf((float)hd);
  const int NN = N * N;
  const int T = C > 1 ? C - 2 : 0;

  float *q0 = sm + L.q0, *k0 = sm + L.k0, *v0 = sm + L.v0, *qc = sm + L.qc, *kc = sm + L.kc, *vc = sm + L.vc;
  float *s0 = sm + L.s0, *sL = sm + L.sL, *quad = sm + L.quad, *mb = sm + L.mb, *sT = sm + L.sT;

  att_load_rows(q0, qkv, tok0, N, C, 0, d3, qcol, hd);
  att_load_rows(k0, qkv, tok0, N, C, 0, d3, kcol, hd);
  att_load_rows(v0, qkv, tok0, N, C, 0, d3, vcol, hd);
  __syncthreads();

  // ---- scores: value, tangents, Laplacian --------------------------------------------------
  for (int pidx = threadIdx.x; pidx < NN; pidx += ATT_THREADS) {
    const int i = pidx / N, j = pidx - i * N;
    float a = 0.f;
    for (int e = 0; e < hd; ++e) a = fmaf(q0[i * row + e], k0[j * row + e], a);
    s0[pidx] = a * scale;
    sL[pidx] = 0.f;
    quad[pidx] = 0.f;
  }
  for (int t = 0; t < T; ++t) {
    __syncthreads();
    att_load_rows(qc, qkv, tok0, N, C, 1 + t, d3, qcol, hd);
    att_load_rows(kc, qkv, tok0, N, C, 1 + t, d3, kcol, hd);
    __syncthreads();
    for (int pidx = threadIdx.x; pidx < NN; pidx += ATT_THREADS) {
      const int i = pidx / N, j = pidx - i * N;
      float a = 0.f, bq = 0.f;
      for (int e = 0; e < hd; ++e) {
        const float qce = qc[i * row + e], kce = kc[j * row + e];
        a = fmaf(qce, k0[j * row + e], a);
        a = fmaf(q0[i * row + e], kce, a);
        bq = fmaf(qce, kce, bq);
      }
      sT[t * NN + pidx] = a * scale;
      sL[pidx] += 2.0f * scale * bq;
    }
  }
  if (C > 1) {
    __syncthreads();
    att_load_rows(qc, qkv, tok0, N, C, C - 1, d3, qcol, hd);
    att_load_rows(kc, qkv, tok0, N, C, C - 1, d3, kcol, hd);
    __syncthreads();
    for (int pidx = threadIdx.x; pidx < NN; pidx += ATT_THREADS) {
      const int i = pidx / N, j = pidx - i * N;
      float a = 0.f;
      for (int e = 0; e < hd; ++e) {
        a = fmaf(qc[i * row + e], k0[j * row + e], a);
        a = fmaf(q0[i * row + e], kc[j * row + e], a);
      }
      sL[pidx] += a * scale;
    }
  }
  __syncthreads();

  // ---- softmax rows (one thread per row i; N <= 16) -----------------------------------------
  if (threadIdx.x < N) {
    const int i = threadIdx.x;
    float mx = -INFINITY;
    for (int j = 0; j < N; ++j) mx = fmaxf(mx, s0[i * N + j]);
    float den = 0.f;
    for (int j = 0; j < N; ++j) {
      const float e = expf(s0[i * N + j] - mx);
      s0[i * N + j] = e;
      den += e;
    }
    const float inv = 1.0f / den;
    for (int j = 0; j < N; ++j) s0[i * N + j] *= inv;  // s0 now holds p0
  }
  __syncthreads();
  // tangents of p: mbar[t][i] = sum_j p0[i][j] sT[t][i][j]; then pT = p0 * (sT - mbar), quad = sum_t (sT - mbar)^2
  for (int idx = threadIdx.x; idx < T * N; idx += ATT_THREADS) {
    const int t = idx / N, i = idx - t * N;
    const float* st = sT + t * NN + i * N;
    const float* p = s0 + i * N;
    float mbar = 0.f;
    for (int j = 0; j < N; ++j) mbar = fmaf(p[j], st[j], mbar);
    mb[idx] = mbar;
  }
  __syncthreads();
  for (int pidx = threadIdx.x; pidx < NN; pidx += ATT_THREADS) {
    const int i = pidx / N;
    const float p = s0[pidx];
    float qd = 0.f;
    for (int t = 0; t < T; ++t) {
      const float dv = sT[t * NN + pidx] - mb[t * N + i];
      sT[t * NN + pidx] = p * dv;
      qd = fmaf(dv, dv, qd);
    }
    quad[pidx] = qd;
  }
  __syncthreads();
  if (C > 1 && threadIdx.x < N) {
    const int i = threadIdx.x;
    const float* p = s0 + i * N;
    float a = 0.f, bq = 0.f;
    for (int j = 0; j < N; ++j) {
      a = fmaf(p[j], sL[i * N + j], a);
      bq = fmaf(p[j], quad[i * N + j], bq);
    }
    for (int j = 0; j < N; ++j) sL[i * N + j] = p[j] * ((sL[i * N + j] - a) + quad[i * N + j] - bq);  // pL
  }
  __syncthreads();

  // ---- outputs: thread owns elements (i, e) = idx, idx + ATT_THREADS, ... --------------------
  constexpr int MAXOWN = (PSIF_MAX_ELEC * 128 + ATT_THREADS - 1) / ATT_THREADS;  // hd <= 128
  float yl[MAXOWN];
#pragma unroll
  for (int r = 0; r < MAXOWN; ++r) yl[r] = 0.f;
  const int NE = N * hd;
#pragma unroll
  for (int r = 0; r < MAXOWN; ++r) {
    const int idx = threadIdx.x + r * ATT_THREADS;
    if (idx < NE) {
      const int i = idx / hd, e = idx - i * hd;
      float a = 0.f;
      for (int j = 0; j < N; ++j) a = fmaf(s0[i * N + j], v0[j * row + e], a);
      out[((tok0 + i) * C + 0) * (long long)d + qcol + e] = a;
    }
  }
  for (int t = 0; t < T; ++t) {
    __syncthreads();
    att_load_rows(vc, qkv, tok0, N, C, 1 + t, d3, vcol, hd);
    __syncthreads();
    const float* pt = sT + t * NN;
#pragma unroll
    for (int r = 0; r < MAXOWN; ++r) {
      const int idx = threadIdx.x + r * ATT_THREADS;
      if (idx < NE) {
        const int i = idx / hd, e = idx - i * hd;
        float a = 0.f, cr = 0.f;
        for (int j = 0; j < N; ++j) {
          const float ptj = pt[i * N + j], vcj = vc[j * row + e];
          a = fmaf(ptj, v0[j * row + e], a);
          a = fmaf(s0[i * N + j], vcj, a);
          cr = fmaf(ptj, vcj, cr);
        }
        out[((tok0 + i) * C + 1 + t) * (long long)d + qcol + e] = a;
        yl[r] += 2.0f * cr;
      }
    }
  }
  if (C > 1) {
    __syncthreads();
    att_load_rows(vc, qkv, tok0, N, C, C - 1, d3, vcol, hd);
    __syncthreads();
#pragma unroll
    for (int r = 0; r < MAXOWN; ++r) {
      const int idx = threadIdx.x + r * ATT_THREADS;
      if (idx < NE) {
        const int i = idx / hd, e = idx - i * hd;
        float a = yl[r];
        for (int j = 0; j < N; ++j) {
          a = fmaf(sL[i * N + j], v0[j * row + e], a);
          a = fmaf(s0[i * N + j], vc[j * row + e], a);
        }
        out[((tok0 + i) * C + C - 1) * (long long)d + qcol + e] = a;
      }
    }
  }
}

inline int32_t attention_payload(const float* qkv, float* out, long long B, int N, int C, int d, int H,
                                 cudaStream_t st) {
  if (B <= 0) return PSIF_OK;
  if (H <= 0 || d % H != 0) return fail(PSIF_E_INVALID, "attention: n_embd must be divisible by n_head%s");
  const int hd = d / H;
  if (N > PSIF_MAX_ELEC || hd > 128) return fail(PSIF_E_INVALID, "attention: N > 16 or head_dim > 128 unsupported%s");
  const AttSmem L = att_layout(N, hd, C);
  const size_t smem = (size_t)L.total * sizeof(float);
  if (smem > 220 * 1024) return fail(PSIF_E_INVALID, "attention: shared memory budget exceeded%s");
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    PSIF_CUDA_CHECK(cudaFuncSetAttribute(attention_payload_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  const long long grid = B * H;
  if (grid > 0x7fffffffLL) return fail(PSIF_E_INVALID, "attention: grid too large%s");
  PSIF_LAUNCH(attention_payload_kernel, (unsigned)grid, ATT_THREADS, smem, st, qkv, out, N, C, d, H);
  return PSIF_OK;
}

}  // namespace psif
