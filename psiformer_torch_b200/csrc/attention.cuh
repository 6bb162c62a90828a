// Per-walker, per-head attention on payloads (psiformer.py:42-62):
//   s = q k^T / sqrt(hd), p = softmax(s), y = p v, for all N electrons (no mask), together with the
//   3N tangent channels and the Laplacian channel (bilinear + softmax rules).
// The sequence is the electron list (N <= 16), so the N x N part lives in shared memory of one CTA
// per (walker, head); tensor cores have nothing to do here.
//
// qkv payload rows are (b, i, c) with 3d columns [q | k | v]; head h owns columns h*hd..(h+1)*hd of each.
#pragma once
#include <cstdlib>
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

// ------------------------------------------------------------------------------------------------
// v2 (head_dim % 4 == 0): same mathematics, organised so that every thread has work for small N too.
//   * the tangent channels are processed in groups of G channels held in shared memory at once; a work item of
//     the score phase is a (channel, i, j) triple and of the output phase a (channel, i, 4 columns) triple,
//     so a 4-electron system still fills 256 threads;
//   * all shared/global traffic is float4 (row stride hd+4 keeps quarter-warps conflict free);
//   * sums over channels (the 2 grad.grad cross terms of the Laplacian rows) go through a per-group scratch
//     and are reduced by a fixed owner thread: deterministic, no atomics.
// ------------------------------------------------------------------------------------------------
constexpr int ATT2_THREADS = 256;

struct Att2Smem {
  int q0, k0, v0, ga, gb, gv, lq, lk, lv, s0, sL, quad, mb, cross, sT, total, G, pre;
};

__host__ __device__ inline Att2Smem att2_layout(int N, int hd, int C) {
  Att2Smem L;
  const int RS = hd + 4;
  const int T = C > 1 ? C - 2 : 0;
  int G = T > 0 ? 96 / N : 0;
  if (G > T) G = T;
  if (T > 0 && G < 1) G = 1;
  const int grows = (G > 1 ? G : 1) * N;
  int o = 0;
  L.q0 = o; o += N * RS;
  L.k0 = o; o += N * RS;
  L.v0 = o; o += N * RS;
  L.ga = o; o += grows * RS;
  L.gb = o; o += grows * RS;
  // "prefetch" layout (small systems): when one group covers all tangent channels and everything fits in 56 KiB,
  // the value group and the Laplacian rows get their own buffers so that EVERY global read of the CTA is issued
  // with cp.async at kernel start (one HBM round trip instead of four dependent ones)
  const int extra = grows * RS + 3 * N * RS;
  const int rest = 3 * N * N + T * N + (G > 1 ? G : 1) * N * N + T * N * N;
  L.pre = (T > 0 && G >= T && (o + extra + rest) * 4 <= 56 * 1024) ? 1 : 0;
  if (L.pre) {
    L.gv = o; o += grows * RS;
    L.lq = o; o += N * RS;
    L.lk = o; o += N * RS;
    L.lv = o; o += N * RS;
  } else {
    L.gv = L.ga; L.lq = L.ga; L.lk = L.gb; L.lv = L.ga;
  }
  L.s0 = o; o += N * N;
  L.sL = o; o += N * N;
  L.quad = o; o += N * N;
  L.mb = o; o += T * N;
  L.cross = o; o += (G > 1 ? G : 1) * N * N;
  L.sT = o; o += T * N * N;
  o = (o + 3) & ~3;
  L.total = o;
  L.G = G;
  return L;
}

// rows [r0, r0+nrows) of a group buffer <- channel c0 + r / N, electron r % N, columns col..col+hd of the qkv payload
__device__ __forceinline__ void att2_load(float* dst, const float* __restrict__ qkv, long long tok0, int N, int C, int c0,
                                          int nrows, int d3, int col, int hd) {
  const int RS = hd + 4, h4 = hd >> 2;
  for (int idx = threadIdx.x; idx < nrows * h4; idx += ATT2_THREADS) {
    const int r = idx / h4, e4 = idx - r * h4;
    const int g = r / N, i = r - g * N;
    const float4 v = __ldg(reinterpret_cast<const float4*>(qkv + ((tok0 + i) * C + c0 + g) * (long long)d3 + col) + e4);
    *reinterpret_cast<float4*>(dst + r * RS + 4 * e4) = v;
  }
}

__device__ __forceinline__ void att2_load_async(float* dst, const float* __restrict__ qkv, long long tok0, int N, int C, int c0,
                                                int nrows, int d3, int col, int hd) {
  const int RS = hd + 4, h4 = hd >> 2;
  for (int idx = threadIdx.x; idx < nrows * h4; idx += ATT2_THREADS) {
    const int r = idx / h4, e4 = idx - r * h4;
    const int g = r / N, i = r - g * N;
    const float* src = qkv + ((tok0 + i) * C + c0 + g) * (long long)d3 + col + 4 * e4;
    const uint32_t sdst = (uint32_t)__cvta_generic_to_shared(dst + r * RS + 4 * e4);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sdst), "l"(src) : "memory");
  }
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int NLEFT>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(NLEFT) : "memory"); }

__device__ __forceinline__ float dot4(const float4 a, const float4 b, float acc) {
  acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
  return acc;
}
__device__ __forceinline__ void axpy4(float4& y, const float a, const float4 x) {
  y.x = fmaf(a, x.x, y.x); y.y = fmaf(a, x.y, y.y); y.z = fmaf(a, x.z, y.z); y.w = fmaf(a, x.w, y.w);
}

__global__ void __launch_bounds__(ATT2_THREADS, 4)
attention_payload_v2_kernel(const float* __restrict__ qkv, float* __restrict__ out, int N, int C, int d, int H) {
  extern __shared__ __align__(16) float sm2[];
  const int hd = d / H, RS = hd + 4, h4 = hd >> 2;
  const Att2Smem L = att2_layout(N, hd, C);
  const long long b = blockIdx.x / H;
  const int h = (int)(blockIdx.x % H);
  const long long tok0 = b * N;
  const int d3 = 3 * d;
  const int qcol = h * hd, kcol = d + h * hd, vcol = 2 * d + h * hd;
  const float scale = rsqrtf((float)hd);
  const int NN = N * N, T = C > 1 ? C - 2 : 0, G = L.G;
  const int tid = threadIdx.x;
  float *q0 = sm2 + L.q0, *k0 = sm2 + L.k0, *v0 = sm2 + L.v0, *ga = sm2 + L.ga, *gb = sm2 + L.gb;
  float *gv = sm2 + L.gv, *lq = sm2 + L.lq, *lk = sm2 + L.lk, *lv = sm2 + L.lv;
  const bool pre = L.pre != 0;
  float *s0 = sm2 + L.s0, *sL = sm2 + L.sL, *quad = sm2 + L.quad, *mb = sm2 + L.mb, *cross = sm2 + L.cross, *sT = sm2 + L.sT;

  if (pre) {
    att2_load_async(q0, qkv, tok0, N, C, 0, N, d3, qcol, hd);
    att2_load_async(k0, qkv, tok0, N, C, 0, N, d3, kcol, hd);
    att2_load_async(v0, qkv, tok0, N, C, 0, N, d3, vcol, hd);
    cp_async_commit();
    att2_load_async(ga, qkv, tok0, N, C, 1, T * N, d3, qcol, hd);
    att2_load_async(gb, qkv, tok0, N, C, 1, T * N, d3, kcol, hd);
    att2_load_async(lq, qkv, tok0, N, C, C - 1, N, d3, qcol, hd);
    att2_load_async(lk, qkv, tok0, N, C, C - 1, N, d3, kcol, hd);
    cp_async_commit();
    att2_load_async(gv, qkv, tok0, N, C, 1, T * N, d3, vcol, hd);
    att2_load_async(lv, qkv, tok0, N, C, C - 1, N, d3, vcol, hd);
    cp_async_commit();
    cp_async_wait<2>();
  } else {
    att2_load(q0, qkv, tok0, N, C, 0, N, d3, qcol, hd);
    att2_load(k0, qkv, tok0, N, C, 0, N, d3, kcol, hd);
    att2_load(v0, qkv, tok0, N, C, 0, N, d3, vcol, hd);
  }
  __syncthreads();
  for (int pidx = tid; pidx < NN; pidx += ATT2_THREADS) {
    const int i = pidx / N, j = pidx - i * N;
    float a = 0.f;
    for (int e4 = 0; e4 < h4; ++e4)
      a = dot4(*reinterpret_cast<const float4*>(q0 + i * RS + 4 * e4), *reinterpret_cast<const float4*>(k0 + j * RS + 4 * e4), a);
    s0[pidx] = a * scale;
    sL[pidx] = 0.f;
  }
  // ---- scores of the tangent channels, G channels at a time -------------------------------------------
  for (int t0 = 0; t0 < T; t0 += G) {
    const int g = (T - t0) < G ? (T - t0) : G;
    if (pre) {
      cp_async_wait<1>();
    } else {
      __syncthreads();
      att2_load(ga, qkv, tok0, N, C, 1 + t0, g * N, d3, qcol, hd);
      att2_load(gb, qkv, tok0, N, C, 1 + t0, g * N, d3, kcol, hd);
    }
    __syncthreads();
    for (int idx = tid; idx < g * NN; idx += ATT2_THREADS) {
      const int gl = idx / NN, pidx = idx - gl * NN;
      const int i = pidx / N, j = pidx - i * N;
      const float* qc = ga + (gl * N + i) * RS;
      const float* kc = gb + (gl * N + j) * RS;
      float a = 0.f, bq = 0.f;
      for (int e4 = 0; e4 < h4; ++e4) {
        const float4 qv = *reinterpret_cast<const float4*>(qc + 4 * e4);
        const float4 kv = *reinterpret_cast<const float4*>(kc + 4 * e4);
        a = dot4(qv, *reinterpret_cast<const float4*>(k0 + j * RS + 4 * e4), a);
        a = dot4(*reinterpret_cast<const float4*>(q0 + i * RS + 4 * e4), kv, a);
        bq = dot4(qv, kv, bq);
      }
      sT[(t0 + gl) * NN + pidx] = a * scale;
      cross[gl * NN + pidx] = bq;
    }
    __syncthreads();
    for (int pidx = tid; pidx < NN; pidx += ATT2_THREADS) {
      float acc = 0.f;
      for (int gl = 0; gl < g; ++gl) acc += cross[gl * NN + pidx];
      sL[pidx] += 2.0f * scale * acc;
    }
  }
  if (C > 1) {  // Laplacian channel of q, k
    if (!pre) {
      __syncthreads();
      att2_load(lq, qkv, tok0, N, C, C - 1, N, d3, qcol, hd);
      att2_load(lk, qkv, tok0, N, C, C - 1, N, d3, kcol, hd);
      __syncthreads();
    }
    for (int pidx = tid; pidx < NN; pidx += ATT2_THREADS) {
      const int i = pidx / N, j = pidx - i * N;
      float a = 0.f;
      for (int e4 = 0; e4 < h4; ++e4) {
        a = dot4(*reinterpret_cast<const float4*>(lq + i * RS + 4 * e4), *reinterpret_cast<const float4*>(k0 + j * RS + 4 * e4), a);
        a = dot4(*reinterpret_cast<const float4*>(q0 + i * RS + 4 * e4), *reinterpret_cast<const float4*>(lk + j * RS + 4 * e4), a);
      }
      sL[pidx] += a * scale;
    }
  }
  __syncthreads();
  // ---- softmax and its derivatives (tiny) -----------------------------------------------------------------
  if (tid < N) {
    const int i = tid;
    float mx = -INFINITY;
    for (int j = 0; j < N; ++j) mx = fmaxf(mx, s0[i * N + j]);
    float den = 0.f;
    for (int j = 0; j < N; ++j) {
      const float e = expf(s0[i * N + j] - mx);
      s0[i * N + j] = e;
      den += e;
    }
    const float inv = 1.0f / den;
    for (int j = 0; j < N; ++j) s0[i * N + j] *= inv;
  }
  __syncthreads();
  for (int idx = tid; idx < T * N; idx += ATT2_THREADS) {
    const int t = idx / N, i = idx - t * N;
    const float* st = sT + t * NN + i * N;
    const float* p = s0 + i * N;
    float m = 0.f;
    for (int j = 0; j < N; ++j) m = fmaf(p[j], st[j], m);
    mb[idx] = m;
  }
  __syncthreads();
  for (int pidx = tid; pidx < NN; pidx += ATT2_THREADS) {
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
  if (C > 1 && tid < N) {
    const int i = tid;
    const float* p = s0 + i * N;
    float a = 0.f, bq = 0.f;
    for (int j = 0; j < N; ++j) {
      a = fmaf(p[j], sL[i * N + j], a);
      bq = fmaf(p[j], quad[i * N + j], bq);
    }
    for (int j = 0; j < N; ++j) sL[i * N + j] = p[j] * ((sL[i * N + j] - a) + quad[i * N + j] - bq);
  }
  __syncthreads();
  // ---- outputs -----------------------------------------------------------------------------------------------
  const int own = N * h4;                    // (i, e4) items owned by threads tid < own (<= 16*32 ... see launcher)
  const int oi = tid / h4, oe = tid - oi * h4;
  float4 yl = make_float4(0.f, 0.f, 0.f, 0.f);
  if (tid < own) {
    float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = 0; j < N; ++j) axpy4(y, s0[oi * N + j], *reinterpret_cast<const float4*>(v0 + j * RS + 4 * oe));
    *(reinterpret_cast<float4*>(out + ((tok0 + oi) * C + 0) * (long long)d + qcol) + oe) = y;
  }
  for (int t0 = 0; t0 < T; t0 += G) {
    const int g = (T - t0) < G ? (T - t0) : G;
    if (pre) {
      cp_async_wait<0>();
    } else {
      __syncthreads();
      att2_load(gv, qkv, tok0, N, C, 1 + t0, g * N, d3, vcol, hd);
    }
    __syncthreads();
    for (int idx = tid; idx < g * own; idx += ATT2_THREADS) {
      const int gl = idx / own, r = idx - gl * own;
      const int i = r / h4, e4 = r - i * h4;
      const float* pt = sT + (t0 + gl) * NN + i * N;
      float4 y = make_float4(0.f, 0.f, 0.f, 0.f), cr = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int j = 0; j < N; ++j) {
        const float4 vcj = *reinterpret_cast<const float4*>(gv + (gl * N + j) * RS + 4 * e4);
        const float ptj = pt[j];
        axpy4(y, ptj, *reinterpret_cast<const float4*>(v0 + j * RS + 4 * e4));
        axpy4(y, s0[i * N + j], vcj);
        axpy4(cr, ptj, vcj);
      }
      *(reinterpret_cast<float4*>(out + ((tok0 + i) * C + 1 + t0 + gl) * (long long)d + qcol) + e4) = y;
      *reinterpret_cast<float4*>(gb + (gl * N + i) * RS + 4 * e4) = cr;
    }
    __syncthreads();
    if (tid < own) {
      for (int gl = 0; gl < g; ++gl) {
        const float4 c4 = *reinterpret_cast<const float4*>(gb + (gl * N + oi) * RS + 4 * oe);
        yl.x += c4.x; yl.y += c4.y; yl.z += c4.z; yl.w += c4.w;
      }
    }
  }
  if (C > 1) {
    if (!pre) {
      __syncthreads();
      att2_load(lv, qkv, tok0, N, C, C - 1, N, d3, vcol, hd);
      __syncthreads();
    }
    if (tid < own) {
      float4 y = make_float4(2.0f * yl.x, 2.0f * yl.y, 2.0f * yl.z, 2.0f * yl.w);
      for (int j = 0; j < N; ++j) {
        axpy4(y, sL[oi * N + j], *reinterpret_cast<const float4*>(v0 + j * RS + 4 * oe));
        axpy4(y, s0[oi * N + j], *reinterpret_cast<const float4*>(lv + j * RS + 4 * oe));
      }
      *(reinterpret_cast<float4*>(out + ((tok0 + oi) * C + C - 1) * (long long)d + qcol) + oe) = y;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Four electrons, head_dim 64 (Be, LiH): one WARP per (walker, head), channels streamed through a cp.async ring.
// The CTA-per-unit kernel above spends most of its time in ~10 __syncthreads phases with 4..64 active threads; here
// nothing is wider than a warp, there is no CTA barrier, and while a warp works on channel c the loads of channels
// c+1 and c+2 are in flight (16 warps per SM -> ~100 KB in flight per SM, enough to cover HBM latency).
//   score lanes : lane = (hf = lane >> 4, i = (lane >> 2) & 3, j = lane & 3); each lane sums the 16-byte chunks of
//                 parity hf of q_i . k_j and the two halves meet through one shuffle
//   output lanes: lane = (ih = lane >> 4, e4 = lane & 15): rows 2 ih, 2 ih + 1, four columns 4 e4 .. 4 e4 + 3
// Per channel the softmax rule needs only that channel's scores:
//   st = (q_c k_0 + q_0 k_c) scale, dv = st - sum_j p st, pt = p dv, y_c = sum_j pt v_0 + p v_c;
//   accumulated for the Laplacian channel: sum_c q_c.k_c, sum_c dv^2, sum_c sum_j pt v_c.
// Rows are padded to 72 floats so that the eight distinct 16-byte chunks a score load touches fall into eight
// different bank groups.
// ------------------------------------------------------------------------------------------------
constexpr int ATT4_WARPS = 4, ATT4_RING = 3, ATT4_RS = 72, ATT4_CH = 12 * ATT4_RS;   // floats per channel buffer
constexpr int ATT4_SMEM_BYTES = ATT4_WARPS * (1 + ATT4_RING) * ATT4_CH * 4;

// SP (first layer): qkv is the COMPACT payload [token][5][3 d] -- value, the token's own three tangents, Laplacian; the
// rows of the other electrons' tangents are zeros and are zero-filled here without touching memory (src-size 0).
template <bool SP>
__device__ __forceinline__ void att4_issue(float* dst, const float* __restrict__ qkv, long long tok0, int C, int c, int d, int col,
                                           int lane) {
  // 12 rows (part q/k/v x electron) x 16 chunks of 16 bytes
#pragma unroll
  for (int it = 0; it < 6; ++it) {
    const int idx = it * 32 + lane;
    const int r = idx >> 4, e4 = idx & 15;
    const int part = r >> 2, i = r & 3;
    const uint32_t sdst = (uint32_t)__cvta_generic_to_shared(dst + r * ATT4_RS + 4 * e4);
    if constexpr (SP) {
      int cc = 0;
      unsigned sz = 16u;
      if (c == C - 1) cc = 4;
      else if (c > 0) { const int j = (c - 1) / 3; cc = c - 3 * j; sz = (i == j) ? 16u : 0u; }
      const float* src = qkv + ((tok0 + i) * 5 + cc) * (long long)(3 * d) + part * d + col + 4 * e4;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sdst), "l"(src), "r"(sz) : "memory");
    } else {
      const float* src = qkv + ((tok0 + i) * C + c) * (long long)(3 * d) + part * d + col + 4 * e4;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sdst), "l"(src) : "memory");
    }
  }
}

template <bool PK, bool SP = false>      // PK: output written as the packed fp16 pair the tensor-core GEMM consumes (common.cuh)
__global__ void __launch_bounds__(ATT4_WARPS * 32, 4)
attention_payload_n4_kernel(const float* __restrict__ qkv, float* __restrict__ out, long long units, int C, int d, int H,
                            unsigned* ovf) {
  extern __shared__ __align__(16) float sm4[];
  constexpr int N = 4, HD = 64, RS = ATT4_RS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long u = (long long)blockIdx.x * ATT4_WARPS + warp;
  if (u >= units) return;
  const long long b = u / H;
  const int h = (int)(u - b * H);
  const long long tok0 = b * N;
  const int col = h * HD;
  float* buf0 = sm4 + warp * (1 + ATT4_RING) * ATT4_CH;       // value channel: q0 | k0 | v0 rows
  float* ring = buf0 + ATT4_CH;
  const float scale = 0.125f;                                   // 1 / sqrt(64)
  const unsigned FULL = 0xffffffffu;

  att4_issue<SP>(buf0, qkv, tok0, C, 0, d, col, lane);
  cp_async_commit();
  if (C > 1) att4_issue<SP>(ring, qkv, tok0, C, 1, d, col, lane);
  cp_async_commit();
  if (C > 2) att4_issue<SP>(ring + ATT4_CH, qkv, tok0, C, 2, d, col, lane);
  cp_async_commit();
  if (C > 3) att4_issue<SP>(ring + 2 * ATT4_CH, qkv, tok0, C, 3, d, col, lane);
  cp_async_commit();
  cp_async_wait<3>();
  __syncwarp();

  const int hf = lane >> 4, si = (lane >> 2) & 3, sj = lane & 3;       // score lanes
  const int ih = lane >> 4, e4 = lane & 15;                             // output lanes
  const float* q0 = buf0 + si * RS + 4 * hf;
  const float* k0 = buf0 + (4 + sj) * RS + 4 * hf;
  // ---- value channel: p = softmax(q0 k0^T scale), y0 = p v0 ------------------------------------------------
  float p;
  {
    float a = 0.f;
#pragma unroll
    for (int s2 = 0; s2 < 8; ++s2)
      a = dot4(*reinterpret_cast<const float4*>(q0 + 8 * s2), *reinterpret_cast<const float4*>(k0 + 8 * s2), a);
    a += __shfl_xor_sync(FULL, a, 16);
    a *= scale;
    float mx = fmaxf(a, __shfl_xor_sync(FULL, a, 1));
    mx = fmaxf(mx, __shfl_xor_sync(FULL, mx, 2));
    const float e = expf(a - mx);
    float den = e + __shfl_xor_sync(FULL, e, 1);
    den += __shfl_xor_sync(FULL, den, 2);
    p = e * (1.0f / den);
  }
  float pr[2][4];            // p for this lane's two output rows
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int j = 0; j < 4; ++j) pr[r][j] = __shfl_sync(FULL, p, (2 * ih + r) * 4 + j);
  float4 v0[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) v0[j] = *reinterpret_cast<const float4*>(buf0 + (8 + j) * RS + 4 * e4);
  float* orow = out + ((tok0 + 2 * ih) * C) * (long long)d;    // payload row (electron 2 ih, channel 0)
  const int ocol = col + 4 * e4;
  const long long rstep = (long long)C * d;                      // next electron
  float amax = 0.f;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < 4; ++j) axpy4(y, pr[r][j], v0[j]);
    st_row4<PK>(orow + r * rstep, d, ocol, y, amax);
  }
  if (C == 1) { if (PK) raise_range_flag(ovf, amax); return; }
  // ---- tangent channels 1 .. C-2, then the Laplacian channel C-1 ---------------------------------------------
  float cross = 0.f, quad = 0.f;
  float4 cr[2] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};
  for (int c = 1; c < C; ++c) {
    if (c > 1) {
      // slot of channel c - 1 is free now: refill it with channel c + 2 (one commit per iteration, empty or not,
      // keeps the group arithmetic of cp.async.wait_group uniform)
      __syncwarp();
      if (c + 2 < C) att4_issue<SP>(ring + ((c + 1) % ATT4_RING) * ATT4_CH, qkv, tok0, C, c + 2, d, col, lane);
      cp_async_commit();
    }
    cp_async_wait<2>();       // everything but channels c + 1, c + 2 has landed
    __syncwarp();
    const float* cb = ring + ((c - 1) % ATT4_RING) * ATT4_CH;
    const float* qc = cb + si * RS + 4 * hf;
    const float* kc = cb + (4 + sj) * RS + 4 * hf;
    const bool lapc = c == C - 1;
    float a = 0.f, bq = 0.f;
#pragma unroll
    for (int s2 = 0; s2 < 8; ++s2) {
      const float4 qv = *reinterpret_cast<const float4*>(qc + 8 * s2);
      const float4 kv = *reinterpret_cast<const float4*>(kc + 8 * s2);
      a = dot4(qv, *reinterpret_cast<const float4*>(k0 + 8 * s2), a);
      a = dot4(*reinterpret_cast<const float4*>(q0 + 8 * s2), kv, a);
      bq = dot4(qv, kv, bq);
    }
    a += __shfl_xor_sync(FULL, a, 16);
    bq += __shfl_xor_sync(FULL, bq, 16);
    float w;       // weight of v_0 in this channel's output: pt (tangent) or the Laplacian of the softmax
    if (!lapc) {
      const float st = a * scale;
      float m = p * st;
      m += __shfl_xor_sync(FULL, m, 1);
      m += __shfl_xor_sync(FULL, m, 2);
      const float dv = st - m;
      w = p * dv;
      quad = fmaf(dv, dv, quad);
      cross += bq;
    } else {
      const float sl = a * scale + 2.0f * scale * cross;
      float ma = p * sl, mq = p * quad;
      ma += __shfl_xor_sync(FULL, ma, 1); mq += __shfl_xor_sync(FULL, mq, 1);
      ma += __shfl_xor_sync(FULL, ma, 2); mq += __shfl_xor_sync(FULL, mq, 2);
      w = p * ((sl - ma) + quad - mq);
    }
    float4 vc[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) vc[j] = *reinterpret_cast<const float4*>(cb + (8 + j) * RS + 4 * e4);
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      float4 y = lapc ? make_float4(2.0f * cr[r].x, 2.0f * cr[r].y, 2.0f * cr[r].z, 2.0f * cr[r].w) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float wj = __shfl_sync(FULL, w, (2 * ih + r) * 4 + j);
        axpy4(y, wj, v0[j]);
        axpy4(y, pr[r][j], vc[j]);
        if (!lapc) axpy4(cr[r], wj, vc[j]);
      }
      st_row4<PK>(orow + r * rstep + (long long)c * d, d, ocol, y, amax);
    }
  }
  if (PK) raise_range_flag(ovf, amax);
}

// ------------------------------------------------------------------------------------------------
// Tiny heads (head_dim 4, N <= 3: the reference's SMALL preset, He: 16 heads x 4): one THREAD per (walker, head).
// A unit is 3 N x 4 floats per channel; the CTA-per-unit kernel above launches B x H CTAs that are all barrier and
// launch latency (0.23 of the 0.41 ms of a 1024-walker He energy pass).  Here everything lives in registers, channels
// are streamed one after the other, consecutive lanes are consecutive heads (16-byte loads, 256 contiguous bytes per
// row for 16 heads).  Mathematics exactly as in the n4 kernel above (per-channel softmax rule, sums for the Laplacian
// channel accumulated on the way).
// ------------------------------------------------------------------------------------------------
template <int N>
__global__ void __launch_bounds__(128)
attention_payload_hd4_kernel(const float* __restrict__ qkv, float* __restrict__ out, long long units, int C, int d, int H) {
  const long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= units) return;
  const long long b = u / H;
  const int h = (int)(u - b * H);
  const long long d3 = 3ll * d;
  const float scale = 0.5f;                         // 1 / sqrt(4)
  auto ld = [&](int i, int c, int which) {
    return __ldg(reinterpret_cast<const float4*>(qkv + ((b * N + i) * C + c) * d3 + (long long)which * d + 4 * h));
  };
  auto st = [&](int i, int c, const float4 v) {
    *reinterpret_cast<float4*>(out + ((b * N + i) * C + c) * (long long)d + 4 * h) = v;
  };
  auto dot = [](const float4 a, const float4 b4) { return fmaf(a.x, b4.x, fmaf(a.y, b4.y, fmaf(a.z, b4.z, a.w * b4.w))); };
  float4 q0[N], k0[N], v0[N];
#pragma unroll
  for (int i = 0; i < N; ++i) { q0[i] = ld(i, 0, 0); k0[i] = ld(i, 0, 1); v0[i] = ld(i, 0, 2); }
  float p[N][N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < N; ++j) { p[i][j] = dot(q0[i], k0[j]) * scale; mx = fmaxf(mx, p[i][j]); }
    float den = 0.f;
#pragma unroll
    for (int j = 0; j < N; ++j) { p[i][j] = expf(p[i][j] - mx); den += p[i][j]; }
    const float inv = 1.0f / den;
    float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < N; ++j) { p[i][j] *= inv; axpy4(y, p[i][j], v0[j]); }
    st(i, 0, y);
  }
  if (C == 1) return;
  float sL[N][N], quad[N][N];
  float4 yl[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    yl[i] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < N; ++j) { sL[i][j] = 0.f; quad[i][j] = 0.f; }
  }
  for (int c = 1; c < C - 1; ++c) {
    float4 qc[N], kc[N], vc[N];
#pragma unroll
    for (int i = 0; i < N; ++i) { qc[i] = ld(i, c, 0); kc[i] = ld(i, c, 1); vc[i] = ld(i, c, 2); }
#pragma unroll
    for (int i = 0; i < N; ++i) {
      float stv[N], m = 0.f;
#pragma unroll
      for (int j = 0; j < N; ++j) {
        stv[j] = (dot(qc[i], k0[j]) + dot(q0[i], kc[j])) * scale;
        sL[i][j] = fmaf(2.0f * scale, dot(qc[i], kc[j]), sL[i][j]);
        m = fmaf(p[i][j], stv[j], m);
      }
      float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int j = 0; j < N; ++j) {
        const float dv = stv[j] - m;
        const float pt = p[i][j] * dv;
        quad[i][j] = fmaf(dv, dv, quad[i][j]);
        axpy4(y, pt, v0[j]);
        axpy4(y, p[i][j], vc[j]);
        axpy4(yl[i], pt, vc[j]);
      }
      st(i, c, y);
    }
  }
  {
    float4 lq[N], lk[N], lv[N];
#pragma unroll
    for (int i = 0; i < N; ++i) { lq[i] = ld(i, C - 1, 0); lk[i] = ld(i, C - 1, 1); lv[i] = ld(i, C - 1, 2); }
#pragma unroll
    for (int i = 0; i < N; ++i) {
      float a = 0.f, bq = 0.f;
#pragma unroll
      for (int j = 0; j < N; ++j) {
        sL[i][j] += (dot(lq[i], k0[j]) + dot(q0[i], lk[j])) * scale;
        a = fmaf(p[i][j], sL[i][j], a);
        bq = fmaf(p[i][j], quad[i][j], bq);
      }
      float4 y = make_float4(2.0f * yl[i].x, 2.0f * yl[i].y, 2.0f * yl[i].z, 2.0f * yl[i].w);
#pragma unroll
      for (int j = 0; j < N; ++j) {
        const float w = p[i][j] * ((sL[i][j] - a) + quad[i][j] - bq);
        axpy4(y, w, v0[j]);
        axpy4(y, p[i][j], lv[j]);
      }
      st(i, C - 1, y);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// 5 .. 14 electrons, head_dim 64 (Ne, N2, ...): one WARP per (walker, head), channels streamed through a cp.async ring,
// every product register-blocked on the FP32 pipe.  This replaces the CTA-per-unit kernel above for these shapes: that
// one ran at 1.2 TB/s (19 % of HBM; ~10 CTA barriers per channel group, 67 M shared-memory bank conflicts per launch,
// profiles/ncu_attention_v2_ne_r01c_summary.txt) and was 32 % / 39 % of a Ne / N2 energy pass.
//
// The work per channel is two small GEMM groups (same mathematics as the n4 kernel above):
//   scores   a = q_c k_0^T + q_0 k_c^T  and  b += q_c k_c^T        (N x N, reduction over the 64 head columns)
//   outputs  y_c = pt v_0 + p v_c       and  cr += pt v_c          (N x 64, reduction over the N electrons)
// with 6 * 64 * N^2 FMAs against 1 KiB * N of HBM traffic: 0.375 N FMA per byte, i.e. balanced (N = 10) to FMA bound
// (N = 14) on a B200 (tools/pipe_rates.cu: 123 FMA/clk/SM; the legacy HMMA path would need a three-pass fp16 split and
// 16-padding and comes out no faster).  So the kernel is organised around the FMA pipe's issue rate:
//   score lanes   lane = (ig, jg, eh): i-half ig, j-half jg of the N x N pairs (TI = ceil(N/2) rows each: a TI x TI
//                 register tile), eh = one of 8 interleaved slices of the head columns.  All operands of a step are
//                 8-byte shared-memory loads that a whole half-warp shares; 3 TI^2 FMAs per 4 TI loads.
//   hand-over     the 8 partial tiles of an (ig, jg) group are reduce-SCATTERed by recursive halving (7 TI shuffles): lane
//                 eh ends up with row eh of the tile, does that row's softmax-derivative arithmetic and writes p~ to
//                 shared memory; the sum-of-squares and the cross terms for the Laplacian channel stay in registers.
//   output lanes  lane = (ig, eg): TI rows x 4 columns; per electron j two broadcast 16-byte loads of p / p~ columns and
//                 two 16-byte loads of v rows feed 12 TI FMAs.
// No CTA barrier anywhere, no shared-memory write/read ping-pong inside a phase; while a warp computes channel c the
// loads of channels c+1 and c+2 are in flight.  Odd N is handled by zero rows (TI = ceil(N/2)).
// ------------------------------------------------------------------------------------------------
constexpr int ATTW_RS = 68;            // floats per staged row (64 + 4: keeps 16-byte alignment, staggers rows over banks)
constexpr int ATTW_RING = 2;           // channel c + 1 lands while channel c is worked on (a channel is ~2000 issue slots)
__host__ __device__ constexpr int attw_arr_floats(int TI) { return 2 * TI * ATTW_RS; }                       // one of q / k / v
// q0 k0 v0 | ring of (qc kc vc) | p | p~ | mbarriers (base + one per ring slot, 8 bytes each, padded to 32 bytes)
__host__ __device__ constexpr int attw_warp_floats(int TI) { return 3 * (1 + ATTW_RING) * attw_arr_floats(TI) + 2 * 256 + 8; }
__host__ __device__ constexpr int attw_warps(int TI) {
  const int w = (220 * 1024) / (attw_warp_floats(TI) * 4);
  return w > 8 ? 8 : w;
}

// rows [0, N) of q | k | v of channel c -> dst (three arrays of attw_arr_floats): one 256-byte bulk copy per row (the
// head's 64 columns are contiguous in global memory), completion counted on the slot's mbarrier.  16-byte cp.async
// cost ~16 issue slots each here (64-bit address arithmetic plus three dummy LDS that ptxas puts in front of every
// LDGSTS on sm_100a), 10 % of the kernel's instructions; the bulk form is 3 N copies per channel and warp.
__device__ __forceinline__ void attw_issue(float* dst, int arr, uint32_t bar, const float* __restrict__ qkv, long long tok0, int N,
                                           int C, int c, int d, int col, int lane) {
  // ONE elected lane issues all 3 N copies from a loop whose addresses are warp-uniform.  With the copies spread over the
  // lanes (idx = lane, ...) ptxas serialises them anyway -- cp.async.bulk takes its operands from uniform registers, so
  // every divergent issue becomes an ELECT / R2UR.BROADCAST x4 / UBLKCP / BRA.U.ANY round trip per lane: 11 % of the
  // kernel's instructions and 22 % of its stall samples (profiles/ncu_attention_warp_ne_r02j_summary.txt).
  (void)lane;
  __syncwarp();
  uint32_t elected;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(elected));
  if (elected) {
    mbar_arrive_expect_tx(bar, (uint32_t)(3 * N * 256));
    const long long rowstep = (long long)C * 3 * d;                   // floats between the same channel of consecutive electrons
    const float* src0 = qkv + (tok0 * C + c) * (3ll * d) + col;
#pragma unroll
    for (int part = 0; part < 3; ++part) {
      const float* src = src0 + part * d;
      uint32_t sdst = smem_u32(dst + part * arr);
      for (int i = 0; i < N; ++i) {
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 256, [%2];"
                     ::"r"(sdst), "l"(src), "r"(bar) : "memory");
        src += rowstep;
        sdst += ATTW_RS * 4;
      }
    }
  }
}

// MODE 0: a += q0 k0^T.   MODE 1: a += qc k0^T + q0 kc^T, b += qc kc^T.   MODE 2: a += qc k0^T + q0 kc^T.
// q*, k*: this lane's row block and column slice (row stride ATTW_RS); the lane covers columns 2 eh + 16 t, t < 4
template <int TI, int MODE>
__device__ __forceinline__ void attw_scores(const float* q0, const float* k0, const float* qc, const float* kc,
                                            float (&a)[TI][TI], float (&b)[TI][TI]) {
  // NOT unrolled over t: one step is ~6 TI^2 FMAs; the unrolled channel body was 30-55 KiB of code, larger than the
  // instruction cache, and each warp walks through it once per channel
#pragma unroll 1
  for (int t = 0; t < 4; ++t) {
    float2 k0v[TI], kcv[TI];
#pragma unroll
    for (int j = 0; j < TI; ++j) {
      k0v[j] = *reinterpret_cast<const float2*>(k0 + j * ATTW_RS + 16 * t);
      if (MODE != 0) kcv[j] = *reinterpret_cast<const float2*>(kc + j * ATTW_RS + 16 * t);
    }
#pragma unroll
    for (int i = 0; i < TI; ++i) {
      const float2 q0v = *reinterpret_cast<const float2*>(q0 + i * ATTW_RS + 16 * t);
      if (MODE == 0) {
#pragma unroll
        for (int j = 0; j < TI; ++j) a[i][j] = fmaf(q0v.y, k0v[j].y, fmaf(q0v.x, k0v[j].x, a[i][j]));
      } else {
        const float2 qcv = *reinterpret_cast<const float2*>(qc + i * ATTW_RS + 16 * t);
#pragma unroll
        for (int j = 0; j < TI; ++j) {
          a[i][j] = fmaf(qcv.y, k0v[j].y, fmaf(qcv.x, k0v[j].x, a[i][j]));
          a[i][j] = fmaf(q0v.y, kcv[j].y, fmaf(q0v.x, kcv[j].x, a[i][j]));
          if (MODE == 1) b[i][j] = fmaf(qcv.y, kcv[j].y, fmaf(qcv.x, kcv[j].x, b[i][j]));
        }
      }
    }
  }
}

// Sum the TI x TI tiles of the 8 lanes that share (ig, jg) (lane bits 0..2) and leave row eh in lane eh (rows >= TI: zero)
template <int TI>
__device__ __forceinline__ void attw_reduce_scatter(const float (&a)[TI][TI], int eh, float (&row)[TI]) {
  const unsigned FULL = 0xffffffffu;
  const bool b2 = (eh & 4) != 0, b1 = (eh & 2) != 0, b0 = (eh & 1) != 0;
  float h4[4][TI];      // rows (eh & 4) + r
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int j = 0; j < TI; ++j) {
      const float lo = r < TI ? a[r < TI ? r : 0][j] : 0.f;
      const float hi = r + 4 < TI ? a[r + 4 < TI ? r + 4 : 0][j] : 0.f;
      const float got = __shfl_xor_sync(FULL, b2 ? lo : hi, 4);
      h4[r][j] = (b2 ? hi : lo) + got;
    }
  float h2[2][TI];      // rows (eh & 6) + r
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int j = 0; j < TI; ++j) {
      const float got = __shfl_xor_sync(FULL, b1 ? h4[r][j] : h4[r + 2][j], 2);
      h2[r][j] = (b1 ? h4[r + 2][j] : h4[r][j]) + got;
    }
#pragma unroll
  for (int j = 0; j < TI; ++j) {
    const float got = __shfl_xor_sync(FULL, b0 ? h2[0][j] : h2[1][j], 1);
    row[j] = (b0 ? h2[1][j] : h2[0][j]) + got;
  }
}

template <int TI, bool PK>
__global__ void __launch_bounds__(attw_warps(TI) * 32, 1)
attention_payload_warp_kernel(const float* __restrict__ qkv, float* __restrict__ out, long long units, int N, int C, int d, int H,
                              unsigned* ovf) {
  extern __shared__ __align__(16) float smw[];
  constexpr int RS = ATTW_RS, ARR = attw_arr_floats(TI), TRI = 3 * ARR;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long u = (long long)blockIdx.x * attw_warps(TI) + warp;
  if (u >= units) return;
  const long long b = u / H;
  const int h = (int)(u - b * H);
  const long long tok0 = b * N;
  const int col = h * 64;
  float* base0 = smw + (size_t)warp * attw_warp_floats(TI);    // q0 | k0 | v0
  float* ring = base0 + TRI;
  float* Pm = ring + ATTW_RING * TRI;                           // p  [j][16]: column j of p, rows ig * 8 + r
  float* PTm = Pm + 256;                                        // p~ (tangent) / softmax Laplacian weights, same layout
  const uint32_t bar0 = smem_u32(PTm + 256);                    // mbarriers: base, ring slot 0, ring slot 1
  const float scale = 0.125f;                                   // 1 / sqrt(64)
  const unsigned FULL = 0xffffffffu;
  // rows N .. 2 TI - 1 (an odd N) and the p / p~ matrices start as zeros: padding contributes exact zeros.  The bulk
  // copies never write these locations, so the two proxies do not meet.
  if (N < 2 * TI)
    for (int i = lane; i < 3 * (1 + ATTW_RING) * RS; i += 32) base0[(i / RS) * ARR + N * RS + (i % RS)] = 0.f;
  for (int i = lane; i < 512; i += 32) Pm[i] = 0.f;
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < 1 + ATTW_RING; ++s) mbar_init(bar0 + 8u * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncwarp();

  attw_issue(base0, ARR, bar0, qkv, tok0, N, C, 0, d, col, lane);
#pragma unroll
  for (int s = 0; s < ATTW_RING; ++s)
    if (1 + s < C) attw_issue(ring + s * TRI, ARR, bar0 + 8u * (1 + s), qkv, tok0, N, C, 1 + s, d, col, lane);
  mbar_wait_warp(bar0, 0);

  // score lanes
  const int eh = lane & 7, jg = (lane >> 3) & 1, ig = lane >> 4;
  const int qoff = ig * TI * RS + 2 * eh, koff = ARR + jg * TI * RS + 2 * eh;
  // output lanes
  const int eg = lane & 15;
  const int voff = 2 * ARR + 4 * eg;
  float* orow0 = out + ((tok0 + ig * TI) * C) * (long long)d;   // payload row (electron ig TI, channel 0)
  const long long rstep = (long long)C * d;
  const int ocol = col + 4 * eg;
  float amax = 0.f;

  float p[TI], quad[TI];          // row `myrow` of p and of sum_c dv^2, columns jg TI + j
  float bacc[TI][TI];             // cross terms sum_c q_c k_c^T (partial over this lane's column slice)
  float4 cr[TI];                  // sum_c p~_c v_c (output-lane layout)
#pragma unroll
  for (int i = 0; i < TI; ++i) {
    quad[i] = 0.f;
    cr[i] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < TI; ++j) bacc[i][j] = 0.f;
  }
  // ---- value channel: p = softmax(q0 k0^T scale), y0 = p v0 ---------------------------------------------------------
  {
    float a[TI][TI];
#pragma unroll
    for (int i = 0; i < TI; ++i)
#pragma unroll
      for (int j = 0; j < TI; ++j) a[i][j] = 0.f;
    attw_scores<TI, 0>(base0 + qoff, base0 + koff, nullptr, nullptr, a, bacc);
    float s[TI];
    attw_reduce_scatter<TI>(a, eh, s);
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < TI; ++j) {
      s[j] = (jg * TI + j < N) ? s[j] * scale : -INFINITY;      // padding columns take no weight
      mx = fmaxf(mx, s[j]);
    }
    mx = fmaxf(mx, __shfl_xor_sync(FULL, mx, 8));
    float den = 0.f;
#pragma unroll
    for (int j = 0; j < TI; ++j) {
      s[j] = expf(s[j] - mx);
      den += s[j];
    }
    den += __shfl_xor_sync(FULL, den, 8);
    const float inv = 1.0f / den;
#pragma unroll
    for (int j = 0; j < TI; ++j) {
      p[j] = s[j] * inv;
      if (eh < TI) Pm[(jg * TI + j) * 16 + ig * 8 + eh] = p[j];
    }
  }
  __syncwarp();
  {
    float4 y[TI];
#pragma unroll
    for (int i = 0; i < TI; ++i) y[i] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
    for (int j = 0; j < 2 * TI; ++j) {
      const float4 v0j = *reinterpret_cast<const float4*>(base0 + voff + j * RS);
      const float4 pa = *reinterpret_cast<const float4*>(Pm + j * 16 + ig * 8);
      const float4 pb = *reinterpret_cast<const float4*>(Pm + j * 16 + ig * 8 + 4);
      const float pv[8] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w};
#pragma unroll
      for (int i = 0; i < TI; ++i) axpy4(y[i], pv[i], v0j);
    }
#pragma unroll
    for (int i = 0; i < TI; ++i)
      if (ig * TI + i < N) st_row4<PK>(orow0 + i * rstep, d, ocol, y[i], amax);
  }
  if (C == 1) { if (PK) raise_range_flag(ovf, amax); return; }

  // ---- tangent channels 1 .. C-2, then the Laplacian channel C-1 ------------------------------------------------------
  for (int c = 1; c < C; ++c) {
    const int slot = (c - 1) % ATTW_RING;
    if (c > 1 && c - 1 + ATTW_RING < C) {
      // the slot of channel c - 1 is free (every lane is past its last read of it): refill it
      __syncwarp();
      const int fs = (c - 2) % ATTW_RING;
      attw_issue(ring + fs * TRI, ARR, bar0 + 8u * (1 + fs), qkv, tok0, N, C, c - 1 + ATTW_RING, d, col, lane);
    }
    mbar_wait_warp(bar0 + 8u * (1 + slot), (uint32_t)(((c - 1) / ATTW_RING) & 1));
    const float* cb = ring + slot * TRI;
    const bool lapc = c == C - 1;
    float a[TI][TI];
#pragma unroll
    for (int i = 0; i < TI; ++i)
#pragma unroll
      for (int j = 0; j < TI; ++j) a[i][j] = 0.f;
    if (!lapc) attw_scores<TI, 1>(base0 + qoff, base0 + koff, cb + qoff, cb + koff, a, bacc);
    else attw_scores<TI, 2>(base0 + qoff, base0 + koff, cb + qoff, cb + koff, a, bacc);
    float st[TI];
    attw_reduce_scatter<TI>(a, eh, st);
    float w[TI];                    // weight of v_0 in this channel's output: p~ (tangent) or the softmax Laplacian
    if (!lapc) {
      float m = 0.f;
#pragma unroll
      for (int j = 0; j < TI; ++j) {
        st[j] *= scale;
        m = fmaf(p[j], st[j], m);
      }
      m += __shfl_xor_sync(FULL, m, 8);
#pragma unroll
      for (int j = 0; j < TI; ++j) {
        const float dv = st[j] - m;
        w[j] = p[j] * dv;
        quad[j] = fmaf(dv, dv, quad[j]);
      }
    } else {
      float cross[TI];
      attw_reduce_scatter<TI>(bacc, eh, cross);
      float ma = 0.f, mq = 0.f;
#pragma unroll
      for (int j = 0; j < TI; ++j) {
        st[j] = st[j] * scale + 2.0f * scale * cross[j];
        ma = fmaf(p[j], st[j], ma);
        mq = fmaf(p[j], quad[j], mq);
      }
      ma += __shfl_xor_sync(FULL, ma, 8);
      mq += __shfl_xor_sync(FULL, mq, 8);
#pragma unroll
      for (int j = 0; j < TI; ++j) w[j] = p[j] * ((st[j] - ma) + quad[j] - mq);
    }
#pragma unroll
    for (int j = 0; j < TI; ++j)
      if (eh < TI) PTm[(jg * TI + j) * 16 + ig * 8 + eh] = w[j];
    __syncwarp();
    float4 y[TI];
#pragma unroll
    for (int i = 0; i < TI; ++i)
      y[i] = lapc ? make_float4(2.0f * cr[i].x, 2.0f * cr[i].y, 2.0f * cr[i].z, 2.0f * cr[i].w) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
    for (int j = 0; j < 2 * TI; ++j) {
      const float4 v0j = *reinterpret_cast<const float4*>(base0 + voff + j * RS);
      const float4 vcj = *reinterpret_cast<const float4*>(cb + voff + j * RS);
      const float4 pa = *reinterpret_cast<const float4*>(Pm + j * 16 + ig * 8);
      const float4 pb = *reinterpret_cast<const float4*>(Pm + j * 16 + ig * 8 + 4);
      const float4 wa = *reinterpret_cast<const float4*>(PTm + j * 16 + ig * 8);
      const float4 wb = *reinterpret_cast<const float4*>(PTm + j * 16 + ig * 8 + 4);
      const float pv[8] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w};
      const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
      // three sweeps over the rows: the two updates of y[i] are then 4 TI instructions apart instead of back to back
#pragma unroll
      for (int i = 0; i < TI; ++i) axpy4(y[i], wv[i], v0j);
      if (!lapc) {
#pragma unroll
        for (int i = 0; i < TI; ++i) axpy4(cr[i], wv[i], vcj);
      }
#pragma unroll
      for (int i = 0; i < TI; ++i) axpy4(y[i], pv[i], vcj);
    }
#pragma unroll
    for (int i = 0; i < TI; ++i)
      if (ig * TI + i < N) st_row4<PK>(orow0 + i * rstep + (long long)c * d, d, ocol, y[i], amax);
  }
  if (PK) raise_range_flag(ovf, amax);
}

// shapes for which the output can be written as the packed fp16 pair
inline bool attention_warp_shape(int N, int d, int H) { return H > 0 && d % H == 0 && d / H == 64 && N >= 5 && N <= 14; }
inline bool attention_can_pack(int N, int d, int H) {
  return H > 0 && d % H == 0 && d / H == 64 && (N == 4 || attention_warp_shape(N, d, H));
}

// two warps per unit (attention_pair.cuh); returns PSIF_OK after launching
inline int32_t attention_pair_launch(const float* qkv, float* out, long long units, int N, int C, int d, int H, cudaStream_t st, bool packed,
                              unsigned* ovf);
// PSIF_ATT_PAIR=0 keeps the one-warp-per-unit kernel (A/B timing); read once per process, never changes afterwards
inline bool attention_use_pair() {
  static const bool on = [] { const char* e = getenv("PSIF_ATT_PAIR"); return e == nullptr || e[0] != '0'; }();
  return on;
}

// first-layer mode (attention_first_layer_sparse): qkv is the compact payload [token][5][3 d]
inline bool attention_first_layer_sparse(int N, int d, int H) { return N == 4 && H > 0 && d % H == 0 && d / H == 64; }

inline int32_t attention_payload(const float* qkv, float* out, long long B, int N, int C, int d, int H,
                                 cudaStream_t st, bool packed = false, unsigned* ovf = nullptr, bool l0_sparse = false) {
  if (B <= 0) return PSIF_OK;
  if (l0_sparse && !(attention_first_layer_sparse(N, d, H) && C == 3 * N + 2 && ((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(out)) & 15) == 0))
    return fail(PSIF_E_INVALID, "attention: first-layer mode not available for this shape%s");
  if (packed && !attention_can_pack(N, d, H)) return fail(PSIF_E_INVALID, "attention: packed output not available for this shape%s");
  if (H <= 0 || d % H != 0) return fail(PSIF_E_INVALID, "attention: n_embd must be divisible by n_head%s");
  const int hd = d / H;
  if (N > PSIF_MAX_ELEC || hd > 128) return fail(PSIF_E_INVALID, "attention: N > 16 or head_dim > 128 unsupported%s");
  const long long grid2 = B * H;
  DevSmemCfg& cfg = dev_smem_cfg();
  if (hd == 4 && (N == 2 || N == 3) && (reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
    const long long nb = (grid2 + 127) / 128;
    if (nb > 0x7fffffffLL) return fail(PSIF_E_INVALID, "attention: grid too large%s");
    if (N == 2) PSIF_LAUNCH(attention_payload_hd4_kernel<2>, (unsigned)nb, 128, 0, st, qkv, out, grid2, C, d, H);
    else PSIF_LAUNCH(attention_payload_hd4_kernel<3>, (unsigned)nb, 128, 0, st, qkv, out, grid2, C, d, H);
    return PSIF_OK;
  }
  if (N == 4 && hd == 64 && d % 4 == 0 && (reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
    if (!cfg.att4) {
      PSIF_CUDA_CHECK(cudaFuncSetAttribute(attention_payload_n4_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT4_SMEM_BYTES));
      PSIF_CUDA_CHECK(cudaFuncSetAttribute(attention_payload_n4_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT4_SMEM_BYTES));
      PSIF_CUDA_CHECK(cudaFuncSetAttribute((attention_payload_n4_kernel<false, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, ATT4_SMEM_BYTES));
      PSIF_CUDA_CHECK(cudaFuncSetAttribute((attention_payload_n4_kernel<true, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, ATT4_SMEM_BYTES));
      cfg.att4 = true;
    }
    const long long nb = (grid2 + ATT4_WARPS - 1) / ATT4_WARPS;
    if (nb > 0x7fffffffLL) return fail(PSIF_E_INVALID, "attention: grid too large%s");
    if (l0_sparse) {
      if (packed) PSIF_LAUNCH((attention_payload_n4_kernel<true, true>), (unsigned)nb, ATT4_WARPS * 32, ATT4_SMEM_BYTES, st, qkv, out, grid2, C, d, H, ovf);
      else PSIF_LAUNCH((attention_payload_n4_kernel<false, true>), (unsigned)nb, ATT4_WARPS * 32, ATT4_SMEM_BYTES, st, qkv, out, grid2, C, d, H, ovf);
      return PSIF_OK;
    }
    if (packed) PSIF_LAUNCH(attention_payload_n4_kernel<true>, (unsigned)nb, ATT4_WARPS * 32, ATT4_SMEM_BYTES, st, qkv, out, grid2, C, d, H, ovf);
    else PSIF_LAUNCH(attention_payload_n4_kernel<false>, (unsigned)nb, ATT4_WARPS * 32, ATT4_SMEM_BYTES, st, qkv, out, grid2, C, d, H, ovf);
    return PSIF_OK;
  }
  if (attention_warp_shape(N, d, H) && (reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
    if (attention_use_pair()) return attention_pair_launch(qkv, out, grid2, N, C, d, H, st, packed, ovf);
    const int TI = (N + 1) / 2;
#define PSIF_ATTW(T)                                                                                                              \
  case T: {                                                                                                                       \
    constexpr int W = attw_warps(T);                                                                                              \
    constexpr size_t smem = (size_t)W * attw_warp_floats(T) * sizeof(float);                                                      \
    if (cfg.attw[T] == 0) {                                                                                                       \
      PSIF_CUDA_CHECK(cudaFuncSetAttribute((attention_payload_warp_kernel<T, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      PSIF_CUDA_CHECK(cudaFuncSetAttribute((attention_payload_warp_kernel<T, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));  \
      cfg.attw[T] = smem;                                                                                                         \
    }                                                                                                                             \
    const long long nb = (grid2 + W - 1) / W;                                                                                     \
    if (nb > 0x7fffffffLL) return fail(PSIF_E_INVALID, "attention: grid too large%s");                                            \
    if (packed) PSIF_LAUNCH((attention_payload_warp_kernel<T, true>), (unsigned)nb, W * 32, smem, st, qkv, out, grid2, N, C, d, H, ovf);     \
    else PSIF_LAUNCH((attention_payload_warp_kernel<T, false>), (unsigned)nb, W * 32, smem, st, qkv, out, grid2, N, C, d, H, ovf);           \
    return PSIF_OK;                                                                                                               \
  }
    switch (TI) { PSIF_ATTW(3) PSIF_ATTW(4) PSIF_ATTW(5) PSIF_ATTW(6) PSIF_ATTW(7) }
#undef PSIF_ATTW
  }
  if (hd % 4 == 0 && N * (hd / 4) <= ATT2_THREADS && d % 4 == 0 && grid2 <= 0x7fffffffLL &&
      (reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
    const Att2Smem L2 = att2_layout(N, hd, C);
    const size_t smem2 = (size_t)L2.total * sizeof(float);
    if (smem2 <= 200 * 1024) {
      if (smem2 > 48 * 1024 && smem2 > cfg.att2) {
        PSIF_CUDA_CHECK(cudaFuncSetAttribute(attention_payload_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
        cfg.att2 = smem2;
      }
      PSIF_LAUNCH(attention_payload_v2_kernel, (unsigned)grid2, ATT2_THREADS, smem2, st, qkv, out, N, C, d, H);
      return PSIF_OK;
    }
  }
  const AttSmem L = att_layout(N, hd, C);
  const size_t smem = (size_t)L.total * sizeof(float);
  if (smem > 220 * 1024) return fail(PSIF_E_INVALID, "attention: shared memory budget exceeded%s");
  if (smem > 48 * 1024 && smem > cfg.att1) {
    PSIF_CUDA_CHECK(cudaFuncSetAttribute(attention_payload_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cfg.att1 = smem;
  }
  const long long grid = B * H;
  if (grid > 0x7fffffffLL) return fail(PSIF_E_INVALID, "attention: grid too large%s");
  PSIF_LAUNCH(attention_payload_kernel, (unsigned)grid, ATT_THREADS, smem, st, qkv, out, N, C, d, H);
  return PSIF_OK;
}

}  // namespace psif
