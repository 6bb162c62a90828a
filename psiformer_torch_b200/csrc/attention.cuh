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

inline int32_t attention_payload(const float* qkv, float* out, long long B, int N, int C, int d, int H,
                                 cudaStream_t st) {
  if (B <= 0) return PSIF_OK;
  if (H <= 0 || d % H != 0) return fail(PSIF_E_INVALID, "attention: n_embd must be divisible by n_head%s");
  const int hd = d / H;
  if (N > PSIF_MAX_ELEC || hd > 128) return fail(PSIF_E_INVALID, "attention: N > 16 or head_dim > 128 unsupported%s");
  const long long grid2 = B * H;
  if (hd % 4 == 0 && N * (hd / 4) <= ATT2_THREADS && d % 4 == 0 && grid2 <= 0x7fffffffLL &&
      (reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
    const Att2Smem L2 = att2_layout(N, hd, C);
    const size_t smem2 = (size_t)L2.total * sizeof(float);
    if (smem2 <= 200 * 1024) {
      static size_t configured2 = 0;
      if (smem2 > 48 * 1024 && smem2 > configured2) {
        PSIF_CUDA_CHECK(cudaFuncSetAttribute(attention_payload_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
        configured2 = smem2;
      }
      PSIF_LAUNCH(attention_payload_v2_kernel, (unsigned)grid2, ATT2_THREADS, smem2, st, qkv, out, N, C, d, H);
      return PSIF_OK;
    }
  }
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
