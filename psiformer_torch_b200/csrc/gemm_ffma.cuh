// FP32 FFMA GEMM for the Linear layers on payload rows:
//     Y[M][N] = X[M][K] * W[N][K]^T  (+ bias[n] on rows r with r % C == 0)  (+ residual)  (GELU if C == 1)
// W is an nn.Linear weight as stored in the state_dict ([out][in], psiformer.py:39,63,74,76,143-144).
// This kernel is exact-fp32 and shape-generic (any M, N, K; it is what the DEBUG preset with
// d = 4 and the K = 4*natom embedding use).  The tcgen05 3xTF32 kernel (gemm_tcgen05.cuh)
// takes over for the large aligned shapes.
#pragma once
#include "common.cuh"

namespace psif {

constexpr int GF_BM = 128, GF_BN = 128, GF_BK = 8, GF_THREADS = 256;

template <bool VEC>
__device__ __forceinline__ float4 gf_load4(const float* __restrict__ base, long long row, long long nrows,
                                           int k, int K, long long ld) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (row < nrows) {
    const float* p = base + row * ld + k;
    if (VEC) {
      if (k < K) v = __ldg(reinterpret_cast<const float4*>(p));
    } else {
      if (k + 0 < K) v.x = __ldg(p + 0);
      if (k + 1 < K) v.y = __ldg(p + 1);
      if (k + 2 < K) v.z = __ldg(p + 2);
      if (k + 3 < K) v.w = __ldg(p + 3);
    }
  }
  return v;
}

template <bool VEC>
__global__ void __launch_bounds__(GF_THREADS)
gemm_tn_ffma_kernel(const float* __restrict__ X, const float* __restrict__ W, const float* __restrict__ bias,
                    const float* res, float* Y, long long M, int N, int K, int C, int act, int tiles_n) {
  __shared__ __align__(16) float As[2][GF_BK][GF_BM];
  __shared__ __align__(16) float Bs[2][GF_BK][GF_BN];

  const int tid = threadIdx.x;
  const long long bid = blockIdx.x;
  const long long m0 = (bid / tiles_n) * GF_BM;
  const int n0 = (int)(bid % tiles_n) * GF_BN;
  const int tx = tid & 15, ty = tid >> 4;
  const int lrow = tid >> 1, lk = (tid & 1) * 4;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  const int nk = (K + GF_BK - 1) / GF_BK;
  float4 ra = gf_load4<VEC>(X, m0 + lrow, M, lk, K, K);
  float4 rb = gf_load4<VEC>(W, (long long)n0 + lrow, N, lk, K, K);
  As[0][lk + 0][lrow] = ra.x; As[0][lk + 1][lrow] = ra.y; As[0][lk + 2][lrow] = ra.z; As[0][lk + 3][lrow] = ra.w;
  Bs[0][lk + 0][lrow] = rb.x; Bs[0][lk + 1][lrow] = rb.y; Bs[0][lk + 2][lrow] = rb.z; Bs[0][lk + 3][lrow] = rb.w;
  __syncthreads();

  int cur = 0;
  for (int kt = 0; kt < nk; ++kt) {
    if (kt + 1 < nk) {
      const int k = (kt + 1) * GF_BK + lk;
      ra = gf_load4<VEC>(X, m0 + lrow, M, k, K, K);
      rb = gf_load4<VEC>(W, (long long)n0 + lrow, N, k, K, K);
    }
#pragma unroll
    for (int k = 0; k < GF_BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[cur][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[cur][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[cur][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[cur][k][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      const int nx = cur ^ 1;
      As[nx][lk + 0][lrow] = ra.x; As[nx][lk + 1][lrow] = ra.y; As[nx][lk + 2][lrow] = ra.z; As[nx][lk + 3][lrow] = ra.w;
      Bs[nx][lk + 0][lrow] = rb.x; Bs[nx][lk + 1][lrow] = rb.y; Bs[nx][lk + 2][lrow] = rb.z; Bs[nx][lk + 3][lrow] = rb.w;
    }
    __syncthreads();
    cur ^= 1;
  }

  // epilogue
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long r = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (r >= M) continue;
    const bool with_bias = (bias != nullptr) && (C == 1 || (r % C) == 0);
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      const int cbase = n0 + jh * 64 + tx * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = cbase + j;
        if (c >= N) continue;
        float v = acc[i][jh * 4 + j];
        if (with_bias) v += __ldg(bias + c);
        if (act) v = gelu_tanh(v);
        if (res != nullptr) v += res[r * (long long)N + c];
        Y[r * (long long)N + c] = v;
      }
    }
  }
}

// host launcher
inline int32_t gemm_ffma(const float* X, const float* W, const float* bias, const float* res, float* Y,
                         long long M, int N, int K, int C, int act, cudaStream_t st) {
  if (M <= 0 || N <= 0 || K <= 0) return PSIF_OK;
  const int tiles_n = cdiv(N, GF_BN);
  const long long tiles = (long long)cdiv(M, GF_BM) * tiles_n;
  if (tiles > 0x7fffffffLL) return fail(PSIF_E_INVALID, "gemm: too many tiles%s");
  const bool vec = (K % 4 == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(W) & 15) == 0);
  if (vec)
    PSIF_LAUNCH(gemm_tn_ffma_kernel<true>, (unsigned)tiles, GF_THREADS, 0, st, X, W, bias, res, Y, M, N, K, C, act, tiles_n);
  else
    PSIF_LAUNCH(gemm_tn_ffma_kernel<false>, (unsigned)tiles, GF_THREADS, 0, st, X, W, bias, res, Y, M, N, K, C, act, tiles_n);
  return PSIF_OK;
}

}  // namespace psif
