// Attention on the payload of the FIRST layer (psiformer.py:42-62 under the forward Laplacian of DESIGN.md section 3).
//
// In front of the first attention token i depends on x_i only, so q_i, k_i, v_i carry five non-zero payload rows: value,
// the three tangents d/dx_{i,alpha} and the Laplacian.  The QKV GEMM of layer 0 therefore runs on a COMPACT payload
// [token][5][3 d] (psif_api.cu), and this kernel turns it into the dense attention output [token][C][d], C = 3 N + 2.
// With S = scale q0 k0^T, P = softmax(S), o_i = sum_k P_ik v0_k, and for channel c = (j, alpha)
//     A_c[k] = scale dq_j[alpha] . k0_k          (row j of dS)          T_c[i] = scale q0_i . dk_j[alpha]   (column j of dS)
// the dense rules of attention.cuh collapse to
//     i != j :  dS_ik = delta_kj T_c[i]                      o_i[c] = P_ij (T_c[i] (v0_j - o_i) + dv_j[alpha])
//     i == j :  dS_jk = A_c[k] + delta_kj T_c[j],  m = sum_k P_jk dS_jk,  dP_jk = P_jk (dS_jk - m)
//                                                             o_j[c] = sum_k dP_jk v0_k + P_jj dv_j[alpha]
//     Laplacian:  LS_ik = scale (lq_i . k0_k + q0_i . lk_k) + delta_ik 2 scale sum_alpha dq_i[alpha] . dk_i[alpha]
//                 quad_ik = sum_c (dS_ik[c] - m_i[c])^2
//                 dLP_ik = P_ik ((LS_ik - sum_m P_im LS_im) + quad_ik - sum_m P_im quad_im)
//                 o_i[L] = sum_k (dLP_ik v0_k + P_ik lv_k) + 2 sum_c dP_{i j(c)}[c] dv_{j(c)}[alpha(c)]
// (restated in fp64 as oracle/forward_laplacian.py::attention_first_layer_payload and checked there against the dense rule)
// i.e. O(N^2) dot products and O(N^2) 64-wide vector updates per (walker, head) instead of O(N^3): the kernel is bound by
// writing its output.  One CTA works through (walker, head) units: inputs to shared memory, the dot products in 4 x 4 register
// blocks, the scalar coefficient tables, then 16 lanes per output row.
#pragma once
#include "common.cuh"

namespace psif {

constexpr int AFL_THREADS = 256, AFL_RS = 68, AFL_HD = 64;

struct AflLayout {
  int vec, P, A, T, LS, LS2, TT, quad, WL, Wown, CX, m, X, O0, total;     // float offsets
};
__host__ __device__ inline AflLayout afl_layout(int N) {
  AflLayout L;
  int o = 0;
  L.vec = o;  o += N * 15 * AFL_RS;        // [i][row 0..4][part q k v][68]
  L.O0 = o;   o += N * AFL_RS;             // value output rows
  L.P = o;    o += N * N;
  L.A = o;    o += 3 * N * N;              // [j][alpha][k]
  L.T = o;    o += 3 * N * N;              // [j][alpha][i]
  L.LS = o;   o += N * N;              // scale lq_i . k0_k  (+ the cross term on the diagonal)
  L.LS2 = o;  o += N * N;              // scale q0_i . lk_k
  L.TT = o;   o += N * N;              // [j][i]: sum_alpha T[j][alpha][i]^2
  L.quad = o; o += N * N;
  L.WL = o;   o += N * N;                  // coefficient of v0_k in the Laplacian row of i
  L.Wown = o; o += 3 * N * N;              // [j][alpha][k]: dP_jk of the own-electron tangent rows
  L.CX = o;   o += 3 * N * N;              // [i][j][alpha]: 2 dP_ij[(j, alpha)]
  L.m = o;    o += 3 * N;
  L.X = o;    o += N;
  L.total = (o + 3) & ~3;
  return L;
}

__device__ __forceinline__ float afl_dot64(const float* a, const float* b) {
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int e = 0; e < AFL_HD; e += 8) {
    const float4 x0 = *reinterpret_cast<const float4*>(a + e), y0 = *reinterpret_cast<const float4*>(b + e);
    const float4 x1 = *reinterpret_cast<const float4*>(a + e + 4), y1 = *reinterpret_cast<const float4*>(b + e + 4);
    s0 = fmaf(x0.x, y0.x, s0); s0 = fmaf(x0.y, y0.y, s0); s0 = fmaf(x0.z, y0.z, s0); s0 = fmaf(x0.w, y0.w, s0);
    s1 = fmaf(x1.x, y1.x, s1); s1 = fmaf(x1.y, y1.y, s1); s1 = fmaf(x1.z, y1.z, s1); s1 = fmaf(x1.w, y1.w, s1);
  }
  return s0 + s1;
}

__device__ __forceinline__ void afl_axpy(float4& y, float a, const float* v) {
  const float4 x = *reinterpret_cast<const float4*>(v);
  y.x = fmaf(a, x.x, y.x); y.y = fmaf(a, x.y, y.y); y.z = fmaf(a, x.z, y.z); y.w = fmaf(a, x.w, y.w);
}

// sum / max over the 16 lanes of a group (both halves of the warp take part)
__device__ __forceinline__ float afl_gsum(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o, 16);
  return v;
}
__device__ __forceinline__ float afl_gmax(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o, 16));
  return v;
}

template <bool PK>
__global__ void __launch_bounds__(AFL_THREADS)
attention_first_layer_kernel(const float* __restrict__ qkv, float* __restrict__ out, long long units, int N, int C, int d, int H,
                             unsigned* ovf) {
  extern __shared__ __align__(16) float afl_sm[];
  const AflLayout L = afl_layout(N);
  float* vec = afl_sm + L.vec;
  float* O0 = afl_sm + L.O0;
  float* P = afl_sm + L.P;
  float* A = afl_sm + L.A;
  float* T = afl_sm + L.T;
  float* LS = afl_sm + L.LS;
  float* LS2 = afl_sm + L.LS2;
  float* TT = afl_sm + L.TT;
  float* quad = afl_sm + L.quad;
  float* WL = afl_sm + L.WL;
  float* Wown = afl_sm + L.Wown;
  float* CX = afl_sm + L.CX;
  float* mown = afl_sm + L.m;
  float* X = afl_sm + L.X;
  const int tid = threadIdx.x;
  const float scale = 0.125f;        // 1 / sqrt(64)
  const int N3 = 3 * N;
  float amax = 0.f;
#define AFL_VEC(i, r, part) (vec + (((i) * 5 + (r)) * 3 + (part)) * AFL_RS)

  for (long long u = blockIdx.x; u < units; u += gridDim.x) {
    const long long b = u / H;
    const int h = (int)(u - b * H);
    const long long tok0 = b * N;
    const int col = h * AFL_HD;
    // ---- inputs: N tokens x 5 rows x 3 parts x 64 floats ----
    for (int idx = tid; idx < N * 15 * 16; idx += AFL_THREADS) {
      const int e4 = idx & 15, v = idx >> 4;            // v = (i * 5 + r) * 3 + part
      const int part = v % 3, ir = v / 3;
      const float* src = qkv + (tok0 * 5 + ir) * (long long)(3 * d) + part * d + col + 4 * e4;
      const uint32_t sdst = (uint32_t)__cvta_generic_to_shared(vec + v * AFL_RS + 4 * e4);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sdst), "l"(src) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    // ---- dot products in 4 x 4 register blocks (8 shared-memory loads feed 64 FMAs) ----
    //   set 1: all 5 N query-side vectors (q0, dq, lq) x k0      -> S, A, scale lq . k0
    //   set 2: q0 x the 4 N key-side vectors dk, lk              -> T, scale q0 . lk
    {
      const int nq1 = (5 * N + 3) >> 2, nk1 = (N + 3) >> 2, nb1 = nq1 * nk1;
      const int nq2 = (N + 3) >> 2, nk2 = N, nb2 = nq2 * nk2;          // 4 N key-side vectors = N blocks of 4
      for (int blk = tid; blk < nb1 + nb2 + N; blk += AFL_THREADS) {
        if (blk >= nb1 + nb2) {                                        // X[j]: the cross term of the diagonal
          const int j = blk - nb1 - nb2;
          float s = 0.f;
          for (int al = 0; al < 3; ++al) s += afl_dot64(AFL_VEC(j, 1 + al, 0), AFL_VEC(j, 1 + al, 1));
          X[j] = 2.0f * scale * s;
          continue;
        }
        const bool s1 = blk < nb1;
        const int bb = s1 ? blk : blk - nb1;
        const int bq = s1 ? bb / nk1 : bb / nk2, bk = s1 ? bb - bq * nk1 : bb - bq * nk2;
        const int nqv = s1 ? 5 * N : N, nkv = s1 ? N : 4 * N;
        const float* qp[4];
        const float* kp[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          int qi = 4 * bq + a; qi = qi < nqv ? qi : nqv - 1;           // query-side index = r * N + i
          int ki = 4 * bk + a; ki = ki < nkv ? ki : nkv - 1;           // key-side index (set 2: (r - 1) * N + k, r = 1 .. 4)
          const int qr = qi / N, qe = qi - qr * N;
          const int kr = s1 ? 0 : 1 + ki / N, ke = s1 ? ki : ki - (kr - 1) * N;
          qp[a] = AFL_VEC(qe, qr, 0);
          kp[a] = AFL_VEC(ke, kr, 1);
        }
        float acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[a][c] = 0.f;
#pragma unroll 4
        for (int e = 0; e < AFL_HD; e += 4) {
          float4 qv[4], kv[4];
#pragma unroll
          for (int a = 0; a < 4; ++a) { qv[a] = *reinterpret_cast<const float4*>(qp[a] + e); kv[a] = *reinterpret_cast<const float4*>(kp[a] + e); }
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              acc[a][c] = fmaf(qv[a].x, kv[c].x, acc[a][c]); acc[a][c] = fmaf(qv[a].y, kv[c].y, acc[a][c]);
              acc[a][c] = fmaf(qv[a].z, kv[c].z, acc[a][c]); acc[a][c] = fmaf(qv[a].w, kv[c].w, acc[a][c]);
            }
        }
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          const int qi = 4 * bq + a;
          if (qi >= nqv) continue;
          const int qr = qi / N, qe = qi - qr * N;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int ki = 4 * bk + c;
            if (ki >= nkv) continue;
            const float v = scale * acc[a][c];
            if (s1) {
              if (qr == 0) P[qe * N + ki] = v;
              else if (qr == 4) LS[qe * N + ki] = v;
              else A[((qe * 3) + qr - 1) * N + ki] = v;                // A[j][alpha][k]
            } else {
              const int kr = 1 + ki / N, ke = ki - (kr - 1) * N;
              if (kr == 4) LS2[qe * N + ke] = v;
              else T[((ke * 3) + kr - 1) * N + qe] = v;                // T[j][alpha][i]
            }
          }
        }
      }
    }
    __syncthreads();
    // ---- tables, phase 1: 16 lanes per electron j (lane = k): softmax row j, then the own-electron tangent rows of j ----
    const int g = tid >> 4, ln = tid & 15;                 // 16 groups of 16 lanes
    const int lk = ln < N ? ln : N - 1;                    // lanes beyond N shadow the last column and contribute nothing
    const bool gv = g < N;                                 // groups beyond N shadow the last electron (the shuffles below need
    const int gi = gv ? g : N - 1;                         // whole warps) and store nothing
    const bool lv = ln < N && gv;
    {
      const int j = gi;
      const float sjk = ln < N ? P[j * N + lk] : -INFINITY;
      const float mx = afl_gmax(sjk);
      const float e = ln < N ? expf(sjk - mx) : 0.f;
      const float p = e * (1.0f / afl_gsum(e));
      if (lv) P[j * N + lk] = p;
#pragma unroll
      for (int al = 0; al < 3; ++al) {
        const int ja = 3 * j + al;
        const float ds = ln < N ? A[ja * N + lk] + (lk == j ? T[ja * N + j] : 0.f) : 0.f;
        const float m = afl_gsum(p * ds);
        if (lv) Wown[ja * N + lk] = p * (ds - m);
        if (ln == 0 && gv) mown[ja] = m;
      }
      // TT[j][i] = sum_alpha T[j][alpha][i]^2 (lane = i)
      if (lv) {
        const float t0 = T[(3 * j) * N + lk], t1 = T[(3 * j + 1) * N + lk], t2 = T[(3 * j + 2) * N + lk];
        TT[j * N + lk] = fmaf(t0, t0, fmaf(t1, t1, t2 * t2));
      }
    }
    __syncthreads();
    // ---- tables, phase 2: 16 lanes per electron i (lane = k): quad_ik, the Laplacian weights, the cross weights ----
    // quad_ik = sum_c (dS_ik[c] - m_i[c])^2.  Channels of another electron j: dS - m = T_c[i] (delta_kj - P_ij), so their sum is
    // B_i + [k != i] TT[k][i] (1 - 2 P_ik),  B_i = sum_{j != i} TT[j][i] P_ij^2
    {
      const int i = gi;
      const bool in = ln < N;
      const float p = in ? P[i * N + lk] : 0.f;
      const float tti = in ? TT[lk * N + i] : 0.f;
      const float Bi = afl_gsum((in && lk != i) ? tti * p * p : 0.f);
      float qd = Bi + ((in && lk != i) ? tti * (1.0f - 2.0f * p) : 0.f);
#pragma unroll
      for (int al = 0; al < 3; ++al) {
        const int ja = 3 * i + al;
        const float dv = A[ja * N + lk] + (lk == i ? T[ja * N + i] : 0.f) - mown[ja];
        qd = fmaf(dv, dv, qd);
      }
      const float ls = LS[i * N + lk] + LS2[i * N + lk] + (lk == i ? X[i] : 0.f);
      const float lm = afl_gsum(p * ls), qm = afl_gsum(p * qd);
      if (lv) WL[i * N + lk] = p * ((ls - lm) + qd - qm);
      for (int ja = gv ? ln : N3; ja < N3; ja += 16) {
        const int j = ja / 3;
        const float pij = P[i * N + j];
        CX[i * N3 + ja] = (j != i) ? 2.0f * pij * T[ja * N + i] * (1.0f - pij)
                                   : 2.0f * pij * (A[ja * N + i] + T[ja * N + i] - mown[ja]);
      }
    }
    // ---- value rows (also kept in shared memory: every foreign tangent row needs o_i) ----
    const int e4 = ln;                                     // 4 columns per lane
    for (int i = g; i < N; i += AFL_THREADS / 16) {
      float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
      for (int k = 0; k < N; ++k) afl_axpy(y, P[i * N + k], AFL_VEC(k, 0, 2) + 4 * e4);
      *reinterpret_cast<float4*>(O0 + i * AFL_RS + 4 * e4) = y;
      st_row4<PK>(out + ((tok0 + i) * C) * (long long)d, d, col + 4 * e4, y, amax);
    }
    __syncthreads();
    // ---- tangent and Laplacian rows: 8 lanes x 8 columns per item; an item is the Laplacian row of an electron, its three
    //      own tangent rows, or the three rows (i, (j, alpha)) of another electron j, which share v0_j - o_i ----
    {
      const int g8 = tid >> 3, l8 = tid & 7;
      const int ca = 4 * l8, cb = 32 + 4 * l8;             // this lane's two 4-column pieces
      const int nitems = N * N + N;
      for (int it = g8; it < nitems; it += AFL_THREADS / 8) {
        if (it < N) {
          const int i = it;
          float4 y0 = make_float4(0.f, 0.f, 0.f, 0.f), y1 = y0;
#pragma unroll 2
          for (int k = 0; k < N; ++k) {
            const float w = WL[i * N + k], pk = P[i * N + k];
            const float* v0 = AFL_VEC(k, 0, 2);
            const float* lvk = AFL_VEC(k, 4, 2);
            afl_axpy(y0, w, v0 + ca); afl_axpy(y1, w, v0 + cb);
            afl_axpy(y0, pk, lvk + ca); afl_axpy(y1, pk, lvk + cb);
          }
#pragma unroll 3
          for (int ja = 0; ja < N3; ++ja) {
            const int j = ja / 3, al = ja - 3 * j;
            const float cx = CX[i * N3 + ja];
            const float* dv = AFL_VEC(j, 1 + al, 2);
            afl_axpy(y0, cx, dv + ca); afl_axpy(y1, cx, dv + cb);
          }
          float* orow = out + ((tok0 + i) * C + C - 1) * (long long)d;
          st_row4<PK>(orow, d, col + ca, y0, amax);
          st_row4<PK>(orow, d, col + cb, y1, amax);
        } else if (it < 2 * N) {
          const int i = it - N;
          const float pii = P[i * N + i];
#pragma unroll 1
          for (int al = 0; al < 3; ++al) {
            const int ja = 3 * i + al;
            float4 y0 = make_float4(0.f, 0.f, 0.f, 0.f), y1 = y0;
#pragma unroll 2
            for (int k = 0; k < N; ++k) {
              const float w = Wown[ja * N + k];
              const float* v0 = AFL_VEC(k, 0, 2);
              afl_axpy(y0, w, v0 + ca); afl_axpy(y1, w, v0 + cb);
            }
            const float* dv = AFL_VEC(i, 1 + al, 2);
            afl_axpy(y0, pii, dv + ca); afl_axpy(y1, pii, dv + cb);
            float* orow = out + ((tok0 + i) * C + 1 + ja) * (long long)d;
            st_row4<PK>(orow, d, col + ca, y0, amax);
            st_row4<PK>(orow, d, col + cb, y1, amax);
          }
        } else {
          const int pr = it - 2 * N, i = pr / (N - 1), jj = pr - i * (N - 1), j = jj + (jj >= i ? 1 : 0);
          const float pij = P[i * N + j];
          const float* vj = AFL_VEC(j, 0, 2);
          const float* oi = O0 + i * AFL_RS;
          const float4 a0 = *reinterpret_cast<const float4*>(vj + ca), a1 = *reinterpret_cast<const float4*>(vj + cb);
          const float4 b0 = *reinterpret_cast<const float4*>(oi + ca), b1 = *reinterpret_cast<const float4*>(oi + cb);
          const float4 d0 = make_float4(a0.x - b0.x, a0.y - b0.y, a0.z - b0.z, a0.w - b0.w);
          const float4 d1 = make_float4(a1.x - b1.x, a1.y - b1.y, a1.z - b1.z, a1.w - b1.w);
          float* orow = out + ((tok0 + i) * C + 1 + 3 * j) * (long long)d;
#pragma unroll
          for (int al = 0; al < 3; ++al) {
            const float t = T[(3 * j + al) * N + i];
            const float* dv = AFL_VEC(j, 1 + al, 2);
            const float4 e0 = *reinterpret_cast<const float4*>(dv + ca), e1 = *reinterpret_cast<const float4*>(dv + cb);
            const float4 y0 = make_float4(pij * fmaf(t, d0.x, e0.x), pij * fmaf(t, d0.y, e0.y), pij * fmaf(t, d0.z, e0.z), pij * fmaf(t, d0.w, e0.w));
            const float4 y1 = make_float4(pij * fmaf(t, d1.x, e1.x), pij * fmaf(t, d1.y, e1.y), pij * fmaf(t, d1.z, e1.z), pij * fmaf(t, d1.w, e1.w));
            st_row4<PK>(orow + (long long)al * d, d, col + ca, y0, amax);
            st_row4<PK>(orow + (long long)al * d, d, col + cb, y1, amax);
          }
        }
      }
    }
    __syncthreads();      // the next unit overwrites the tables
  }
#undef AFL_VEC
  if (PK) raise_range_flag(ovf, amax);
}

inline bool attention_first_layer_shape(int N, int d, int H) { return H > 0 && d % H == 0 && d / H == AFL_HD && N >= 2 && N <= PSIF_MAX_ELEC; }

// qkv: compact payload [B N][5][3 d]; out: dense payload [B N][C][d] (fp32 rows or the packed fp16 pair)
inline int32_t attention_first_layer(const float* qkv, float* out, long long B, int N, int C, int d, int H, cudaStream_t st, bool packed,
                                     unsigned* ovf, int sms) {
  if (B <= 0) return PSIF_OK;
  if (!attention_first_layer_shape(N, d, H) || C != 3 * N + 2 ||
      ((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(out)) & 15) != 0)
    return fail(PSIF_E_INVALID, "attention: first-layer kernel not available for this shape%s");
  const size_t smem = (size_t)afl_layout(N).total * sizeof(float);
  size_t& configured = dev_smem_cfg().afl;
  if (smem > 48 * 1024 && smem > configured) {
    PSIF_CUDA_CHECK(cudaFuncSetAttribute(attention_first_layer_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PSIF_CUDA_CHECK(cudaFuncSetAttribute(attention_first_layer_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  const long long units = B * H;
  const int per_sm = (int)(200 * 1024 / (smem + 1024)) < 8 ? (int)(200 * 1024 / (smem + 1024)) : 8;
  const long long cap = (long long)(sms > 0 ? sms : 148) * (per_sm > 0 ? per_sm : 1);
  const unsigned grid = (unsigned)(units < cap ? units : cap);
  if (packed) PSIF_LAUNCH(attention_first_layer_kernel<true>, grid, AFL_THREADS, smem, st, qkv, out, units, N, C, d, H, ovf);
  else PSIF_LAUNCH(attention_first_layer_kernel<false>, grid, AFL_THREADS, smem, st, qkv, out, units, N, C, d, H, ovf);
  return PSIF_OK;
}

}  // namespace psif
