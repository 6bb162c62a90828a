// Per-token payload kernels: feature embedding, LayerNorm, GELU, orbital*envelope product.
// Payload layout P[b][i][c][e]; C = 1 (value only) or 3N+2 (value, 3N tangents, Laplacian).
// Propagation rules: SURVEY App. B (checked against autograd in tests/test_forward_laplacian.py).
#pragma once
#include "common.cuh"

namespace psif {

struct Nuclei {
  int natom;
  float R[PSIF_MAX_ATOMS][3];
  float Z[PSIF_MAX_ATOMS];
};

// ------------------------------------------------------------------------------------------
// embed: features [r_i - R_I, |r_i - R_I|] -> l_0   (psiformer.py:233-236)
// grid = B*N tokens, block = min(d rounded to 32, 256)
// ------------------------------------------------------------------------------------------
__global__ void embed_kernel(const float* __restrict__ x, const float* __restrict__ W0,
                             const float* __restrict__ b0, float* __restrict__ out, int N, int C, int d,
                             Nuclei nuc) {
  const long long tok = blockIdx.x;
  const int i = (int)(tok % N);
  const float px = x[tok * 3 + 0], py = x[tok * 3 + 1], pz = x[tok * 3 + 2];
  float disp[PSIF_MAX_ATOMS][3], r[PSIF_MAX_ATOMS], rinv[PSIF_MAX_ATOMS];
#pragma unroll
  for (int a = 0; a < PSIF_MAX_ATOMS; ++a) {
    if (a < nuc.natom) {
      disp[a][0] = px - nuc.R[a][0];
      disp[a][1] = py - nuc.R[a][1];
      disp[a][2] = pz - nuc.R[a][2];
      r[a] = sqrtf(disp[a][0] * disp[a][0] + disp[a][1] * disp[a][1] + disp[a][2] * disp[a][2]);
      rinv[a] = 1.0f / r[a];
    }
  }
  const int nf = 4 * nuc.natom;
  float* o = out + tok * (long long)C * d;
  for (int e = threadIdx.x; e < d; e += blockDim.x) {
    const float* w = W0 + (long long)e * nf;
    float val = b0[e], t0 = 0.f, t1 = 0.f, t2 = 0.f, lap = 0.f;
#pragma unroll
    for (int a = 0; a < PSIF_MAX_ATOMS; ++a) {
      if (a < nuc.natom) {
        const float w0 = w[4 * a + 0], w1 = w[4 * a + 1], w2 = w[4 * a + 2], w3 = w[4 * a + 3];
        val += w0 * disp[a][0] + w1 * disp[a][1] + w2 * disp[a][2] + w3 * r[a];
        t0 += w0 + w3 * disp[a][0] * rinv[a];
        t1 += w1 + w3 * disp[a][1] * rinv[a];
        t2 += w2 + w3 * disp[a][2] * rinv[a];
        lap += w3 * 2.0f * rinv[a];
      }
    }
    o[e] = val;
    if (C > 1) {
      for (int c = 1; c < C - 1; ++c) {
        const int own = c - (1 + 3 * i);
        o[(long long)c * d + e] = own == 0 ? t0 : own == 1 ? t1 : own == 2 ? t2 : 0.f;
      }
      o[(long long)(C - 1) * d + e] = lap;
    }
  }
}

// value path (C == 1): one WARP per token, 8 tokens per CTA -- with one 256-thread CTA per token the kernel was launch
// bound (49 us for 20480 tokens on Ne, ncu round 2: as long as the QKV GEMM behind it).  Same expressions as above.
__global__ void __launch_bounds__(256)
embed_value_kernel(const float* __restrict__ x, const float* __restrict__ W0, const float* __restrict__ b0,
                   float* __restrict__ out, long long T, int d, Nuclei nuc) {
  const long long tok = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (tok >= T) return;
  const int lane = threadIdx.x & 31;
  const float px = x[tok * 3 + 0], py = x[tok * 3 + 1], pz = x[tok * 3 + 2];
  float disp[PSIF_MAX_ATOMS][3], r[PSIF_MAX_ATOMS];
#pragma unroll
  for (int a = 0; a < PSIF_MAX_ATOMS; ++a) {
    if (a < nuc.natom) {
      disp[a][0] = px - nuc.R[a][0];
      disp[a][1] = py - nuc.R[a][1];
      disp[a][2] = pz - nuc.R[a][2];
      r[a] = sqrtf(disp[a][0] * disp[a][0] + disp[a][1] * disp[a][1] + disp[a][2] * disp[a][2]);
    }
  }
  const int nf = 4 * nuc.natom;
  float* o = out + tok * (long long)d;
  for (int e = lane; e < d; e += 32) {
    const float* w = W0 + (long long)e * nf;
    float val = b0[e];
#pragma unroll
    for (int a = 0; a < PSIF_MAX_ATOMS; ++a) {
      if (a < nuc.natom) {
        const float w0 = w[4 * a + 0], w1 = w[4 * a + 1], w2 = w[4 * a + 2], w3 = w[4 * a + 3];
        val += w0 * disp[a][0] + w1 * disp[a][1] + w2 * disp[a][2] + w3 * r[a];
      }
    }
    o[e] = val;
  }
}

// ------------------------------------------------------------------------------------------
// LayerNorm over d (psiformer.py:86-87), App. B LayerNorm row.
// One CTA per token, LN_WARPS warps; warp w handles tangent channels w, w+LN_WARPS, ...
// Each lane owns elements e = lane + 32*t, t < EPL.
// ------------------------------------------------------------------------------------------
constexpr int LN_WARPS = 8;

template <int EPL>
__global__ void __launch_bounds__(LN_WARPS * 32)
layernorm_payload_kernel(const float* __restrict__ in, const float* __restrict__ gamma,
                         const float* __restrict__ beta, float* __restrict__ out, int C, int d) {
  extern __shared__ float ln_smem[];  // [LN_WARPS][d] Laplacian corrections
  const long long tok = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* ip = in + tok * (long long)C * d;
  float* op = out + tok * (long long)C * d;
  const float inv_d = 1.0f / (float)d;

  float ah[EPL], gam[EPL], corr[EPL];
  float s;
  {
    float v[EPL];
    float sum = 0.f;
#pragma unroll
    for (int t = 0; t < EPL; ++t) {
      const int e = lane + 32 * t;
      v[t] = e < d ? ip[e] : 0.f;
      gam[t] = e < d ? gamma[e] : 0.f;
      sum += v[t];
    }
    const float mean = warp_sum(sum) * inv_d;
    float sq = 0.f;
#pragma unroll
    for (int t = 0; t < EPL; ++t) {
      const int e = lane + 32 * t;
      v[t] = e < d ? v[t] - mean : 0.f;
      sq += v[t] * v[t];
    }
    s = rsqrtf(warp_sum(sq) * inv_d + kLnEps);
#pragma unroll
    for (int t = 0; t < EPL; ++t) {
      ah[t] = v[t] * s;
      corr[t] = 0.f;
    }
    if (warp == 0) {
#pragma unroll
      for (int t = 0; t < EPL; ++t) {
        const int e = lane + 32 * t;
        if (e < d) op[e] = gam[t] * ah[t] + beta[e];
      }
    }
  }
  if (C == 1) return;

  const float s2 = s * s;
  for (int c = 1 + warp; c < C - 1; c += LN_WARPS) {
    float v[EPL];
    float sum = 0.f;
#pragma unroll
    for (int t = 0; t < EPL; ++t) {
      const int e = lane + 32 * t;
      v[t] = e < d ? ip[(long long)c * d + e] : 0.f;
      sum += v[t];
    }
    const float mu = warp_sum(sum) * inv_d;
    float sm = 0.f, sq = 0.f;
#pragma unroll
    for (int t = 0; t < EPL; ++t) {
      const int e = lane + 32 * t;
      v[t] = e < d ? v[t] - mu : 0.f;
      sm += ah[t] * v[t];
      sq += v[t] * v[t];
    }
    const float m = warp_sum(sm) * inv_d;
    const float q = warp_sum(sq) * inv_d;
    const float k2 = s2 * (q - m * m);
#pragma unroll
    for (int t = 0; t < EPL; ++t) {
      const int e = lane + 32 * t;
      const float w = v[t] - ah[t] * m;
      if (e < d) op[(long long)c * d + e] = gam[t] * s * w;
      corr[t] += -2.0f * s2 * m * w - ah[t] * k2;
    }
  }
#pragma unroll
  for (int t = 0; t < EPL; ++t) {
    const int e = lane + 32 * t;
    if (e < d) ln_smem[warp * d + e] = corr[t];
  }
  __syncthreads();
  if (warp == 0) {
    float v[EPL];
    float sum = 0.f;
#pragma unroll
    for (int t = 0; t < EPL; ++t) {
      const int e = lane + 32 * t;
      v[t] = e < d ? ip[(long long)(C - 1) * d + e] : 0.f;
      sum += v[t];
    }
    const float mu = warp_sum(sum) * inv_d;
    float sm = 0.f;
#pragma unroll
    for (int t = 0; t < EPL; ++t) {
      const int e = lane + 32 * t;
      v[t] = e < d ? v[t] - mu : 0.f;
      sm += ah[t] * v[t];
    }
    const float m = warp_sum(sm) * inv_d;
#pragma unroll
    for (int t = 0; t < EPL; ++t) {
      const int e = lane + 32 * t;
      if (e < d) {
        float tot = 0.f;
#pragma unroll
        for (int w = 0; w < LN_WARPS; ++w) tot += ln_smem[w * d + e];
        op[(long long)(C - 1) * d + e] = gam[t] * (s * (v[t] - ah[t] * m) + tot);
      }
    }
  }
}

// value-only LayerNorm: one warp per token
template <int EPL>
__global__ void __launch_bounds__(256)
layernorm_value_kernel(const float* __restrict__ in, const float* __restrict__ gamma,
                       const float* __restrict__ beta, float* __restrict__ out, long long tokens, int d) {
  const long long tok = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (tok >= tokens) return;
  const int lane = threadIdx.x & 31;
  const float* ip = in + tok * d;
  float* op = out + tok * d;
  const float inv_d = 1.0f / (float)d;
  float v[EPL];
  float sum = 0.f;
#pragma unroll
  for (int t = 0; t < EPL; ++t) {
    const int e = lane + 32 * t;
    v[t] = e < d ? ip[e] : 0.f;
    sum += v[t];
  }
  const float mean = warp_sum(sum) * inv_d;
  float sq = 0.f;
#pragma unroll
  for (int t = 0; t < EPL; ++t) {
    const int e = lane + 32 * t;
    v[t] = e < d ? v[t] - mean : 0.f;
    sq += v[t] * v[t];
  }
  const float s = rsqrtf(warp_sum(sq) * inv_d + kLnEps);
#pragma unroll
  for (int t = 0; t < EPL; ++t) {
    const int e = lane + 32 * t;
    if (e < d) op[e] = gamma[e] * (v[t] * s) + beta[e];
  }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm on payloads, warp-per-token variant for d % 128 == 0 (float4 per lane, V = d/128 of them):
// no shared memory, no block barrier; the token's C rows stream through one warp, U rows in flight at a time.
// ------------------------------------------------------------------------------------------------
template <int V>
__device__ __forceinline__ float ln_sum4(const float4 (&v)[V]) {
  float s = 0.f;
#pragma unroll
  for (int t = 0; t < V; ++t) s += (v[t].x + v[t].y) + (v[t].z + v[t].w);
  return s;
}

// PK: the output is written as the packed fp16 pair the tensor-core GEMM consumes (common.cuh); ovf = range flag.
// SP (first layer, `nel` electrons per walker): in front of the first attention token i only depends on x_i, so of its
// C = 3 nel + 2 rows only the value, its own three tangents and the Laplacian are non-zero; only those are read, and the
// output is the COMPACT payload [token][5][d] (the zero rows add exact zeros to every sum below, so the five rows are
// bit-identical to the dense result).
template <int V, bool PK, bool SP = false>
__global__ void __launch_bounds__(256)
layernorm_payload_warp_kernel(const float* __restrict__ in, const float* __restrict__ gamma,
                              const float* __restrict__ beta, float* __restrict__ out, long long tokens, int C, unsigned* ovf,
                              int nel = 0) {
  constexpr int d = 128 * V;
  constexpr int U = 4;  // tangent rows in flight
  const long long tok = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (tok >= tokens) return;
  const int lane = threadIdx.x & 31;
  const float4* ip = reinterpret_cast<const float4*>(in + tok * (long long)C * d) + lane;
  const int own = SP ? 1 + 3 * (int)(tok % nel) : 1;     // first and one-past-last tangent row that is read
  const int c_end = SP ? own + 3 : C - 1;
  const int CO = SP ? 5 : C;                              // rows per token in the output
  float* orow = out + tok * (long long)CO * d;           // output row c of the token starts at orow + c * d
  const float inv_d = 1.0f / (float)d;
  constexpr int RV = d / 4;  // float4 per row
  float amax = 0.f;

  float4 gam[V], ah[V], corr[V];
  float s;
  {
    float4 v[V];
#pragma unroll
    for (int t = 0; t < V; ++t) {
      v[t] = __ldg(ip + 32 * t);
      gam[t] = __ldg(reinterpret_cast<const float4*>(gamma) + lane + 32 * t);
    }
    const float mean = warp_sum(ln_sum4<V>(v)) * inv_d;
    float sq = 0.f;
#pragma unroll
    for (int t = 0; t < V; ++t) {
      v[t].x -= mean; v[t].y -= mean; v[t].z -= mean; v[t].w -= mean;
      sq += (v[t].x * v[t].x + v[t].y * v[t].y) + (v[t].z * v[t].z + v[t].w * v[t].w);
    }
    s = rsqrtf(warp_sum(sq) * inv_d + kLnEps);
#pragma unroll
    for (int t = 0; t < V; ++t) {
      ah[t] = make_float4(v[t].x * s, v[t].y * s, v[t].z * s, v[t].w * s);
      corr[t] = make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + lane + 32 * t);
      st_row4<PK>(orow, d, 4 * (lane + 32 * t), make_float4(gam[t].x * ah[t].x + b.x, gam[t].y * ah[t].y + b.y,
                                                            gam[t].z * ah[t].z + b.z, gam[t].w * ah[t].w + b.w), amax);
    }
  }
  if (C == 1) { if (PK) raise_range_flag(ovf, amax); return; }
  const float s2 = s * s;
  for (int c0 = own; c0 < c_end; c0 += U) {
    float4 v[U][V];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int t = 0; t < V; ++t)
        v[u][t] = (c0 + u < c_end) ? __ldg(ip + (long long)(c0 + u) * RV + 32 * t) : make_float4(0.f, 0.f, 0.f, 0.f);
    float mu[U];
#pragma unroll
    for (int u = 0; u < U; ++u) mu[u] = ln_sum4<V>(v[u]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int u = 0; u < U; ++u) mu[u] += __shfl_xor_sync(0xffffffffu, mu[u], o);
    float sm[U], sq[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const float m0 = mu[u] * inv_d;
      float a = 0.f, q = 0.f;
#pragma unroll
      for (int t = 0; t < V; ++t) {
        v[u][t].x -= m0; v[u][t].y -= m0; v[u][t].z -= m0; v[u][t].w -= m0;
        a += (ah[t].x * v[u][t].x + ah[t].y * v[u][t].y) + (ah[t].z * v[u][t].z + ah[t].w * v[u][t].w);
        q += (v[u][t].x * v[u][t].x + v[u][t].y * v[u][t].y) + (v[u][t].z * v[u][t].z + v[u][t].w * v[u][t].w);
      }
      sm[u] = a; sq[u] = q;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int u = 0; u < U; ++u) {
        sm[u] += __shfl_xor_sync(0xffffffffu, sm[u], o);
        sq[u] += __shfl_xor_sync(0xffffffffu, sq[u], o);
      }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (c0 + u < c_end) {
        const float m = sm[u] * inv_d, q = sq[u] * inv_d;
        const float k2 = s2 * (q - m * m), k1 = -2.0f * s2 * m;
#pragma unroll
        for (int t = 0; t < V; ++t) {
          float4 w;
          w.x = v[u][t].x - ah[t].x * m; w.y = v[u][t].y - ah[t].y * m;
          w.z = v[u][t].z - ah[t].z * m; w.w = v[u][t].w - ah[t].w * m;
          st_row4<PK>(orow + (long long)(SP ? 1 + c0 + u - own : c0 + u) * d, d, 4 * (lane + 32 * t),
                      make_float4(gam[t].x * s * w.x, gam[t].y * s * w.y, gam[t].z * s * w.z, gam[t].w * s * w.w), amax);
          corr[t].x += k1 * w.x - ah[t].x * k2; corr[t].y += k1 * w.y - ah[t].y * k2;
          corr[t].z += k1 * w.z - ah[t].z * k2; corr[t].w += k1 * w.w - ah[t].w * k2;
        }
      }
    }
  }
  {
    float4 v[V];
#pragma unroll
    for (int t = 0; t < V; ++t) v[t] = __ldg(ip + (long long)(C - 1) * RV + 32 * t);
    const float mu = warp_sum(ln_sum4<V>(v)) * inv_d;
    float a = 0.f;
#pragma unroll
    for (int t = 0; t < V; ++t) {
      v[t].x -= mu; v[t].y -= mu; v[t].z -= mu; v[t].w -= mu;
      a += (ah[t].x * v[t].x + ah[t].y * v[t].y) + (ah[t].z * v[t].z + ah[t].w * v[t].w);
    }
    const float m = warp_sum(a) * inv_d;
#pragma unroll
    for (int t = 0; t < V; ++t)
      st_row4<PK>(orow + (long long)(CO - 1) * d, d, 4 * (lane + 32 * t),
                  make_float4(gam[t].x * (s * (v[t].x - ah[t].x * m) + corr[t].x), gam[t].y * (s * (v[t].y - ah[t].y * m) + corr[t].y),
                              gam[t].z * (s * (v[t].z - ah[t].z * m) + corr[t].z), gam[t].w * (s * (v[t].w - ah[t].w * m) + corr[t].w)),
                  amax);
  }
  if (PK) raise_range_flag(ovf, amax);
}

// fp32 payload -> packed fp16 pair (the residual stream in front of the orbital-head GEMM): thread per float4
__global__ void __launch_bounds__(256)
pack_payload_kernel(const float* __restrict__ in, float* __restrict__ out, long long rows, int w, unsigned* ovf) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int w4 = w >> 2;
  float amax = 0.f;
  if (idx < rows * w4) {
    const long long r = idx / w4;
    const int c4 = (int)(idx - r * w4);
    const float4 v = __ldg(reinterpret_cast<const float4*>(in + r * w) + c4);
    st_row4<true>(out + r * w, w, 4 * c4, v, amax);
  }
  raise_range_flag(ovf, amax);
}

// layernorm shapes whose output can be written as the packed fp16 pair (the warp-per-token kernel)
inline bool layernorm_can_pack(int d) { return d == 128 || d == 256 || d == 512; }

// packed: write the output as the packed fp16 pair (common.cuh) and raise *ovf when a value does not fit fp16
// sparse_nel > 0: first-layer mode, compact [token][5][d] output (see the kernel)
inline int32_t layernorm_payload(const float* in, const float* gamma, const float* beta, float* out,
                                 long long tokens, int C, int d, cudaStream_t st, bool packed = false, unsigned* ovf = nullptr,
                                 int sparse_nel = 0) {
  if (tokens <= 0) return PSIF_OK;
  if (d > 1024) return fail(PSIF_E_INVALID, "layernorm: n_embd > 1024 unsupported%s");
  if (sparse_nel > 0) {
    const bool al = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(gamma) |
                      reinterpret_cast<uintptr_t>(beta)) & 15) == 0;
    if (!(al && layernorm_can_pack(d)) || C != 3 * sparse_nel + 2) return fail(PSIF_E_INVALID, "layernorm: first-layer mode needs n_embd in {128, 256, 512} and C = 3 N + 2%s");
    const unsigned grid = (unsigned)cdiv(tokens, 8);
#define PSIF_LNS(VV) do { if (packed) PSIF_LAUNCH((layernorm_payload_warp_kernel<VV, true, true>), grid, 256, 0, st, in, gamma, beta, out, tokens, C, ovf, sparse_nel); \
                          else PSIF_LAUNCH((layernorm_payload_warp_kernel<VV, false, true>), grid, 256, 0, st, in, gamma, beta, out, tokens, C, ovf, sparse_nel); } while (0)
    if (d == 128) PSIF_LNS(1); else if (d == 256) PSIF_LNS(2); else PSIF_LNS(4);
#undef PSIF_LNS
    return PSIF_OK;
  }
  const bool al16 = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(gamma) |
                      reinterpret_cast<uintptr_t>(beta)) & 15) == 0;
  if (packed && !(al16 && layernorm_can_pack(d))) return fail(PSIF_E_INVALID, "layernorm: packed output needs n_embd in {128, 256, 512}%s");
  if (al16 && layernorm_can_pack(d)) {   // also for C == 1: value and energy paths share the arithmetic
    const unsigned grid = (unsigned)cdiv(tokens, 8);
#define PSIF_LNW(VV) do { if (packed) PSIF_LAUNCH((layernorm_payload_warp_kernel<VV, true>), grid, 256, 0, st, in, gamma, beta, out, tokens, C, ovf, 0); \
                          else PSIF_LAUNCH((layernorm_payload_warp_kernel<VV, false>), grid, 256, 0, st, in, gamma, beta, out, tokens, C, ovf, 0); } while (0)
    if (d == 128) PSIF_LNW(1); else if (d == 256) PSIF_LNW(2); else PSIF_LNW(4);
#undef PSIF_LNW
    return PSIF_OK;
  }
  if (C == 1) {
    const unsigned grid = (unsigned)cdiv(tokens, 8);
#define PSIF_LNV(E) PSIF_LAUNCH(layernorm_value_kernel<E>, grid, 256, 0, st, in, gamma, beta, out, tokens, d)
    if (d <= 32) PSIF_LNV(1); else if (d <= 64) PSIF_LNV(2); else if (d <= 128) PSIF_LNV(4);
    else if (d <= 256) PSIF_LNV(8); else if (d <= 512) PSIF_LNV(16); else PSIF_LNV(32);
#undef PSIF_LNV
    return PSIF_OK;
  }
  const size_t smem = (size_t)LN_WARPS * d * sizeof(float);
#define PSIF_LNP(E) PSIF_LAUNCH(layernorm_payload_kernel<E>, (unsigned)tokens, LN_WARPS * 32, smem, st, in, gamma, beta, out, C, d)
  if (d <= 32) PSIF_LNP(1); else if (d <= 64) PSIF_LNP(2); else if (d <= 128) PSIF_LNP(4);
  else if (d <= 256) PSIF_LNP(8); else if (d <= 512) PSIF_LNP(16); else PSIF_LNP(32);
#undef PSIF_LNP
  return PSIF_OK;
}

// ------------------------------------------------------------------------------------------
// GELU(tanh) on payloads (psiformer.py:75): thread per (token, e), loops over channels
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gelu_payload_kernel(const float* __restrict__ in, float* __restrict__ out, long long tokens, int C, int width) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= tokens * width) return;
  const long long tok = idx / width;
  const int e = (int)(idx % width);
  const float* ip = in + tok * (long long)C * width + e;
  float* op = out + tok * (long long)C * width + e;
  float g, g1, g2;
  gelu_tanh_d2(ip[0], g, g1, g2);
  op[0] = g;
  if (C == 1) return;
  float ss = 0.f;
  for (int c = 1; c < C - 1; ++c) {
    const float t = ip[(long long)c * width];
    ss = fmaf(t, t, ss);
    op[(long long)c * width] = g1 * t;
  }
  op[(long long)(C - 1) * width] = fmaf(g1, ip[(long long)(C - 1) * width], g2 * ss);
}

inline int32_t gelu_payload(const float* in, float* out, long long tokens, int C, int width, cudaStream_t st) {
  const long long n = tokens * width;
  if (n <= 0) return PSIF_OK;
  PSIF_LAUNCH(gelu_payload_kernel, (unsigned)cdiv(n, 256), 256, 0, st, in, out, tokens, C, width);
  return PSIF_OK;
}

// ------------------------------------------------------------------------------------------
// orbital * envelope (psiformer.py:115-120, 172): in place on the orbital-linear payload
//   lin[b][i][c][col], col in [0, K*n_up) for the up head, [K*n_up, K*(n_up+n_dn)) for down.
// Only the columns of token i's own spin are transformed (the others are never read).
// sigma/pi are the already clamped values, laid out [natom][Korb] with the same column index.
// ------------------------------------------------------------------------------------------
// tpt threads per token (a power of two dividing 128): 128 in energy mode (one CTA per token, C rows per column), 32 in value
// mode, where one row per token left a 128-thread CTA with 16 .. 224 columns of work and the launch latency bound.
__global__ void __launch_bounds__(128)
orbital_envelope_kernel(float* __restrict__ lin, const float* __restrict__ x, const float* __restrict__ sigma,
                        const float* __restrict__ pi, int N, int n_up, int C, int Kup, int Korb, Nuclei nuc, long long tokens,
                        int tpt) {
  const long long tok = (long long)blockIdx.x * (128 / tpt) + threadIdx.x / tpt;
  if (tok >= tokens) return;
  const int tl = threadIdx.x % tpt;
  const int i = (int)(tok % N);
  const int col0 = i < n_up ? 0 : Kup;
  const int ncol = i < n_up ? Kup : Korb - Kup;
  const float px = x[tok * 3 + 0], py = x[tok * 3 + 1], pz = x[tok * 3 + 2];
  float* base = lin + tok * (long long)C * Korb;
  for (int cc = tl; cc < ncol; cc += tpt) {
    const int col = col0 + cc;
    float e0 = 0.f, g0 = 0.f, g1 = 0.f, g2 = 0.f, el = 0.f;
#pragma unroll
    for (int a = 0; a < PSIF_MAX_ATOMS; ++a) {
      if (a < nuc.natom) {
        const float dx = px - nuc.R[a][0], dy = py - nuc.R[a][1], dz = pz - nuc.R[a][2];
        const float r = sqrtf(dx * dx + dy * dy + dz * dz);
        const float sg = sigma[a * Korb + col];
        const float ev = pi[a * Korb + col] * expf(-r * sg);
        const float rinv = 1.0f / r;
        e0 += ev;
        const float k = -sg * ev * rinv;
        g0 += k * dx; g1 += k * dy; g2 += k * dz;
        el += ev * (sg * sg - 2.0f * sg * rinv);
      }
    }
    float* p = base + col;
    const float l0 = p[0];
    p[0] = l0 * e0;
    if (C > 1) {
      float cross = 0.f;
      // sixteen rows at a time: all loads first, then the arithmetic and the stores.  In place and with a run-time row
      // pitch the compiler must assume that a store may feed the next load, so the plain loop was one global round trip
      // per row (1.6 TB/s on N2, where this kernel moves 36 GB per energy pass).
      for (int c0 = 1; c0 < C - 1; c0 += 16) {
        float lc[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) lc[u] = (c0 + u < C - 1) ? p[(long long)(c0 + u) * Korb] : 0.f;
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const int c = c0 + u;
          if (c < C - 1) {
            const int own = c - (1 + 3 * i);
            float v = lc[u] * e0;
            if (own >= 0 && own < 3) {
              const float eg = own == 0 ? g0 : own == 1 ? g1 : g2;
              v += l0 * eg;
              cross += lc[u] * eg;
            }
            p[(long long)c * Korb] = v;
          }
        }
      }
      const float ll = p[(long long)(C - 1) * Korb];
      p[(long long)(C - 1) * Korb] = ll * e0 + l0 * el + 2.0f * cross;
    }
  }
}

}  // namespace psif
