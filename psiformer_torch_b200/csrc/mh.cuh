// Metropolis-Hastings walker step (mcmc.py:31-49): all-electron Gaussian proposal, psi-ratio,
// accept/reject, with Philox4x32-10 randomness keyed by (seed, GLOBAL walker id, step) so a chain's
// stream does not depend on how walkers are sharded over GPUs; or with injected noise/uniforms for
// bit-exact replay of the reference's decisions.
//
// Walker state is stored as the reference stores it, x[B][N][3] fp32; one walker's coordinates are
// 12 N contiguous bytes, so a warp moves whole walkers with coalesced accesses.
#pragma once
#include "common.cuh"

namespace psif {

struct Philox4 {
  uint32_t v[4];
};

__host__ __device__ inline Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                 uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
    const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  Philox4 o;
  o.v[0] = c0; o.v[1] = c1; o.v[2] = c2; o.v[3] = c3;
  return o;
}

// counter layout: (walker_lo, walker_hi, step_lo, (step_hi << 8) | slot); slot = electron index, 0xFF = uniform
__device__ __forceinline__ Philox4 mh_random(uint64_t seed, uint64_t walker, uint64_t step, uint32_t slot) {
  return philox4x32_10((uint32_t)walker, (uint32_t)(walker >> 32), (uint32_t)step,
                       ((uint32_t)(step >> 32) << 8) | (slot & 0xFFu), (uint32_t)seed, (uint32_t)(seed >> 32));
}

// three N(0,1) for one electron: Box-Muller on (v0,v1) and (v2,v3)
__device__ __forceinline__ void mh_normals3(const Philox4& r, float& n0, float& n1, float& n2) {
  const float k = 5.9604644775390625e-08f;  // 2^-24
  const float u0 = ((float)(r.v[0] >> 8) + 1.0f) * k;  // (0,1]
  const float u1 = (float)(r.v[1] >> 8) * k;           // [0,1)
  const float u2 = ((float)(r.v[2] >> 8) + 1.0f) * k;
  const float u3 = (float)(r.v[3] >> 8) * k;
  const float ra = sqrtf(-2.0f * logf(u0)), rb = sqrtf(-2.0f * logf(u2));
  float s, c;
  sincosf(6.283185307179586f * u1, &s, &c);
  n0 = ra * c;
  n1 = ra * s;
  n2 = rb * cosf(6.283185307179586f * u3);
}

__device__ __forceinline__ float mh_uniform(const Philox4& r) {
  return (float)(r.v[0] >> 8) * 5.9604644775390625e-08f;  // [0,1), 24 bits like torch.rand
}

// trial = state + step_size * eps   (mcmc.py:33; separate multiply and add, like the eager reference)
// thread per electron
__global__ void __launch_bounds__(256)
mh_propose_kernel(const float* __restrict__ x, float* __restrict__ trial, long long B, int N, float step_size,
                  uint64_t seed, uint64_t walker_id0, uint64_t step, const uint64_t* __restrict__ step_counter,
                  int step_offset, const float* __restrict__ noise) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * N) return;
  float e0, e1, e2;
  if (noise != nullptr) {
    e0 = noise[idx * 3 + 0]; e1 = noise[idx * 3 + 1]; e2 = noise[idx * 3 + 2];
  } else {
    const uint64_t st = (step_counter ? *step_counter : step) + (uint64_t)step_offset;
    const long long b = idx / N;
    const int i = (int)(idx - b * N);
    mh_normals3(mh_random(seed, walker_id0 + (uint64_t)b, st, (uint32_t)i), e0, e1, e2);
  }
  trial[idx * 3 + 0] = __fadd_rn(x[idx * 3 + 0], __fmul_rn(step_size, e0));
  trial[idx * 3 + 1] = __fadd_rn(x[idx * 3 + 1], __fmul_rn(step_size, e1));
  trial[idx * 3 + 2] = __fadd_rn(x[idx * 3 + 2], __fmul_rn(step_size, e2));
}

// accept iff log(u) < min(2 (log|psi'| - log|psi|), 0)   (mcmc.py:40-43); NaN alpha rejects.
// Thread b of a CTA decides for walker b (256 walkers per CTA) and leaves the decision in shared memory; then ALL threads
// walk over the CTA's contiguous span of 256 * 3N coordinates and copy the accepted walkers' trial positions: consecutive
// lanes touch consecutive floats (one thread copying its whole walker strode 12 N bytes from lane to lane).
__global__ void __launch_bounds__(256)
mh_accept_kernel(float* __restrict__ x, const float* __restrict__ trial, float* __restrict__ logabs,
                 const float* __restrict__ logabs_trial, float* __restrict__ sign, const float* __restrict__ sign_trial,
                 uint32_t* __restrict__ status, const uint32_t* __restrict__ status_trial, long long B, int N,
                 uint64_t seed, uint64_t walker_id0, uint64_t step, const uint64_t* __restrict__ step_counter,
                 int step_offset, const float* __restrict__ uniforms, uint8_t* __restrict__ accept_out,
                 unsigned long long* __restrict__ n_accept) {
  __shared__ uint8_t s_acc[256];
  const long long b0 = (long long)blockIdx.x * blockDim.x;
  const long long b = b0 + threadIdx.x;
  bool acc = false;
  if (b < B) {
    float u;
    if (uniforms != nullptr) {
      u = uniforms[b];
    } else {
      const uint64_t st = (step_counter ? *step_counter : step) + (uint64_t)step_offset;
      u = mh_uniform(mh_random(seed, walker_id0 + (uint64_t)b, st, 0xFFu));
    }
    const float lt = logabs_trial[b];
    const float alpha = __fmul_rn(2.0f, __fsub_rn(lt, logabs[b]));
    const float log_accept = fminf(alpha, 0.0f);
    acc = (alpha == alpha) && (logf(u) < log_accept);
    if (acc) {
      logabs[b] = lt;
      if (sign) sign[b] = sign_trial[b];
      if (status) status[b] = status_trial[b];
    }
    if (accept_out) accept_out[b] = acc ? 1 : 0;
  }
  s_acc[threadIdx.x] = acc ? 1 : 0;
  if (n_accept != nullptr) {
    const unsigned m = __ballot_sync(0xffffffffu, acc);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(n_accept, (unsigned long long)__popc(m));
  }
  __syncthreads();
  const int w3 = 3 * N;
  const long long nb = (B - b0) < (long long)blockDim.x ? (B - b0) : (long long)blockDim.x;      // walkers of this CTA
  const long long e0 = b0 * w3;
  for (int e = threadIdx.x; e < (int)(nb * w3); e += blockDim.x)
    if (s_acc[e / w3]) x[e0 + e] = trial[e0 + e];
}

__global__ void mh_advance_counter_kernel(uint64_t* step_counter, int n) { *step_counter += (uint64_t)n; }

// test hook: the raw normal / uniform streams
__global__ void philox_dump_kernel(uint64_t seed, uint64_t walker_id0, uint64_t step, long long B, int N,
                                   float* __restrict__ normals, float* __restrict__ uniform) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * N) return;
  const long long b = idx / N;
  const int i = (int)(idx - b * N);
  float e0, e1, e2;
  mh_normals3(mh_random(seed, walker_id0 + (uint64_t)b, step, (uint32_t)i), e0, e1, e2);
  normals[idx * 3 + 0] = e0; normals[idx * 3 + 1] = e1; normals[idx * 3 + 2] = e2;
  if (i == 0 && uniform) uniform[b] = mh_uniform(mh_random(seed, walker_id0 + (uint64_t)b, step, 0xFFu));
}

}  // namespace psif
