/*
 * psiformer_b200.h — C ABI of libpsiformer_b200.so (sm_100a).
 *
 * The reference (jorgemunozl/psiformer_torch) has no FFI layer: its boundary is
 * the Python API.  This header is the boundary a replacement binds instead; each
 * entry point names the reference code it stands in for (paths relative to
 * /root/reference/src/psiformer_torch/).  The ctypes stubs that the Python
 * drop-in package uses are shown in INTEGRATION.md.
 *
 * Conventions
 *   - every function returns 0 on success or a negative PSIF_E_* code, never
 *     throws; psif_last_error() gives a thread-local message;
 *   - all array pointers are CUDA DEVICE pointers unless marked "host";
 *   - work is enqueued on the caller's stream (a cudaStream_t passed as void*)
 *     and the call returns without synchronising;
 *   - the caller owns all memory (inputs, outputs, workspace).  The handle owns
 *     only its configuration and a private copy of the packed parameters;
 *   - one host thread per handle; handles are independent (one per GPU / rank): the library keeps no process-wide
 *     mutable state apart from the thread-local error string, the launch counter (an atomic statistic) and an
 *     idempotent per-device record of kernel attributes already set;
 *
 * Tensor layouts are row-major and contiguous, fp32 unless stated:
 *   walkers      x[B][N][3]         N = n_up + n_dn, first n_up electrons are spin-up
 *   parameters   one blob in state_dict order:
 *                l_0.{weight[d][4*natom],bias[d]},
 *                per layer: attn.c_attn.{weight[3d][d],bias}, attn.c_proj.{weight[d][d],bias},
 *                           mlp.c_fc.{weight[4d][d],bias}, mlp.c_proj.{weight[d][4d],bias},
 *                           ln_1.{weight,bias}, ln_2.{weight,bias},
 *                orbital_head.det_logits[K], envelope_up.{pi,raw_sigma}[natom][K*n_up],
 *                envelope_down.{pi,raw_sigma}[natom][K*n_dn],
 *                orb_up.{weight[K*n_up][d],bias}, orb_down.{weight[K*n_dn][d],bias},
 *                jastrow.alpha_anti[1], jastrow.alpha_par[1]
 */
#ifndef PSIFORMER_B200_H
#define PSIFORMER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PSIF_MAX_ATOMS 8
#define PSIF_MAX_ELEC 16      /* electrons per walker                      */
#define PSIF_MAX_SPIN 8       /* electrons per spin channel (LU size)       */
#define PSIF_MAX_DET 64

#define PSIF_OK 0
#define PSIF_E_INVALID (-1)   /* bad argument / unsupported shape           */
#define PSIF_E_CUDA (-2)      /* a CUDA runtime call failed                 */
#define PSIF_E_WORKSPACE (-3) /* workspace too small                        */
#define PSIF_E_STATE (-4)     /* parameters not set                         */

/* per-walker status bits written by psif_logpsi / psif_local_energy / psif_mh_steps */
#define PSIF_ST_NONFINITE_LOGDET 1u /* psiformer.py:256-257 would raise ValueError           */
#define PSIF_ST_CLAMP_SUSPECT 2u    /* a block has a singular value < 1e-6: the clamp of logdet_matmul.py:50-51 is active
                                       (value and, in energy mode, derivatives follow the clamp; informational) */
#define PSIF_ST_NONFINITE_ELOC 4u   /* train.py:86-90 would drop the entry                   */
#define PSIF_ST_FLOOR 8u            /* |sum_k ...| < 1e-12 floor of logdet_matmul.py:68 hit   */
#define PSIF_ST_FP16_RANGE 16u      /* an activation of this walker's chunk exceeded fp16's range in a split-fp16 GEMM:
                                       repeat the call after psif_set_gemm_mode(h, PSIF_GEMM_TF32_SPLIT)             */

/* operand split of the tensor-core Linear (nn.Linear in psiformer.py:39,63,74,76); both give fp32-grade results */
#define PSIF_GEMM_FP16_SPLIT 0 /* x = h0 + 2^-11 h1 in fp16: default, twice the tensor throughput, |x| < 65504 */
#define PSIF_GEMM_TF32_SPLIT 1 /* x = hi + lo in tf32: full fp32 range                                        */

#define PSIF_MODE_VALUE 0  /* log|psi| only (Metropolis)                     */
#define PSIF_MODE_ENERGY 1 /* value + 3N tangents + Laplacian                */

typedef struct PsifConfig {
  int32_t n_layer, n_head, n_embd, n_det, n_up, n_dn, natom, reserved;
  double Z[PSIF_MAX_ATOMS];    /* nuclear charges                           */
  double R[PSIF_MAX_ATOMS][3]; /* nuclear positions (bohr)                  */
} PsifConfig;

typedef struct PsifHandle PsifHandle;

/* Model_Config -> handle (config.py:6-16; psiformer.py:204-217). host_cfg is a host pointer. */
int32_t psif_create(const PsifConfig* host_cfg, PsifHandle** out);
int32_t psif_destroy(PsifHandle* h);

/* number of floats in the packed parameter blob (utils/count_params.py:32-51) */
int32_t psif_param_count(const PsifHandle* h, size_t* n_floats);

/* load_state_dict equivalent: copies the blob and derives softmax(det_logits)
 * (psiformer.py:190), clamped envelope sigma/pi (:115-117), fused orbital weights. */
int32_t psif_set_params(PsifHandle* h, const float* packed_params, size_t n_floats, void* stream);

/* operand split of the tensor-core Linear: PSIF_GEMM_FP16_SPLIT (default) or PSIF_GEMM_TF32_SPLIT; the host side
 * switches to the latter and repeats a call whose status words carry PSIF_ST_FP16_RANGE. */
int32_t psif_set_gemm_mode(PsifHandle* h, int32_t mode);

/* The same event as PSIF_ST_FP16_RANGE, for the host side: *out = 1 if a forward pass of this handle (psif_logpsi,
 * psif_local_energy, psif_mh_steps) raised it since the last call of this function, else 0; clears it.  The event is
 * mirrored into pinned host memory by the pass itself, so after synchronising the pass' stream this is a host read --
 * no reduction over the status array, no device-to-host copy.  Replaces nothing in the reference (its fp32 matmuls
 * have no range limit); it is how Engine.local_energy / logpsi decide to repeat a call in PSIF_GEMM_TF32_SPLIT mode. */
int32_t psif_take_range_event(PsifHandle* h, int32_t* out);

/* bytes of scratch the caller must pass as `ws` for B walkers in `mode` */
int32_t psif_workspace_bytes(const PsifHandle* h, int64_t B, int32_t mode, size_t* out);

/* PsiFormer.forward (psiformer.py:220-264) plus the sign dropped at :191. */
int32_t psif_logpsi(PsifHandle* h, const float* x, int64_t B, float* logabs, float* sign,
                    uint32_t* status, void* ws, size_t ws_bytes, void* stream);

/* Hamiltonian.local_energy (hamiltonian.py:46-54) with grad_log_psi (:56-68),
 * laplacian_log_psi (:70-95) and Potential.potential (:15-35) as by-products.
 * grad[B][N][3], lap[B], pot[B] may be NULL.  accum (3 doubles on the device, may be NULL)
 * is atomically incremented by {sum E_L, sum E_L^2, n} over the walkers the reference's masks keep (finite log|psi|
 * and E_L, train.py:79-90: neither PSIF_ST_NONFINITE_* bit set) -- the collective point of train.py:138. */
int32_t psif_local_energy(PsifHandle* h, const float* x, int64_t B, float* e_loc, float* logabs,
                          float* sign, float* grad, float* lap, float* pot, double* accum,
                          uint32_t* status, void* ws, size_t ws_bytes, void* stream);

/* MH._run_steps (mcmc.py:31-54) with log|psi(current)| cached between steps.
 *   x_inout[B][N][3], logabs_inout[B]: chain state; if have_logabs == 0 the cache is
 *   (re)computed first.  Randomness: Philox4x32-10 keyed by (seed, walker_id0 + b, step)
 *   where step = step0 + s, or *step_counter (device, incremented by n_steps) if non-NULL;
 *   or injected: noise[n_steps][B][N][3] replaces randn_like (:33), uniforms[n_steps][B]
 *   replaces rand_like (:42).  accept_out[n_steps][B] (may be NULL) receives the masks of
 *   :43; n_accept (device, may be NULL) is incremented by the number of accepted moves. */
int32_t psif_mh_steps(PsifHandle* h, float* x_inout, float* logabs_inout, float* sign_inout,
                      int64_t B, int32_t n_steps, float step_size, int32_t have_logabs,
                      uint64_t seed, uint64_t walker_id0, uint64_t step0,
                      uint64_t* step_counter, const float* noise, const float* uniforms,
                      uint8_t* accept_out, unsigned long long* n_accept, uint32_t* status,
                      void* ws, size_t ws_bytes, void* stream);

/* Fused sampler + energy evaluation (MH.sampler's inner loop mcmc.py:76-81 followed by Hamiltonian.local_energy on the
 * new sample, train.py:128-130): n_steps Metropolis steps as psif_mh_steps with the device Philox stream, then one
 * local-energy pass on the resident state.  e_loc[B], logabs_energy[B] (log|psi| as computed by the energy pass: the
 * operand of the score-function loss, train.py:141), accum / status as psif_local_energy.  ws must hold
 * max(psif_workspace_bytes(VALUE), psif_workspace_bytes(ENERGY)). */
int32_t psif_sample_energy(PsifHandle* h, float* x_inout, float* logabs_inout, float* sign_inout, int64_t B,
                           int32_t n_steps, float step_size, int32_t have_logabs, uint64_t seed, uint64_t walker_id0,
                           uint64_t step0, uint64_t* step_counter, unsigned long long* n_accept, float* e_loc,
                           float* logabs_energy, double* accum, uint32_t* status, void* ws, size_t ws_bytes,
                           void* stream);

/* logdet_matmul value (logdet_matmul.py:35-70) for one weight column:
 * phi_up[B][K][nu][nu], phi_dn[B][K][nd][nd], w[K] -> logabs[B], sign[B]. */
int32_t psif_slogdet_multi(const float* phi_up, const float* phi_dn, const float* w, int64_t B,
                           int32_t K, int32_t nu, int32_t nd, float* logabs, float* sign,
                           uint32_t* status, void* stream);

/* LogDetMatmul.backward (logdet_matmul.py:94-120): grad_log[B] -> dx1 (shape of phi_up), dx2 (shape of phi_dn) and the
 * per-walker terms of dw, dw_per_walker[B][K] (the caller sums over B).  Singular values clamped at 1e-6 contribute
 * no gradient, as with torch.clamp in the reference; below the 1e-12 output floor everything is zero. */
int32_t psif_logdet_matmul_grad(const float* phi_up, const float* phi_dn, const float* w, const float* grad_log,
                                int64_t B, int32_t K, int32_t nu, int32_t nd, float* dx1, float* dx2,
                                float* dw_per_walker, void* stream);

/* Backward of the function above (the reference differentiates its backward again, create_graph at :110): with
 * cotangents v1, v2 (shapes of dx1, dx2) and vw[K] of (dx1, dx2, dw) it returns the gradients with respect to
 * grad_log[B], phi_up, phi_dn and (per walker, [B][K]) w. */
int32_t psif_logdet_matmul_grad_grad(const float* phi_up, const float* phi_dn, const float* w, const float* grad_log,
                                     const float* v1, const float* v2, const float* vw, int64_t B, int32_t K,
                                     int32_t nu, int32_t nd, float* d_grad_log, float* d1, float* d2,
                                     float* dw_per_walker, void* stream);

/* Jastrow.forward (jastrow.py:67-87): x[B][N][3] -> out[B]. */
int32_t psif_jastrow(const float* x, int64_t B, int32_t n_up, int32_t n_dn, float alpha_par,
                     float alpha_anti, float* out, void* stream);

/* Potential.potential (hamiltonian.py:15-35; nuclear repulsion added for natom > 1). */
int32_t psif_potential(const float* x, int64_t B, int32_t n_elec, int32_t natom, const double* host_Z,
                       const double* host_R, float* out, void* stream);

/* Fill out[n_walkers][n_per_walker] with the N(0,1) stream psif_mh_steps would draw at `step`
 * (exposed so tests can check the generator against the numpy Philox in oracle/). */
int32_t psif_philox_normal(uint64_t seed, uint64_t walker_id0, uint64_t step, int64_t n_walkers,
                           int32_t n_elec, float* out_normals, float* out_uniform, void* stream);

/* d(sum_b grad_out[b] * log|psi|(x_b)) / d(params): the parameter backward that
 * loss.backward() needs at train.py:148.  grad_params has psif_param_count floats and is
 * OVERWRITTEN.  range_flag (device, one word, may be NULL) gets PSIF_ST_FP16_RANGE OR-ed in when an activation of the
 * forward recompute left fp16's range: the gradient is then invalid and the call must be repeated in
 * PSIF_GEMM_TF32_SPLIT mode. */
int32_t psif_logpsi_backward(PsifHandle* h, const float* x, const float* grad_out, int64_t B,
                             float* grad_params, uint32_t* range_flag, void* ws, size_t ws_bytes, void* stream);
int32_t psif_backward_workspace_bytes(const PsifHandle* h, int64_t B, size_t* out);

/* debug / test hooks: run ONE stage kernel of the pipeline on caller-provided payloads.
 * Payload layout P[b][i][c][e], C = 1 (value) or 3N+2 (energy).  See tests/test_stages_gpu.py. */
int32_t psif_stage_embed(PsifHandle* h, const float* x, int64_t B, int32_t C, float* out, void* stream);
int32_t psif_stage_linear(const float* in, const float* W, const float* bias, const float* residual,
                          int64_t rows, int32_t C, int32_t k_in, int32_t n_out, int32_t gelu,
                          float* out, void* stream);
/* same contract on the tcgen05 split-precision kernel; gemm_mode = PSIF_GEMM_FP16_SPLIT or PSIF_GEMM_TF32_SPLIT;
 * a_packed != 0 (fp16 mode only): `in` holds rows in the packed fp16-pair format of psif_stage_pack, and with gelu == 2
 * `out` is written in that format too; scratch holds 3*n_out*k_in + 4 floats (tf32 W_hi, W_lo, the fp16 halves, the
 * range flag); trace (tools only, may be NULL): device buffer of 2*18*512 int64 that receives a clock64 timeline of
 * cluster 0.  Stateless. */
int32_t psif_stage_linear_tc(const float* in, const float* W, const float* bias, const float* residual,
                             int64_t rows, int32_t C, int32_t k_in, int32_t n_out, int32_t gelu,
                             int32_t gemm_mode, int32_t a_packed, float* out, float* scratch, long long* trace,
                             void* stream);
/* fp32 rows[rows][width] -> the packed fp16 pair that the kernels in front of a tensor-core Linear write instead of fp32
 * in energy mode: per row, width halves h0 = fp16(x) followed by width halves h1 = fp16(2^11 (x - h0)), the same
 * 4*width bytes.  range_flag (device, may be NULL) is set when |x| >= 65504. */
int32_t psif_stage_pack(const float* in, int64_t rows, int32_t width, float* out, uint32_t* range_flag, void* stream);
/* the determinant stage with derivatives (logdet_matmul.py:35-70 under the forward Laplacian) on a caller-provided
 * orbital payload phi[B][N][C][K (n_up + n_dn)], C = tangent channels + 2: up electron i, determinant k, column j at
 * [b][i][c][k n_up + j], down electron i at [b][n_up + i][c][K n_up + k n_dn + j].  No Jastrow, no potential:
 * logabs = log|sum_k w_k det_up det_dn|, grad[B][C - 2], lap[B], e_loc = -(lap + |grad|^2) / 2.  fixup != 0 runs the
 * clamp fix-up kernel behind it, as the energy pass does: walkers with a singular value below 1e-6 then carry the
 * derivatives the reference's torch.clamp defines. */
int32_t psif_stage_det_energy(const float* phi, const float* w, int64_t B, int32_t N, int32_t n_up, int32_t C, int32_t K,
                              int32_t fixup, float* e_loc, float* logabs, float* sign, float* grad, float* lap,
                              uint32_t* status, void* stream);
/* packed != 0: the output is written in the packed fp16-pair format (shapes for which the pipeline does so) */
int32_t psif_stage_layernorm(const float* in, const float* gamma, const float* beta, int64_t tokens,
                             int32_t C, int32_t d, int32_t packed, float* out, void* stream);
int32_t psif_stage_attention(const float* qkv, int64_t B, int32_t N, int32_t C, int32_t d,
                             int32_t n_head, int32_t packed, float* out, void* stream);
/* first-layer attention: qkv5 is the compact payload [B N][5][3 d] (value, the token's own three tangents, Laplacian), out the
 * dense payload [B N][3 N + 2][d]; head_dim 64 only */
int32_t psif_stage_attention_first_layer(const float* qkv5, int64_t B, int32_t N, int32_t d, int32_t n_head, int32_t packed,
                                         float* out, void* stream);
int32_t psif_stage_gelu(const float* in, int64_t tokens, int32_t C, int32_t width, float* out, void* stream);

/* Per-kernel-class device timing for roofline reports (bench.py), per handle: when enabled, CUDA events bracket
 * every launch group of this handle's calls; psif_profile_read synchronises and fills host_out[9][4] =
 * {groups, total ms, algorithmic FLOPs, algorithmic bytes} for the classes
 * gemm, attention, layernorm, gelu, embed, orbital, det, jastrow, mh (in that order). */
int32_t psif_profile_enable(PsifHandle* h, int32_t on);
int32_t psif_profile_read(PsifHandle* h, double* host_out, int32_t n_classes);

/* number of kernels this library has launched since load (all handles, this process) */
int64_t psif_launch_count(void);
const char* psif_last_error(void);
const char* psif_version(void);

#ifdef __cplusplus
}
#endif
#endif /* PSIFORMER_B200_H */
