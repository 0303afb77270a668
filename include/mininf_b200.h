/*
 * mininf_b200 — C-ABI of the B200-native ELBO engine.
 *
 * The reference (tillahoffmann/mininf) has no FFI: its ELBO hot path is
 * `EvidenceLowerBoundLoss.forward` (mininf/nn.py:212-228), which re-executes the Python model
 * under a `LogProbTracer` (mininf/core.py:207-277) and lets torch autograd differentiate it.
 * This header is the boundary a maintainer would bind instead (ctypes stub in INTEGRATION.md):
 * the host traces the model ONCE into flat tables of plain-old-data and hands them, together
 * with borrowed device pointers, to the entry points below. Every function
 *   - is `extern "C"`, takes only pointers / integers / POD structs (no torch types),
 *   - only ENQUEUES work on `stream` (a cudaStream_t passed as void*), never synchronises,
 *   - returns 0 on success or a negative MNF_E_* code; `mnf_last_error()` has the text,
 *   - borrows every pointer for the duration of the enqueued work; it owns no user memory.
 *
 * Vocabulary (the reference's): a *site* is one `mininf.sample(name, dist, shape)` statement
 * (mininf/core.py:300-328); a *latent* site is one the approximation draws
 * (mininf/nn.py:133-145), an *observed* site is one `condition` pinned (mininf/core.py:331-387);
 * a *particle* is one reparameterised draw of all latent sites (the reference uses exactly one,
 * mininf/nn.py:217; S particles := mean over S reference evaluations).
 *
 * Packed layouts (all row-major fp32 unless stated):
 *   z     [S][D]  particle values of all *global* latent sites, D = sum of their element counts
 *   noise [S][D]  the reparameterisation noise of z: eps ~ N(0,1) for Normal sites, the standard
 *                 gamma draw for Gamma sites, the drawn value itself for Beta sites
 *   acc   [S][1+D] fp64 step accumulator: column 0 = sum of scaled log-densities of particle s,
 *                 column 1+d = d(that sum)/d z[s][d]
 */
#ifndef MININF_B200_H_
#define MININF_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MNF_ABI_VERSION 7

/* error codes */
#define MNF_OK 0
#define MNF_E_INVALID (-1)      /* bad argument (null pointer, unsupported enum, bad shape)   */
#define MNF_E_UNSUPPORTED (-2)  /* a shape/family combination this build has no kernel for     */
#define MNF_E_CUDA (-3)         /* a CUDA runtime call failed; text in mnf_last_error()        */

/* distribution families; log-densities restate torch.distributions */
#define MNF_NORMAL 0            /* params: loc, scale            torch/distributions/normal.py:87-103    */
#define MNF_GAMMA 1             /* params: concentration, rate   torch/distributions/gamma.py:89-98      */
#define MNF_BETA 2              /* params: concentration1, concentration0   beta.py:87-91 -> dirichlet.py:90-97 */
#define MNF_BERNOULLI_PROBS 3   /* params: probs                 bernoulli.py:121-125 + utils.py:101-137 */
#define MNF_BERNOULLI_LOGITS 4  /* params: logits                bernoulli.py:121-125                    */
#define MNF_POISSON 5           /* params: rate                  poisson.py:75-79                        */
#define MNF_NUM_FAMILIES 6

/* inverse links applied to a linear predictor */
#define MNF_T_ID 0
#define MNF_T_EXP 1
#define MNF_T_MASK 0x0F         /* transform codes of mnf_adam_t: low bits = MNF_T_*              */
#define MNF_T_FROZEN 0x80       /*   | this flag = the parameter is a constant, never updated     */

/* device status bits, OR-ed into `status[0]` by kernels (checked by the host at its sync point) */
#define MNF_ST_BAD_PARAM 1u     /* a distribution parameter left its arg constraint (e.g. scale<=0, NaN) */
#define MNF_ST_BAD_VALUE 2u     /* an unmasked value left the support of its distribution      */
#define MNF_ST_NONFINITE 4u     /* the loss or a gradient is NaN/Inf                           */
#define MNF_ST_XRANK_TIMEOUT 8u /* a peer rank did not deliver its accumulator within ~2 s      */
#define MNF_ST_RANGE 16u        /* MNF_DENSE_F16: a design-matrix entry is NaN/Inf or >= 2^15: outside */
                                /* the range the fp16 operand format is used for (use MNF_DENSE_TF32)  */

/*
 * One distribution parameter as a scalar link:  value_i = T(A_i + B_i * x_i)  with
 *   A_i = a_const + (a_lat < 0 ? 0 : z[s][a_lat + a_stride * i])
 *   B_i = b_const + (b_lat < 0 ? 0 : z[s][b_lat + b_stride * i])
 *   x_i = x == NULL ? 1       : x[x_stride * i]
 * This closed form covers the user link code of SURVEY §8a row a10 short of matrix products:
 * constants, data tensors, a latent itself (prior sites), `c + d*x`, `exp(a + b*x)`.
 */
typedef struct mnf_link {
  float a_const;
  float b_const;
  int32_t a_lat;
  int32_t b_lat;
  int32_t a_stride;
  int32_t b_stride;
  const float* x;
  int32_t x_stride;
  int32_t transform; /* MNF_T_* */
} mnf_link_t;

/* A latent site of the approximation (mininf/nn.py:100-159); parameters are the CONSTRAINED
 * tensors of the torch distribution, already expanded to `numel` elements. */
typedef struct mnf_latent {
  int32_t family;  /* MNF_NORMAL | MNF_GAMMA | MNF_BETA */
  int32_t numel;
  int32_t offset;  /* first column in z / noise */
  int32_t reserved;
  const float* p0; /* Normal loc   | Gamma concentration | Beta concentration1 */
  const float* p1; /* Normal scale | Gamma rate          | Beta concentration0 */
} mnf_latent_t;

/* A model site evaluated element-wise (everything LogProbTracer.sample records,
 * mininf/core.py:211-245, except matrix-product links). */
typedef struct mnf_site {
  int32_t family;
  int32_t value_lat;    /* >= 0: the value is z[s][value_lat + i] (prior of a latent site)     */
  const float* value;   /* else: observed data, numel elements                                 */
  const uint8_t* mask;  /* optional torch.bool mask (MaskedTensor sites, core.py:231-239,265)  */
  int64_t numel;
  double scale;         /* batch rescaling declared/actual (core.py:267-271), 1 otherwise      */
  mnf_link_t param[2];
} mnf_site_t;

/* A site whose first parameter is a dense linear predictor  eta_i = a + X[i,:] . theta
 * (`X @ theta`, tests/test_mininf.py:11, examples/minibatch.md:33). */
typedef struct mnf_dense_site {
  int32_t family;       /* MNF_NORMAL (loc=eta) | MNF_BERNOULLI_LOGITS (logits=eta) | MNF_POISSON (rate=exp(eta)) */
  int32_t p;            /* features */
  int64_t n_rows;       /* rows held by THIS rank */
  int64_t ldx;          /* row stride of X in elements (>= p) */
  const float* X;       /* [n_rows][ldx] */
  const float* y;       /* [n_rows] observed values */
  const uint8_t* mask;  /* optional row mask */
  int32_t theta_lat;    /* column of theta[0] in z */
  int32_t icpt_lat;     /* column of a latent scalar intercept, or -1 */
  float icpt_const;     /* constant intercept when icpt_lat < 0 */
  int32_t reserved;
  mnf_link_t scale;     /* MNF_NORMAL only: the scale parameter as a scalar link with x == NULL  */
  double weight;        /* batch rescaling (core.py:267-271) */
} mnf_dense_site_t;

/*
 * A per-observation ("row") latent matrix Z [n_rows][p] with a mean-field Normal approximation
 * q(Z) = Normal(loc, scale) and the model sites that touch it
 * (examples/regression-with-feature-uncertainty.md:28-38, widened to p features):
 *   prior     Z_ij ~ Normal(prior_loc, prior_scale)            value is the latent itself
 *   features  X_ij ~ Normal(Z_ij, feat_scale)                  observed, optional
 *   response  y_i  ~ F(T(icpt + Z_i . beta))                   observed, optional; F = Poisson (T = exp),
 *                                                              Normal (T = id) or Bernoulli logits
 * Z is never materialised: particle s draws eps_ij from Philox4x32-10 (key = seed, counter =
 * (offset, i*p + j, s/4)) unless `eps` supplies it ([S][n_rows][p], parity tests). The kernel
 * writes d loss/d loc and d loss/d scale (entropy term included, mean over particles) and adds
 * the sites' log-densities, the entropy of q(Z) and the gradients w.r.t. global latents to acc.
 */
typedef struct mnf_rowlatent {
  int64_t n_rows;
  int32_t p;               /* features per row, 1..32 in this build */
  int32_t resp_family;     /* MNF_POISSON | MNF_NORMAL | MNF_BERNOULLI_LOGITS, or -1 for none */
  const float* loc;        /* [n_rows][p] */
  const float* scale;      /* [n_rows][p], positive */
  float* grad_loc;         /* out [n_rows][p] */
  float* grad_scale;       /* out [n_rows][p] */
  const float* eps;        /* optional external noise [S][n_rows][p] */
  mnf_link_t prior_loc;    /* scalar links (x == NULL): constant or scalar latent */
  mnf_link_t prior_scale;
  const float* feat;       /* observed features [n_rows][p] or NULL */
  mnf_link_t feat_scale;
  const float* resp;       /* observed response [n_rows] or NULL */
  int32_t beta_lat;        /* column of beta[0] in z */
  int32_t icpt_lat;        /* scalar latent intercept column or -1 */
  float icpt_const;
  int32_t resp_transform;  /* MNF_T_* applied to icpt + Z.beta */
  mnf_link_t resp_scale;   /* Normal response only */
} mnf_rowlatent_t;

/* precision modes of the dense sweep */
#define MNF_DENSE_FP32 0   /* SIMT fp32 FMA, any p / S                                          */
#define MNF_DENSE_TF32 1   /* tcgen05 kind::tf32, one evaluation per (particle, observation): the */
                           /* black-box sweep. X rounded to TF32 by the TMA unit, theta as hi + lo */
                           /* TF32 pairs, fp32 accumulate in TMEM; shapes: mnf_dense_tf32_kernel   */
#define MNF_DENSE_TF32_CLOSED_FORM 2 /* as 1, but a Normal site with p <= 64 may be reduced to its  */
                           /* Gram statistics (no per-particle work; csrc/dense_gram.cuh)          */

#define MNF_DENSE_F16 3    /* as 1 with fp16 operands (the same 11-bit significand as TF32, K = 16 per */
                           /* MMA: half the tensor-pipe time; csrc/dense_th.cuh). p == 64, S <= 64;    */
                           /* |X| must stay below 2^15 (checked per step -> MNF_ST_RANGE)              */

/* flags of mnf_site_sweep / mnf_plan_desc_t */
#define MNF_SWEEP_CLOSED_FORM 1u /* allow the data-only sufficient-statistics paths (six sums for   */
                           /* Normal(A + B x, sigma), Chebyshev moments for Poisson(exp(A + B x)))  */
                           /* instead of one evaluation per (particle, element)                     */

typedef struct mnf_device_info {
  int32_t sm_count;
  int32_t cc_major;
  int32_t cc_minor;
  int32_t max_smem_optin;
  int64_t total_mem;
} mnf_device_info_t;

int mnf_abi_version(void);
const char* mnf_last_error(void);
int mnf_device_info(int device, mnf_device_info_t* out);

/* Bytes of scratch `mnf_dense_sweep` / `mnf_site_sweep` need for their per-CTA partial sums. */
size_t mnf_workspace_bytes(int n_particles, int n_latent_total, int device);

/*
 * Reparameterised draw of every global latent site for S particles
 * (FactorizedDistribution.rsample, mininf/nn.py:133-145; Normal.rsample torch normal.py:82-85;
 * Gamma.rsample gamma.py:79-87; Beta.rsample beta.py:84-85).
 *   noise_in != NULL : external noise [S][D] (parity mode: the host drew it with the torch
 *                      generator in the reference's order) and is copied to noise_out;
 *   noise_in == NULL : every site draws from Philox4x32-10 (key = seed, counter = (offset,
 *                      particle*D + column [, rejection round])): Normal eps by Box-Muller, standard
 *                      gammas by Marsaglia-Tsang, Beta as G1 / (G1 + G0).
 * Also zeroes `acc` [S][1+D] for the step and clears nothing else.
 * `offset_dev` (may be NULL): a device-resident counter added to `offset` inside the kernel. A
 * step captured in a CUDA graph passes it together with `mnf_finalize(step_counter=...)`, which
 * increments it, so every replay draws with a new call index although its arguments are frozen.
 */
int mnf_rsample(const mnf_latent_t* latents_dev, int n_latents, int n_particles, int n_latent_total,
                const float* noise_in, uint64_t seed, uint64_t offset, const uint64_t* offset_dev,
                float* z, float* noise_out, double* acc, uint32_t* status, void* stream);

/*
 * Dense-link sweep: for every local row i and particle s evaluates the site's log-density at
 * eta = a_s + X[i,:].theta_s and its score, and accumulates
 *   acc[s][0]               += weight * sum_i log p(y_i | eta_is, ...)
 *   acc[s][1+theta_lat+j]   += weight * sum_i dlogp/deta * X[i,j]
 *   (+ intercept / scale latent columns)
 * X and y are read exactly once per call for all particles and both directions
 * (replaces aten::mv + MvBackward + the elementwise log_prob chain).
 */
/*
 * Which tcgen05 kernel MNF_DENSE_TF32 runs for a dense site of `family` with p features and S
 * particles (host-only query, no device needed): 0 = shape not covered (use MNF_DENSE_FP32),
 * 1 = csrc/dense_tc.cuh (p == 64, S <= 64), 2 = csrc/dense_tcr.cuh (p % 4 == 0, ceil(p/64) chunks within 512 TMEM
 * columns and 227 KB of shared memory; up to 32 particles per sweep, 33..128 particles run as 2-4
 * sweeps over X). Intercepts are supported by both. X must also be 16-byte aligned with
 * ldx % 4 == 0 and fewer than 2^31 rows.
 * 3 = only the Gram path below applies (Normal, p <= 64, p % 4 == 0, more than 128 particles).
 * With MNF_DENSE_TF32_CLOSED_FORM a Normal site with p <= 64 and p % 4 == 0 (with or without a mask,
 * any particle count) is swept by csrc/dense_gram.cuh instead: X'X, X'r0, X'1, sum r0, sum r0^2
 * around the particle mean in one pass over X (tcgen05, TF32 X, exact products), then the closed
 * forms of every particle's log-density and gradients in fp64 - same inputs, outputs and bytes
 * read, no per-particle work. MNF_DENSE_TF32 never takes that path.
 */
int mnf_dense_tf32_kernel(int family, int p, int n_particles);

int mnf_dense_sweep(const mnf_dense_site_t* site, int mode, const float* z, int n_particles,
                    int n_latent_total, double* acc, void* workspace, size_t workspace_bytes,
                    uint32_t* status, void* stream);

/*
 * Element-wise sweep over up to MNF_MAX_FUSED_SITES sites of equal numel whose values are
 * observed data and whose links reference only scalar latents (stride 0): one pass over the
 * data for all particles (LogProbTracer.sample + contribution, core.py:211-273).
 * Normal(A + B x, sigma) and Poisson(exp(A + B x)) sites are reduced to data-only sufficient
 * statistics (six sums / up to 33 Chebyshev moments of x) and cost the same for any number of
 * particles; the Poisson expansion is checked on the device per call and a per-particle kernel
 * takes over when the check fails. `workspace` must hold mnf_workspace_bytes(); with less the
 * call still works but without the moment path / the overlap of the two statistics passes.
 * Uses one internal side stream per device (forked from and joined to `stream` inside the call).
 */
#define MNF_MAX_FUSED_SITES 4
int mnf_site_sweep(const mnf_site_t* sites, int n_sites, const float* z, int n_particles,
                   int n_latent_total, double* acc, void* workspace, size_t workspace_bytes,
                   uint32_t flags, uint32_t* status, void* stream);

/*
 * Small sites (priors of global latents, short observed vectors): any link stride, any value
 * kind; accumulates into acc with fp64 atomics. `sites_dev` is a DEVICE array.
 */
int mnf_small_sites(const mnf_site_t* sites_dev, int n_sites, int64_t max_numel, const float* z,
                    int n_particles, int n_latent_total, double* acc, uint32_t* status,
                    void* stream);

/*
 * Combine: loss = -( mean_s acc[s][0] + entropy(q) )   (mininf/nn.py:226-228, entropy
 * nn.py:124-131) and its gradient w.r.t. the constrained parameters of every latent site through
 * the reparameterisation (pathwise) derivative, plus the analytic entropy gradient.
 *   out[0] = loss; out[1 + offset + i] = dloss/dp0_i; out[1 + D + offset + i] = dloss/dp1_i.
 */
int mnf_finalize(const mnf_latent_t* latents_dev, int n_latents, int n_particles,
                 int n_latent_total, const float* z, const float* noise, const double* acc,
                 int with_entropy, float* out, uint64_t* step_counter, uint32_t* status, void* stream);

/*
 * Row-latent sweep (see mnf_rowlatent_t): one pass over loc / scale / features / response for all
 * particles; `entropy_weight` (1 or 0) switches the entropy of q(Z) on.
 */
int mnf_rowlatent_sweep(const mnf_rowlatent_t* desc, const float* z, int n_particles,
                        int n_latent_total, uint64_t seed, uint64_t offset, const uint64_t* offset_dev,
                        int with_entropy, double* acc, void* workspace, size_t workspace_bytes,
                        uint32_t* status, void* stream);

/*
 * ------------------------------------------------------------------------------------------------
 * One call per step: the host traces the model ONCE, hands the flat tables to
 * mnf_plan_create, and every evaluation of EvidenceLowerBoundLoss.forward (mininf/nn.py:212-228)
 * is one mnf_elbo_fwd_bwd call that enqueues every kernel of the step. A C consumer needs nothing
 * else (tests/c/plan_step.c runs a regression step from host tables without Python).
 * ------------------------------------------------------------------------------------------------
 */
typedef struct mnf_plan mnf_plan_t;   /* opaque; owns only its own host / device copies of the tables */
typedef struct mnf_xrank mnf_xrank_t; /* opaque; peer-memory exchange of the step accumulator        */

typedef struct mnf_plan_desc {
  int32_t n_particles;
  int32_t n_latent_total;              /* D */
  int32_t n_latents;
  int32_t n_dense;
  int32_t n_groups;                    /* fused element-wise sweeps (mnf_site_sweep calls) */
  int32_t n_small_observed;
  int32_t n_small_global;
  int32_t n_rowlatent;
  const mnf_latent_t* latents;         /* HOST arrays below; p0 / p1 / data pointers are device pointers */
  const mnf_dense_site_t* dense;
  const int32_t* dense_mode;           /* MNF_DENSE_* per dense site */
  const mnf_site_t* group_sites;       /* the sites of all groups, concatenated */
  const int32_t* group_sizes;          /* 1..MNF_MAX_FUSED_SITES each */
  const mnf_site_t* small_observed;    /* short observed sites (any link stride) */
  const mnf_site_t* small_global;      /* latent-valued (prior) sites: counted once, after the exchange */
  const mnf_rowlatent_t* rowlatent;    /* loc / scale / grad_* / eps are taken from mnf_buffers_t per step */
  uint32_t flags;                      /* MNF_SWEEP_CLOSED_FORM */
  int32_t device;                      /* -1: the current device */
} mnf_plan_desc_t;

/* per-step pointers of one row latent (read in place; the gradients are written by the kernel) */
typedef struct mnf_row_buffers {
  const float* loc;
  const float* scale;
  float* grad_loc;
  float* grad_scale;
  const float* eps;                    /* optional external noise [S][n_rows][p] */
  /* mnf_svi_step only: the N*p variational parameters are trained INSIDE the row-latent sweep
   * (transform, chain rule and Adam per element, in place; csrc/rowlatent.cuh). `loc_rw` is the
   * unconstrained location (identity transform, == loc), `raw_scale` the unconstrained scale
   * (scale = exp(raw_scale); `scale`, `grad_loc`, `grad_scale` are then unused), m / v the Adam
   * moments, all [n_rows][p]. Single-pass sweeps only (at most 32 particles). */
  float* loc_rw;
  float* raw_scale;
  float* m_loc;
  float* v_loc;
  float* m_scale;
  float* v_scale;
} mnf_row_buffers_t;

/* Caller-owned device buffers of a step; every pointer is borrowed for the enqueued work. */
typedef struct mnf_buffers {
  float* z;                            /* [S][D] */
  float* noise;                        /* [S][D] */
  double* acc;                         /* [S][1+D] */
  float* out;                          /* [1+2D]: loss, d loss/d p0, d loss/d p1 */
  void* workspace;
  size_t workspace_bytes;              /* >= mnf_plan_workspace_bytes */
  uint32_t* status;
  const float* noise_in;               /* optional external noise [S][D] (parity mode) */
  uint64_t* step_counter;              /* optional device counter added to `offset` (CUDA-graph replays) */
  const mnf_row_buffers_t* rows;       /* HOST array, n_rowlatent entries */
  mnf_xrank_t* xrank;                  /* optional: every observed site is this rank's row shard */
} mnf_buffers_t;

/* torch.optim.Adam (no weight decay, no amsgrad) over the UNCONSTRAINED parameters of the packed
 * latent sites, fused into the last kernel of the step: raw[d] is the storage behind p0 of latent
 * column d, raw[D+d] behind p1; `constrained` [2D] receives transform(raw) at the start of the
 * step and must be where the plan's latent table points (p0 = constrained + offset,
 * p1 = constrained + D + offset). README.md:63-69's zero_grad / backward / optimizer.step(). */
typedef struct mnf_adam {
  float lr, beta1, beta2, eps;
  float* raw;
  const uint8_t* transform;            /* device [2D]: MNF_T_ID / MNF_T_EXP, | MNF_T_FROZEN */
  float* m;
  float* v;
  float* constrained;
  int64_t* step;                       /* device: updates applied so far */
} mnf_adam_t;

#define MNF_STEP_ENTROPY 1u            /* add the entropy of the approximation (mininf/nn.py:226) */
#define MNF_STEP_PRE 2u                /* rsample + every sweep over observed sites              */
#define MNF_STEP_POST 4u               /* [exchange] + latent-valued sites + finalize [+ Adam]   */
#define MNF_STEP_ALL (MNF_STEP_PRE | MNF_STEP_POST)

int mnf_plan_create(const mnf_plan_desc_t* desc, mnf_plan_t** out);
/* Same table shapes, new data pointers (the next minibatch); device tables are refreshed on `stream`. */
int mnf_plan_update(mnf_plan_t* plan, const mnf_plan_desc_t* desc, void* stream);
int mnf_plan_workspace_bytes(const mnf_plan_t* plan, size_t* bytes);
/* Kernels of this library the most recent step call on this plan enqueued (0 before the first). */
int mnf_plan_launches(const mnf_plan_t* plan, int* kernels);
int mnf_plan_destroy(mnf_plan_t* plan);

/* Enqueue one ELBO + gradient evaluation. Without MNF_STEP_PRE / MNF_STEP_POST both halves run
 * (a caller that reduces `acc` itself, e.g. with ncclAllReduce, calls PRE, reduces, calls POST). */
int mnf_elbo_fwd_bwd(mnf_plan_t* plan, const mnf_buffers_t* buffers, uint64_t seed, uint64_t offset,
                     uint32_t flags, void* stream);
/* The whole SVI step: parameter transforms + mnf_elbo_fwd_bwd + Adam, 4-6 kernels, no host work. */
int mnf_svi_step(mnf_plan_t* plan, const mnf_buffers_t* buffers, const mnf_adam_t* adam, uint64_t seed,
                 uint64_t offset, uint32_t flags, void* stream);

/*
 * Cross-rank exchange (replaces the ncclAllReduce of SURVEY §8e; csrc/small.cuh): every rank pushes
 * its [S][1+D] accumulator into an inbox on each peer over NVLink and raises a flag; each rank adds
 * the inboxes in rank order inside its tail kernel - bit-identical totals on every rank, no host
 * call, capturable. The inbox is one cudaMalloc allocation shared through CUDA IPC: create it,
 * exchange the MNF_XRANK_HANDLE_BYTES handles of all ranks by any means (torch.distributed
 * all_gather in INTEGRATION.md), connect.
 */
#define MNF_XRANK_HANDLE_BYTES 64
int mnf_xrank_create(int world, int rank, int64_t n_doubles, mnf_xrank_t** out, void* handle_out);
int mnf_xrank_connect(mnf_xrank_t* xr, const void* handles /* world x MNF_XRANK_HANDLE_BYTES */);
int mnf_xrank_destroy(mnf_xrank_t* xr);

/*
 * Batched posterior predictive (`broadcast_samples`, mininf/core.py:548-584): for each of B
 * posterior samples, walk the traced sites in model order; `value` sites are evaluated from their
 * link, sites the samples do not provide are drawn from their distribution. z is [B][n_columns]:
 * the given samples occupy their columns, every site writes its result to out_col .. out_col+numel.
 * One launch for all samples (csrc/predict.cuh); Philox stream (seed, offset).
 */
#define MNF_PRED_DRAW 0
#define MNF_PRED_VALUE 1
typedef struct mnf_pred_site {
  int32_t kind;          /* MNF_PRED_DRAW | MNF_PRED_VALUE (then param[0] / the dense link IS the value) */
  int32_t family;        /* MNF_* of a drawn site */
  int64_t numel;
  int32_t out_col;       /* first column of this site's values in z */
  int32_t transform;     /* dense link only: MNF_T_* applied to icpt + X theta */
  mnf_link_t param[2];   /* scalar links (X == NULL) */
  const float* X;        /* dense link of the first parameter: [numel][ldx], or NULL */
  int64_t ldx;
  int32_t p;
  int32_t theta_lat;
  int32_t icpt_lat;
  float icpt_const;
} mnf_pred_site_t;

int mnf_predictive(const mnf_pred_site_t* sites_dev, int n_sites, int n_samples, int n_columns, float* z,
                   uint64_t seed, uint64_t offset, uint32_t* status, void* stream);

/* Counting scan used by the integer-exact parity checks: out[0]=sum(mask), out[1]=sum(mask*value)
 * as int64 (value must hold integers); mask may be NULL (all ones). */
int mnf_masked_count(const float* value, const uint8_t* mask, int64_t numel, int64_t* out,
                     void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MININF_B200_H_ */
