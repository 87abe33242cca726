/*
 * scorus_b200.h -- C ABI of the B200-native ensemble sampler (libscorus_b200.so).
 *
 * This is the drop-in boundary for scorus's data-parallel MCMC hot path.  The reference
 * (astrojhgu/scorus, pure Rust) has no FFI of its own; each entry point below names the
 * reference function (path:line, relative to the reference root) it replaces, and is what a
 * thin Rust `-sys` crate binds (INTEGRATION.md shows the shim; rust/ holds its source).
 *
 * Conventions
 *  - plain pointers and sizes only; no C++ / torch types.
 *  - every call returns 0 (SCORUS_OK) or a positive status code; nothing unwinds.
 *    Codes 1..4 are the reference's McmcErr variants (src/mcmc/mcmc_errors.rs:1-7) in
 *    declaration order; the reference itself assert!s / panics on them
 *    (src/mcmc/ensemble_sample.rs:130-133, src/mcmc/utils.rs:86,93-95).
 *  - a context is NOT thread-safe (mirrors the reference's `&mut` exclusivity).
 *  - ensembles cross the ABI flat and walker-major, double[n][dim]: row i is walker i,
 *    rung r owns walkers [r*n/nbeta, (r+1)*n/nbeta) (src/mcmc/ensemble_sample.rs:137,181).
 *  - there is no CPU fallback: without a CUDA device scorus_ctx_create fails.
 */
#ifndef SCORUS_B200_H
#define SCORUS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCORUS_B200_ABI_VERSION 1

typedef struct scorus_ctx scorus_ctx;

/* status codes */
enum {
    SCORUS_OK = 0,
    SCORUS_ERR_NWALKERS_IS_ZERO = 1,          /* McmcErr::NWalkersIsZero          */
    SCORUS_ERR_NWALKERS_IS_NOT_EVEN = 2,      /* McmcErr::NWalkersIsNotEven       */
    SCORUS_ERR_NWALKERS_MISMATCHES_NBETA = 3, /* McmcErr::NWalkersMismatchesNBeta */
    SCORUS_ERR_BETA_NOT_IN_DECR_ORD = 4,      /* McmcErr::BetaNotInDecrOrd        */
    SCORUS_ERR_LENGTH_MISMATCH = 5,           /* assert at ensemble_sample.rs:133 */
    SCORUS_ERR_INVALID_ARG = 6,
    SCORUS_ERR_NPHI_ZERO = 7,                 /* `nphi - 1` on usize, ensemble_sample.rs:184 */
    SCORUS_ERR_NO_LOGPROB = 8,                /* no log-density plugin registered */
    SCORUS_ERR_NOT_UPLOADED = 9,
    SCORUS_ERR_DIM_TOO_LARGE = 10,            /* the reference is dimension-generic; here the stretch move needs two 32-row
                                               * tiles in 227 KB of shared memory above dim 256 (dim <= ~440), the
                                               * t-walk and HMC keep a row in registers (dim <= 256) */
    SCORUS_ERR_CUDA = 100,                    /* see scorus_last_error */
    SCORUS_ERR_NO_DEVICE = 101
};

/* log-density plugins: the device-side replacements of the `flogprob: &F` closure
 * (src/mcmc/ensemble_sample.rs:88,102).  params layouts in the comments. */
enum {
    SCORUS_LP_GAUSSIAN = 0,         /* {c}: -sum c*x^2; c=0.5 examples/test_emcee.rs:18-24, c=1 examples/sample_high_dim.rs:22-28 */
    SCORUS_LP_BOUNDED_GAUSSIAN = 1, /* {limit}: examples/sample_bimodal.rs:35-44 */
    SCORUS_LP_ROSENBROCK = 2,       /* {}: examples/test_emcee.rs:26-32 */
    SCORUS_LP_RASTRIGIN = 3,        /* {a}: examples/sample_rastrigin.rs:17-25 */
    SCORUS_LP_BIMODAL2D = 4,        /* {} or {lo0,hi0,lo1,hi1,split,mu_a,sigma_a,mu_b,sigma_b}: examples/sample_bimodal.rs:46-54 */
    SCORUS_LP_STEP2D = 5,           /* {}: examples/sample_bimodal.rs:56-65 */
    SCORUS_LP_CORR_GAUSSIAN = 6,    /* {mu[D], L[D][D] row-major}: -0.5*|L(x-mu)|^2 (not in the reference; BASELINE config 2) */
    SCORUS_LP_MIXTURE = 7,          /* {sigma1, sigma2, m[D]}: two-component mixture (not in the reference; BASELINE config 3) */
    SCORUS_LP_COUNT = 8
};

/* UpdateFlagSpec (src/mcmc/ensemble_sample.rs:25-55) */
enum {
    SCORUS_UFS_PROB = 0,          /* Prob(p): D Bernoulli(p), redrawn until one is set (:43-50) */
    SCORUS_UFS_ALL = 1,           /* All (:51) */
    SCORUS_UFS_ONE_HOT_CYCLE = 2, /* the cycling Func closure of examples/sample_high_dim.rs:59-65 */
    SCORUS_UFS_MASK = 3           /* Func(f) in general: the caller evaluates f() once per walker
                                     on the host and passes the [n][mask_len] byte matrix (:52) */
};
typedef struct scorus_ufs {
    int32_t kind;
    int32_t reserved;
    double prob;          /* SCORUS_UFS_PROB */
    const uint8_t *mask;  /* SCORUS_UFS_MASK: [n][mask_len], nonzero = move */
    size_t mask_len;      /* <= dim; coordinates >= mask_len always move (:68) */
} scorus_ufs;

#define SCORUS_CFG_DEFAULT 0u
/* Evaluate every log-density with one lane per walker, summing left to right exactly like the
 * reference closures (bit-identical log-probs for the transcendental-free targets).  The default
 * cooperative kernel sums per-lane partials in a fixed tree instead: same positions and accept
 * decisions, log-probs equal to a few ulp, about 2x faster. */
#define SCORUS_CFG_SEQUENTIAL_SUM 1u
/* T = f32: the reference is generic over the walkers' scalar type (`T: Float`, src/mcmc/ensemble_sample.rs:95).  An f32
 * context holds the ensemble and the log-probs as float and runs the f32 instantiation of the fused step kernel, of
 * init_logprob and of the ladder swap (all targets but the tensor-core dense Gaussian); host arrays go through the
 * *_f32 entry points below, scalars (a, beta, plugin parameters) stay double in the ABI.  Everything else -- t-walk,
 * HMC, statistics, sharding, host-buffer calls, SEQUENTIAL_SUM -- answers SCORUS_ERR_INVALID_ARG on such a context. */
#define SCORUS_CFG_F32 2u
typedef struct scorus_cfg {
    uint64_t n_walkers; /* ensemble.len() */
    uint32_t dim;       /* ensemble[i].dimension() */
    int32_t device;     /* CUDA device ordinal, -1 = the current device */
    uint64_t seed;      /* key of the counter-based device RNG (replaces `rng: &mut U`) */
    uint32_t flags;     /* SCORUS_CFG_* */
    uint32_t reserved;
} scorus_cfg;

/* ---- library ---- */
int scorus_abi_version(void);
const char *scorus_status_string(int status);
int scorus_device_count(void);

/* Precondition checks, callable without a GPU.
 * scorus_check_sample_args: asserts of src/mcmc/ensemble_sample.rs:127-133, in order.
 * scorus_check_beta_order:  the panic of src/mcmc/utils.rs:91-95 (equal betas tolerated). */
int scorus_check_sample_args(size_t n_walkers, size_t logprob_len, size_t nbeta);
int scorus_check_beta_order(const double *beta_list, size_t nbeta);

/* ---- context: owns the device copy of (ensemble, logprob), i.e. of the `ensemble: &mut [V]` and
 * `cached_logprob: &mut [T]` arguments every reference call takes (src/mcmc/ensemble_sample.rs:89-90, :109-110;
 * src/mcmc/utils.rs:73-74) ---- */
int scorus_ctx_create(scorus_ctx **out, const scorus_cfg *cfg);
void scorus_ctx_destroy(scorus_ctx *ctx);
const char *scorus_last_error(const scorus_ctx *ctx);

/* registers the log-density; replaces the `flogprob: &F` argument of every reference call
 * (src/mcmc/ensemble_sample.rs:78, :88, :108; bounds :102) */
int scorus_set_logprob(scorus_ctx *ctx, int kind, const double *params, size_t nparams);

/* host <-> device: the caller's `&mut [V]` / `&mut [T]` slices (src/mcmc/ensemble_sample.rs:109-110) as flat
 * walker-major arrays.  logprob == NULL on upload runs init_logprob (:77-85) on the device. */
/* scorus_upload is asynchronous on the context's stream when the source is page-locked (scorus_host_alloc): the
 * caller must not reuse or free the buffers before scorus_sync (or any call that returns host data).  Pageable
 * sources are staged by the driver and may be reused on return. */
int scorus_upload(scorus_ctx *ctx, const double *ensemble, const double *logprob);
int scorus_download(scorus_ctx *ctx, double *ensemble, double *logprob);

/* init_logprob, src/mcmc/ensemble_sample.rs:77-85 */
int scorus_init_logprob(scorus_ctx *ctx);

/* sample, src/mcmc/ensemble_sample.rs:87-105 (= sample_pt with beta_list = [1]) */
int scorus_sample(scorus_ctx *ctx, double a, const scorus_ufs *ufs, uint64_t nsteps);

/* sample_pt, src/mcmc/ensemble_sample.rs:107-190: nsteps fused steps on the device copy.
 * Asynchronous on the context's stream; scorus_sync / scorus_download wait. */
int scorus_sample_pt(scorus_ctx *ctx, double a, const scorus_ufs *ufs, const double *beta_list,
                     size_t nbeta, uint64_t nsteps);

/* sample_pt with every random draw supplied by the caller (the parity entry point):
 * partner[n] (ensemble_sample.rs:135-148), z[n] (:149), flags[n][flag_len] bytes (:150-153),
 * u[n] (:185).  accepted_out[n] (may be NULL) receives the accept decisions.  Synchronous. */
int scorus_sample_pt_fixed(scorus_ctx *ctx, const double *beta_list, size_t nbeta,
                           const uint32_t *partner, const double *z, const uint8_t *flags,
                           size_t flag_len, const double *u, uint8_t *accepted_out);

/* draw_z, src/mcmc/utils.rs:33-45: `count` stretch factors g(z) ~ 1/sqrt(z) on [1/a, a) from the
 * device stream (the draws walkers 0..count-1 would get at the current step; the step counter
 * advances by one).  out is a host array.  count <= n_walkers. */
int scorus_draw_z(scorus_ctx *ctx, double a, size_t count, double *out);

/* swap_walkers, src/mcmc/utils.rs:72-112 */
int scorus_swap_walkers(scorus_ctx *ctx, const double *beta_list, size_t nbeta);
/* ... with the draws supplied: jvec[(nbeta-1)][n/nbeta] (utils.rs:97) and r (same shape,
 * :104) in sweep order, stage s = rung nbeta-1-s.  swapped_out same shape or NULL. */
int scorus_swap_walkers_fixed(scorus_ctx *ctx, const double *beta_list, size_t nbeta,
                              const uint32_t *jvec, const double *r, uint8_t *swapped_out);

/* The literal drop-in for the reference's `&mut [V]`, `&mut [T]` call shape: host buffers in,
 * one sample_pt step, host buffers out (upload + step + download on the context's stream).
 * logprob_len is cached_logprob.len() (checked like ensemble_sample.rs:133).
 * This per-step form exists for parity with the reference's call sites, not for speed: every call is a PCIe round
 * trip plus a synchronise (~50 us even for the 16 walkers of examples/test_emcee.rs, where the reference's own CPU
 * step takes 3 us; PCIe-bound at 2^22 walkers).  Drivers should use scorus_sample_pt_host_n with the number of steps
 * between two recorded samples, or keep the ensemble on the device (scorus_upload once, scorus_sample_pt with nsteps,
 * scorus_trace_* / scorus_moments_* for the recording): a one-block ensemble then runs all nsteps in ONE launch.
 * On page-locked buffers (scorus_host_alloc / cudaHostAlloc) only the walkers accepted in the call are written back; with
 * rows of 512 bytes and more the ensemble is not uploaded either: the first step reads the rows out of the caller's buffer
 * over PCIe itself and writes accepted rows straight back, both directions of the link at once. */
int scorus_sample_pt_host(scorus_ctx *ctx, double *ensemble, double *logprob, size_t logprob_len,
                          double a, const scorus_ufs *ufs, const double *beta_list, size_t nbeta);
/* The same with nsteps steps between the upload and the download: what a driver that records every
 * k-th iteration needs (examples/test_emcee.rs:82-84 writes every 10th), one PCIe round trip per k. */
int scorus_sample_pt_host_n(scorus_ctx *ctx, double *ensemble, double *logprob, size_t logprob_len,
                            double a, const scorus_ufs *ufs, const double *beta_list, size_t nbeta,
                            uint64_t nsteps);
/* swap_walkers on host buffers; a logprob_len mismatch is the reference's silent no-op
 * (src/mcmc/utils.rs:88). */
int scorus_swap_walkers_host(scorus_ctx *ctx, double *ensemble, double *logprob,
                             size_t logprob_len, const double *beta_list, size_t nbeta);

/* ---- the ensemble / parallel-tempering t-walk, src/mcmc/twalk.rs ----
 * TWalkParams (:75-83): aw, at, pphi and the CUMULATIVE kernel weights fw (Walk below fw[0], Traverse
 * up to fw[1], :57-62); scorus_twalk_params_default is TWalkParams::new(dim) (:89-110). */
typedef struct scorus_twalk_params {
    double aw, at, pphi;
    double fw[2];
} scorus_twalk_params;
int scorus_twalk_params_default(size_t dim, scorus_twalk_params *out);
/* twalk::sample (:586-648) on the resident ensemble, nsteps times: a random perfect matching per rung, then
 * two half-passes (sample1, :650-762) in which one member of every pair is proposed a Walk (:274-304) or
 * Traverse (:324-353) move against the other and accepted with calc_a (:463-503).  The pairs of rung r are
 * walkers of rung r (the reference's sample1 reads them unshifted, :686-690, and writes them shifted, :757-758;
 * with one rung the two agree -- DESIGN.md 4.7).  Each half-pass advances the step counter by one.  The
 * nthreads argument of the reference has no counterpart.  dim <= 256. */
int scorus_twalk_sample(scorus_ctx *ctx, const scorus_twalk_params *tw, const double *beta_list, size_t nbeta,
                        uint64_t nsteps);
/* One half-pass with every random draw supplied, per pair k < n/2 in matching order: pair k moves walker
 * order[2k + half] against order[2k + 1 - half] (order: a permutation of the walkers, rung by rung);
 * kernel[k] 0 = Walk, 1 = Traverse; b[k] (sim_b, :306-322); flags[k][dim] (:259-272); walk_u[k][dim]
 * (sim_walk's uniforms, flagged coordinates only); u[k] the accept uniform (:756).  accepted_out[k] may be
 * NULL.  Synchronous. */
int scorus_twalk_half_fixed(scorus_ctx *ctx, const scorus_twalk_params *tw, const double *beta_list, size_t nbeta,
                            const uint32_t *order, int half, const uint8_t *kernel, const double *b,
                            const uint8_t *flags, const double *walk_u, const double *u, uint8_t *accepted_out);
/* twalk::sample on caller-owned host buffers (`ensemble_logprob: &mut (W, X)`, :588): upload, nsteps, download */
int scorus_twalk_sample_host(scorus_ctx *ctx, double *ensemble, double *logprob, size_t logprob_len,
                             const scorus_twalk_params *tw, const double *beta_list, size_t nbeta, uint64_t nsteps);
/* the device stream's sim_b value (src/mcmc/twalk.rs:306-322) of walkers [0, count) at step counter `step` (the one draw that goes through
 * pow(); exposed so that a device-stream run can be replayed exactly through scorus_twalk_half_fixed) */
int scorus_twalk_draw_b(scorus_ctx *ctx, double at, uint64_t step, size_t count, double *out);

/* ---- ensemble / parallel-tempering Hamiltonian Monte Carlo, src/mcmc/hmc/naive.rs ----
 * HmcParam (:17-56): after an accepted step epsilon of the rung grows by (1 + adj_factor) with probability
 * 1 - target_accept_ratio, after a rejected one it shrinks by the same factor with probability
 * target_accept_ratio (:269-283).  quick_adj = {t, 0.01}, slow_adj = {t, 1e-6}, fixed = {t, 0}, default {0.99, 1e-6}. */
typedef struct scorus_hmc_param {
    double target_accept_ratio, adj_factor;
} scorus_hmc_param;
/* sample_ensemble_pt (:122-292) on the resident ensemble, nsteps times: fresh N(0, 1) momenta, l leapfrog steps
 * (src/mcmc/hmc/utils.rs:6-30) of size epsilon[rung] under the tempered force beta * grad(logprob), accept with
 * exp(beta (lp' - lp) + K - K') when that is finite.  The gradient -- the reference's second closure, grad_logprob --
 * belongs to the log-density plugin (Gaussian, bounded Gaussian, Rosenbrock, Rastrigin, mixture and the dense Gaussian
 * have one; the bounded 2-D targets answer SCORUS_ERR_INVALID_ARG) and is recomputed from q0 at entry, as sample_ensemble_pt does (:139).  epsilon[nbeta] is
 * in/out (host): for rungs of up to 4096 walkers it is updated exactly as the reference's sequential accept loop
 * would (bit-identical); for larger rungs the same factors are applied at once, epsilon (1 + adj_factor)^(grown -
 * shrunk), within n/nbeta * 2^-53 relative of the sequential result.  accepted_cnt[nbeta] (the reference's return
 * value) may be NULL.  Walkers need not pair up (any n divisible by nbeta).  Synchronous. */
int scorus_hmc_sample_ensemble_pt(scorus_ctx *ctx, double *epsilon, const double *beta_list, size_t nbeta, size_t l,
                                  const scorus_hmc_param *param, uint64_t *accepted_cnt, uint64_t nsteps);
/* one step with the draws supplied: p0[n][dim] momenta (:182-192), u1[n] accept uniforms (:269), u2[n] adaptation
 * uniforms (:275 / :280); accepted_out[n] may be NULL */
int scorus_hmc_sample_ensemble_pt_fixed(scorus_ctx *ctx, double *epsilon, const double *beta_list, size_t nbeta, size_t l,
                                        const scorus_hmc_param *param, const double *p0, const double *u1,
                                        const double *u2, uint8_t *accepted_out, uint64_t *accepted_cnt);
/* sample_ensemble_pt on caller-owned host buffers (q0: &mut [V], lp: &mut [T], src/mcmc/hmc/naive.rs:125-126): upload, one step, download */
int scorus_hmc_sample_ensemble_pt_host(scorus_ctx *ctx, double *q0, double *logprob, size_t logprob_len, double *epsilon,
                                       const double *beta_list, size_t nbeta, size_t l, const scorus_hmc_param *param,
                                       uint64_t *accepted_cnt);
/* #grow - #shrink of the step-size adaptation codes the LAST HMC step wrote for this context's walkers
 * (src/mcmc/hmc/naive.rs:274-283).  Walker shards of one rung run the step with HmcParam::fixed, sum this over the ranks and
 * apply epsilon (1 + adj)^net once: the single-GPU result for rungs above 4096 walkers, bit for bit. */
int scorus_hmc_adapt_net(scorus_ctx *ctx, int64_t *net_out);
/* the device stream's momenta (the draws of src/mcmc/hmc/naive.rs:182-192) of walkers [0, count) at step counter `step`, out[count][dim] (they go through
 * log / sqrt / sincos; exposed so that a device-stream run can be replayed exactly through the _fixed entry point) */
int scorus_hmc_draw_momenta(scorus_ctx *ctx, uint64_t step, size_t count, double *out);

/* ---- periphery of the path ----
 * get_one_init_realization, src/mcmc/init_ensemble.rs:19-32, for every walker on the device:
 * x[d] ~ U[lo[d], hi[d]) from the device stream; runs init_logprob if a log-density is registered. */
int scorus_init_ensemble(scorus_ctx *ctx, const double *lo, const double *hi);
/* mean / cov, src/linear_space/utils.rs:4-15, 17-41, of walkers [first, first+count) (a rung or the
 * whole ensemble) computed on the device: mean_out[dim]; cov_out[dim][dim], biased (/N), about the mean. */
int scorus_mean(scorus_ctx *ctx, size_t first, size_t count, double *mean_out);
int scorus_cov(scorus_ctx *ctx, size_t first, size_t count, double *cov_out);
/* The examples' recording loops (`results[i].push(ensemble[i*n+0][0])`, examples/sample_high_dim.rs:84-86)
 * without a download per step: after every sample_pt step the values ens[walkers[c]][coords[c]] are
 * captured on the device (up to capacity_steps rows).  scorus_trace_download copies min(recorded,
 * max_steps) rows of ncapture doubles, reports how many steps were recorded and rewinds the buffer. */
int scorus_trace_enable(scorus_ctx *ctx, const uint32_t *walkers, const uint32_t *coords, size_t ncapture,
                        size_t capacity_steps);
int scorus_trace_download(scorus_ctx *ctx, double *out, size_t max_steps, size_t *nrecorded);
/* capture after every `every`-th step only (thinned output, examples/test_emcee.rs:82-84); default 1 */
int scorus_trace_stride(scorus_ctx *ctx, uint64_t every);

/* ---- chain statistics on the device: what the example drivers compute from
 * per-step downloads (examples/sample_bimodal.rs:137-161, examples/sample_high_dim.rs:84-108), without
 * the downloads ----
 * Acceptance per rung and exchange acceptance per ladder stage since creation: accepted / proposed have
 * nrungs entries (indexed like beta_list), swapped / swap_proposed nrungs-1 (entry k = exchanges between
 * rung k and rung k+1).  Any output may be NULL.  Synchronises. */
int scorus_get_rung_stats(scorus_ctx *ctx, size_t nrungs, uint64_t *accepted, uint64_t *proposed,
                          uint64_t *swapped, uint64_t *swap_proposed);
/* Running moments of each of nrungs equal rungs, accumulated over the walkers of the rung after every
 * `every`-th step from now on: mean and (with_cov, dim <= 128) the biased covariance as
 * src/linear_space/utils.rs:4-41 define them, over steps x walkers.  nrungs == 0 disables.
 * scorus_moments_download reports nsamples = accumulations x walkers per rung; cov_out may be NULL. */
int scorus_moments_enable(scorus_ctx *ctx, size_t nrungs, uint64_t every, int with_cov);
int scorus_moments_download(scorus_ctx *ctx, size_t rung, double *mean_out, double *cov_out, uint64_t *nsamples);
/* No reference counterpart (the drivers only record the chains, examples/sample_bimodal.rs:137-139).
 * Integrated autocorrelation time of every captured series (scorus_trace_enable), from the rows recorded so
 * far (the buffer is not rewound): C(k) = sum_{t<T-k} (x_t - m)(x_{t+k} - m) / T for k <= max_lag (0 = T-1),
 * tau(M) = 1 + 2 sum_{k=1..M} C(k)/C(0) with Sokal's window, M = the first lag with M >= c_window * tau(M)
 * (max_lag if none: check window_out).  tau_out / window_out / mean_out / var_out (= C(0)) have ncapture
 * entries; all but tau_out may be NULL. */
int scorus_trace_iat(scorus_ctx *ctx, double c_window, size_t max_lag, double *tau_out, uint32_t *window_out,
                     double *mean_out, double *var_out);

/* ---- sharding: one process per GPU.  No reference counterpart (the reference is one
 * process); what makes it exact is that the pairing never leaves a rung (src/mcmc/ensemble_sample.rs:136-146) and
 * that swap_walkers only couples adjacent rungs (src/mcmc/utils.rs:89-108) ----
 * scorus_set_shard tells a context which slice of the whole ensemble it holds: its walker i is
 * global walker walker_offset + i, its rung r is global rung rung_offset + r.  The device streams are
 * keyed by global ids, so a sharded run draws exactly what a single-GPU run of the whole ensemble
 * draws.  With whole rungs per rank (tempered runs) scorus_sample_pt then needs no communication. */
int scorus_set_shard(scorus_ctx *ctx, uint64_t walker_offset, uint32_t rung_offset, uint64_t n_walkers_global);
/* One sample_pt step (src/mcmc/ensemble_sample.rs:107-190) of a walker-sharded ensemble with the
 * reference's whole-rung matching: ens_all / lp_all are DEVICE copies of the whole ensemble
 * ([n_global][pitch], [n_global]; e.g. an NCCL all-gather of every rank's scorus_device_buffers),
 * beta_list is the global list.  Updates this context's shard only; identical to the single-GPU step. */
int scorus_sample_pt_global(scorus_ctx *ctx, double a, const scorus_ufs *ufs, const double *beta_list,
                            size_t nbeta, const double *ens_all, const double *lp_all);
/* swap_walkers (src/mcmc/utils.rs:72-112) decided for the whole ladder from the gathered log-probs
 * (DEVICE, [n_global], updated in place to the post-swap values): src_all[slot] (DEVICE, [n_global])
 * receives the global slot whose walker ends in `slot`.  Every rank computes the same plan; rows are
 * then exchanged by the caller (scorus_b200/distributed.py). */
int scorus_swap_plan_device(scorus_ctx *ctx, const double *beta_list, size_t nbeta, double *lp_all,
                            uint32_t *src_all);

/* One-sided partner exchange (the rows `ensemble[i.1]` of src/mcmc/ensemble_sample.rs:157 that live on another GPU):
 * scorus_ipc_export fills two 64-byte CUDA IPC handles (ensemble,
 * log-probs); after the ranks have exchanged them (any transport), scorus_ipc_attach maps every
 * peer's buffers so that kernels of this context can read partner rows straight over NVLink. */
int scorus_ipc_export(scorus_ctx *ctx, void *handle_ens_64b, void *handle_lp_64b);
int scorus_ipc_attach(scorus_ctx *ctx, int world, int rank, const void *handles_ens, const void *handles_lp);

/* sample_pt of a walker shard with the reference's whole-ensemble matching (src/mcmc/ensemble_sample.rs:135-148), WITHOUT an
 * all-gather and without staging copies: every pair of the matching is owned by the rank holding its even-position
 * member; the owner reads the partner's row where it lives (one-sided loads over NVLink from the peers mapped by
 * scorus_ipc_attach), runs the fused step for BOTH walkers and writes accepted rows / log-probs back to their home
 * ranks.  Each row crosses NVLink at most once per direction per step and nothing is computed twice.  Synchronisation:
 * one flag barrier in peer memory before each step (inside the pair-list kernel) and one after the last step -- no
 * NCCL call, no host synchronisation.  Every rank calls it with the same arguments.  nsteps consecutive steps;
 * identical to the single-GPU steps (streams are keyed by global walker ids).
 * UpdateFlagSpec::Func (SCORUS_UFS_MASK; src/mcmc/ensemble_sample.rs:31,52,150-153): here and in
 * scorus_sample_pt_global the mask covers the WHOLE ensemble, [n_walkers_global][mask_len] indexed by global walker id
 * -- the closure is called once per walker of the whole ensemble in order, so every rank evaluates the same closure
 * n_global times, as the single-GPU run does; the owner of a pair needs both walkers' flags.  One step per call. */
int scorus_sample_pt_peer(scorus_ctx *ctx, double a, const scorus_ufs *ufs, const double *beta_list, size_t nbeta,
                          uint64_t nsteps);
/* scorus_sample_pt_peer on caller-owned host buffers of this rank's shard (the reference-shaped call, collective over
 * the ranks): upload, nsteps one-sided steps, write-back of the shard's accepted walkers (pinned buffers: only those). */
int scorus_sample_pt_peer_host(scorus_ctx *ctx, double *ensemble, double *logprob, size_t logprob_len, double a,
                               const scorus_ufs *ufs, const double *beta_list, size_t nbeta, uint64_t nsteps);
/* The barrier on its own (stream-ordered; every attached rank must call it the same number of times). */
int scorus_peer_barrier(scorus_ctx *ctx);

/* Measurement only (the roofline denominator of the one-sided step): read bandwidth of this GPU from rank `peer`'s
 * mapped ensemble over NVLink, by a copy-engine memcpy, by an SM kernel of 16-byte streaming loads, and by the same
 * loads on whole rows in a scrambled order (the step kernel's access pattern; may be NULL); GB/s.
 * bytes == 0: the whole buffer. */
int scorus_peer_read_bandwidth(scorus_ctx *ctx, int peer, size_t bytes, int iters, double *gbps_copy_engine,
                               double *gbps_sm_loads, double *gbps_random_rows);

/* Blocked matchings.  The reference pairs walkers over the whole rung (src/mcmc/ensemble_sample.rs:135-148).  With
 * nblocks > 1 a "local" step cuts every rung into nblocks contiguous blocks that pair among themselves -- what the
 * shards of a multi-GPU run do on the steps where they exchange nothing; global_every = k >= 1 makes every step whose
 * index is a multiple of k a whole-rung step again (0: never).  The schedule depends on the step index only, so each
 * step's matching is independent of the state and the move stays a valid stretch move.  block_offset shifts the
 * stream keys of the blocks (a shard that is block c of the whole ensemble passes c), so that one GPU with
 * (nblocks = world, global_every = k) and `world` shards running scorus_sample_pt / scorus_sample_pt_peer on the same
 * schedule produce the same chain bit for bit.  Default (1, 0, 0): the reference's matching on every step. */
int scorus_set_matching_blocks(scorus_ctx *ctx, uint32_t nblocks, uint32_t block_offset, uint64_t global_every);

/* Applies a swap plan (the row and log-prob exchanges of src/mcmc/utils.rs:106-107) to this rank's slots with one-sided
 * reads of the peers (scorus_ipc_attach):
 * phase 0 pulls every incoming row into scratch, the caller barriers across ranks, phase 1 commits
 * rows and log-probs.  Equal shards (n_global = world * n_local). */
int scorus_swap_apply_p2p(scorus_ctx *ctx, const uint32_t *src_all, const double *lp_all, int phase);

/* swap_walkers (src/mcmc/utils.rs:72-112) over rung shards in one call: barrier, the whole ladder decided on every rank
 * from log-probs read one-sidedly out of the peers' memory, incoming rows pulled over NVLink, barrier, commit.  No
 * NCCL, no host synchronisation; every rank calls it with the GLOBAL beta list.  Identical to the single-GPU swap. */
int scorus_swap_walkers_peer(scorus_ctx *ctx, const double *beta_list, size_t nbeta);

/* ---- T = f32 (SCORUS_CFG_F32): the calls that carry walker data, with float arrays.  scorus_set_logprob,
 * scorus_init_logprob, scorus_sample, scorus_sample_pt, scorus_swap_walkers, scorus_get_stats, the counters and the
 * stream calls are shared with f64 contexts. ---- */
int scorus_upload_f32(scorus_ctx *ctx, const float *ensemble, const float *logprob);
int scorus_download_f32(scorus_ctx *ctx, float *ensemble, float *logprob);
/* sample_pt with every draw supplied (src/mcmc/ensemble_sample.rs:107-190 at T = f32): as scorus_sample_pt_fixed */
int scorus_sample_pt_fixed_f32(scorus_ctx *ctx, const double *beta_list, size_t nbeta, const uint32_t *partner,
                               const float *z, const uint8_t *flags, size_t flag_len, const float *u,
                               uint8_t *accepted_out);
/* swap_walkers with every draw supplied (src/mcmc/utils.rs:72-112 at T = f32): as scorus_swap_walkers_fixed */
int scorus_swap_walkers_fixed_f32(scorus_ctx *ctx, const double *beta_list, size_t nbeta, const uint32_t *jvec,
                                  const float *r, uint8_t *swapped_out);

/* ---- plumbing (no reference counterpart: the reference call returns when the step is done; these calls are
 * asynchronous on the context's stream unless they return host data) ---- */
int scorus_sync(scorus_ctx *ctx);
/* run on a caller-owned cudaStream_t (e.g. torch's current stream; NULL is CUDA's legacy default
 * stream, as everywhere in CUDA); scorus_reset_stream returns to the context's own stream */
int scorus_set_stream(scorus_ctx *ctx, void *cuda_stream);
int scorus_reset_stream(scorus_ctx *ctx);
void *scorus_get_stream(scorus_ctx *ctx);
/* device pointers of the resident state (the device-side `ensemble` / `cached_logprob`, src/mcmc/ensemble_sample.rs:
 * 109-110): ensemble rows are pitch_elems doubles apart.  Both pointers are valid until the next
 * scorus_swap_walkers on this context and must be read again after it: a ladder swap leaves its rows to the stretch
 * step that follows, which writes every row to its new slot in the context's second buffer (the two buffers change
 * roles); this call -- like every other entry point that looks at the ensemble -- first moves rows and log-probs of a
 * pending swap to their slots, so what it returns is the state after the swap (contexts attached to peers by
 * scorus_ipc_attach keep their buffer). */
int scorus_device_buffers(scorus_ctx *ctx, double **ensemble, size_t *pitch_elems, double **logprob);
/* what is left of `rng: &mut U` (src/mcmc/ensemble_sample.rs:111, src/mcmc/utils.rs:75) on the device:
 * RNG position = (step counter, swap-call counter, one-hot cycle counter); with the seed and a
 * download this is a complete checkpoint */
int scorus_get_counters(const scorus_ctx *ctx, uint64_t *step, uint64_t *swap_call, uint64_t *onehot);
int scorus_set_counters(scorus_ctx *ctx, uint64_t step, uint64_t swap_call, uint64_t onehot);
/* the accept branch of src/mcmc/ensemble_sample.rs:185-188 and the swap branch of src/mcmc/utils.rs:105-107, counted:
 * running totals since creation (synchronises): accepted / proposed moves, accepted / proposed swaps */
int scorus_get_stats(scorus_ctx *ctx, uint64_t *accepted, uint64_t *proposed, uint64_t *swapped,
                     uint64_t *swap_proposed);
/* number of kernels this library has launched on this context since creation */
uint64_t scorus_launch_count(const scorus_ctx *ctx);
/* pinned host memory for callers that want full-speed scorus_*_host / upload / download */
void *scorus_host_alloc(size_t bytes);
void scorus_host_free(void *p);

#ifdef __cplusplus
}
#endif
#endif /* SCORUS_B200_H */
