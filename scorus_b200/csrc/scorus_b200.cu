// scorus_b200.cu -- context + C ABI of libscorus_b200.so (see include/scorus_b200.h).
//
// Host side of the B200-native ensemble sampler: owns the device-resident (ensemble, logprob),
// validates the reference's preconditions, and launches the kernels of step_kernel.cuh and
// swap_kernel.cuh.  There is deliberately no CPU path in this file.
#include "../../include/scorus_b200.h"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "launch.cuh"
#include "stats_kernel.cuh"
#include "swap_kernel.cuh"
#include "tile_common.cuh"
#include "twalk_kernel.cuh"
#include "hmc_kernel.cuh"

using namespace scorus;

struct scorus_ctx {
    int device = 0;
    uint64_t n = 0;
    uint32_t dim = 0, pitch = 0;
    uint64_t seed = 0;
    uint64_t step = 0, swap_call = 0, onehot = 0;
    uint64_t w_off = 0, n_global = 0;  // shard position in the whole ensemble (scorus_set_shard)
    uint32_t rung_off = 0;
    uint64_t launches = 0, proposed = 0, swap_proposed = 0;
    int num_sms = 0, smem_optin = 0;

    cudaStream_t own_stream = nullptr, stream = nullptr;
    double *d_ens = nullptr, *d_lp = nullptr, *d_scratch = nullptr;
    double *d_beta = nullptr;   // the beta list of the current call (one of beta_cache)
    struct BetaEntry { std::vector<double> host; double *dev = nullptr; uint64_t used = 0; };
    std::vector<BetaEntry> beta_cache;  // callers alternate between a few lists (local / global ladder)
    uint64_t beta_clock = 0;
    double *d_params = nullptr; size_t params_cap = 0;
    uint32_t *d_order = nullptr, *d_src = nullptr;
    uint32_t *d_gorder = nullptr; size_t gorder_cap = 0;   // global-matching mode: exact order of the whole ensemble
    uint32_t *d_pairs = nullptr;  size_t pairs_cap = 0;    // ... local pair list
    uint32_t *d_count = nullptr;
    // one-sided partner reads over NVLink (scorus_ipc_attach): peers' ensembles / log-probs mapped here
    int world = 1, rank = 0;
    double *peer_ens[16] = {nullptr}, *peer_lp[16] = {nullptr};
    double **d_peer_ens = nullptr, **d_peer_lp = nullptr;   // device copies of the tables
    double *d_stage = nullptr, *d_stage_lp = nullptr;       // pulled partner rows, one per local walker
    // K6: statistics scratch and the chain trace
    double *d_part = nullptr; size_t part_cap = 0;
    double *d_stat = nullptr; size_t stat_cap = 0;
    uint32_t *d_tr_walker = nullptr, *d_tr_coord = nullptr;
    double *d_trace = nullptr;
    size_t tr_ncap = 0, tr_capacity = 0, tr_count = 0;
    uint64_t tr_every = 1, tr_tick = 0;                     // capture after every tr_every-th step
    // chain statistics: per-rung accepted moves [0, rung_cap) and per-stage accepted exchanges [rung_cap, 2 rung_cap)
    unsigned long long *d_rung_stats = nullptr; size_t rung_cap = 0;
    std::vector<uint64_t> rung_proposed, stage_proposed;
    // running per-rung moments: sums of (x - shift) and of its outer products, taken every mom_every-th step
    double *d_mom = nullptr;   // [shift: rungs*dim][s1: rungs*dim][s2: rungs*dim*dim]
    double *d_iat = nullptr; size_t iat_cap = 0;
    uint32_t mom_rungs = 0; bool mom_cov = false;
    uint64_t mom_every = 1, mom_tick = 0, mom_count = 0;
    bool p2p_pending = false;                                // scorus_sample_pt_p2p_begin without _end
    StepParams p2p_params;
    Geometry p2p_geom;
    bool p2p_all_flags = false;
    unsigned long long *d_stats = nullptr;
    // scratch for caller-supplied streams
    double *d_z = nullptr, *d_u = nullptr;
    uint8_t *d_flags = nullptr; size_t flags_cap = 0;
    uint8_t *d_acc = nullptr;
    uint8_t *d_dirty = nullptr;     // host-buffer call: walkers accepted since the upload
    unsigned long long *d_tile_counter = nullptr;  // dynamic tile scheduling of the register kernel
    unsigned long long tile_ticket = 0;
    bool track_dirty = false;
    uint32_t *d_jinv = nullptr; size_t jinv_cap = 0;
    double *d_r = nullptr;      size_t r_cap = 0;
    uint8_t *d_swapped = nullptr; size_t swapped_cap = 0;
    int8_t *d_hmc_adapt = nullptr;                              // HMC: step-size adaptation codes, one per walker
    double *d_hmc_eps = nullptr; size_t hmc_eps_cap = 0;        // ... step size per rung
    double *d_hmc_p = nullptr; size_t hmc_p_cap = 0;            // ... caller-supplied momenta
    uint8_t *d_tw_kernel = nullptr; size_t tw_kernel_cap = 0;   // t-walk, caller-supplied draws
    double *d_tw_walk = nullptr;    size_t tw_walk_cap = 0;

    int lp_kind = -1;
    uint32_t nparams = 0;
    bool uploaded = false;
    bool sequential = false;  // SCORUS_CFG_SEQUENTIAL_SUM
    std::string err;
};

static int fail(scorus_ctx *c, int code, const char *fmt, ...)
{
    if (c) {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof(buf), fmt, ap);
        va_end(ap);
        c->err = buf;
    }
    return code;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(ctx, SCORUS_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                        __FILE__, __LINE__);                                                       \
    } while (0)

// ------------------------------------------------------------------------------------------
// library-level helpers
// ------------------------------------------------------------------------------------------
extern "C" int scorus_abi_version(void) { return SCORUS_B200_ABI_VERSION; }

extern "C" const char *scorus_status_string(int s)
{
    switch (s) {
    case SCORUS_OK: return "ok";
    case SCORUS_ERR_NWALKERS_IS_ZERO: return "NWalkersIsZero";
    case SCORUS_ERR_NWALKERS_IS_NOT_EVEN: return "NWalkersIsNotEven";
    case SCORUS_ERR_NWALKERS_MISMATCHES_NBETA: return "NWalkersMismatchesNBeta";
    case SCORUS_ERR_BETA_NOT_IN_DECR_ORD: return "BetaNotInDecrOrd";
    case SCORUS_ERR_LENGTH_MISMATCH: return "ensemble and logprob lengths differ";
    case SCORUS_ERR_INVALID_ARG: return "invalid argument";
    case SCORUS_ERR_NPHI_ZERO: return "update flags of a walker are all false";
    case SCORUS_ERR_NO_LOGPROB: return "no log-density registered";
    case SCORUS_ERR_NOT_UPLOADED: return "no ensemble uploaded";
    case SCORUS_ERR_DIM_TOO_LARGE: return "dimension too large for one shared-memory tile";
    case SCORUS_ERR_CUDA: return "CUDA error";
    case SCORUS_ERR_NO_DEVICE: return "no CUDA device";
    default: return "unknown status";
    }
}

extern "C" int scorus_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" int scorus_check_sample_args(size_t n, size_t lp_len, size_t nbeta)
{
    // src/mcmc/ensemble_sample.rs:127-133, in assert order
    if (nbeta == 0) return SCORUS_ERR_INVALID_ARG;  // division by zero in the reference
    size_t npb = n / nbeta;
    if (npb * nbeta != n) return SCORUS_ERR_NWALKERS_MISMATCHES_NBETA;
    if (n == 0) return SCORUS_ERR_NWALKERS_IS_ZERO;
    if (npb % 2 != 0) return SCORUS_ERR_NWALKERS_IS_NOT_EVEN;
    if (n != lp_len) return SCORUS_ERR_LENGTH_MISMATCH;
    return SCORUS_OK;
}

extern "C" int scorus_check_beta_order(const double *beta, size_t nbeta)
{
    // src/mcmc/utils.rs:91-95: panic if beta[i] > beta[i-1]; equal values pass
    if (!beta && nbeta) return SCORUS_ERR_INVALID_ARG;
    for (size_t i = 1; i < nbeta; ++i)
        if (beta[i] > beta[i - 1]) return SCORUS_ERR_BETA_NOT_IN_DECR_ORD;
    return SCORUS_OK;
}

extern "C" void *scorus_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
extern "C" void scorus_host_free(void *p)
{
    if (p) cudaFreeHost(p);
}

// ------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------
extern "C" int scorus_ctx_create(scorus_ctx **out, const scorus_cfg *cfg)
{
    if (!out || !cfg) return SCORUS_ERR_INVALID_ARG;
    *out = nullptr;
    if (cfg->n_walkers == 0) return SCORUS_ERR_NWALKERS_IS_ZERO;
    // pairing moves (stretch, t-walk) check the parity of a rung per call (scorus_check_sample_args); HMC has no pairs
    if (cfg->dim == 0 || cfg->n_walkers > 0xFFFFFFFFull) return SCORUS_ERR_INVALID_ARG;
    int ndev = scorus_device_count();
    if (ndev <= 0) return SCORUS_ERR_NO_DEVICE;  // no CPU fallback, by design
    scorus_ctx *ctx = new scorus_ctx();
    int dev = cfg->device;
    cudaError_t e;
    if (dev < 0) {
        e = cudaGetDevice(&dev);
        if (e != cudaSuccess) { delete ctx; return SCORUS_ERR_CUDA; }
    }
    if (dev >= ndev) { delete ctx; return SCORUS_ERR_INVALID_ARG; }
    ctx->device = dev;
    ctx->n = cfg->n_walkers;
    ctx->dim = cfg->dim;
    ctx->pitch = cfg->dim + (cfg->dim & 1u);  // rows are 16-byte multiples for cp.async.bulk
    ctx->seed = cfg->seed;
    ctx->n_global = cfg->n_walkers;
    ctx->sequential = (cfg->flags & SCORUS_CFG_SEQUENTIAL_SUM) != 0;
    auto bail = [&](cudaError_t err, const char *what) {
        fprintf(stderr, "scorus_ctx_create: %s: %s\n", what, cudaGetErrorString(err));
        scorus_ctx_destroy(ctx);
        return SCORUS_ERR_CUDA;
    };
    if ((e = cudaSetDevice(dev)) != cudaSuccess) return bail(e, "cudaSetDevice");
    cudaDeviceGetAttribute(&ctx->num_sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&ctx->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if ((e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking)) != cudaSuccess)
        return bail(e, "cudaStreamCreate");
    ctx->stream = ctx->own_stream;
    size_t ens_bytes = (size_t)ctx->n * ctx->pitch * sizeof(double);
    if ((e = cudaMalloc(&ctx->d_ens, ens_bytes)) != cudaSuccess) return bail(e, "cudaMalloc ensemble");
    if ((e = cudaMalloc(&ctx->d_lp, ctx->n * sizeof(double))) != cudaSuccess) return bail(e, "cudaMalloc logprob");
    if ((e = cudaMalloc(&ctx->d_stats, 4 * sizeof(unsigned long long))) != cudaSuccess) return bail(e, "cudaMalloc stats");
    if ((e = cudaMalloc(&ctx->d_tile_counter, sizeof(unsigned long long))) != cudaSuccess) return bail(e, "cudaMalloc tile counter");
    cudaMemsetAsync(ctx->d_tile_counter, 0, sizeof(unsigned long long), ctx->stream);
    cudaMemsetAsync(ctx->d_ens, 0, ens_bytes, ctx->stream);
    cudaMemsetAsync(ctx->d_stats, 0, 4 * sizeof(unsigned long long), ctx->stream);
    *out = ctx;
    return SCORUS_OK;
}

extern "C" void scorus_ctx_destroy(scorus_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    void *bufs[] = {ctx->d_ens, ctx->d_lp, ctx->d_scratch, ctx->d_params, ctx->d_order,
                    ctx->d_src, ctx->d_stats, ctx->d_z, ctx->d_u, ctx->d_flags, ctx->d_acc,
                    ctx->d_jinv, ctx->d_r, ctx->d_swapped, ctx->d_gorder, ctx->d_pairs, ctx->d_count, ctx->d_dirty, ctx->d_tile_counter};
    for (void *b : bufs)
        if (b) cudaFree(b);
    for (auto &e : ctx->beta_cache)
        if (e.dev) cudaFree(e.dev);
    for (int q = 0; q < ctx->world; ++q)
        if (q != ctx->rank) {
            if (ctx->peer_ens[q]) cudaIpcCloseMemHandle(ctx->peer_ens[q]);
            if (ctx->peer_lp[q]) cudaIpcCloseMemHandle(ctx->peer_lp[q]);
        }
    void *p2p[] = {ctx->d_peer_ens, ctx->d_peer_lp, ctx->d_stage, ctx->d_stage_lp, ctx->d_part, ctx->d_stat,
                   ctx->d_tr_walker, ctx->d_tr_coord, ctx->d_trace, ctx->d_rung_stats, ctx->d_mom, ctx->d_iat, ctx->d_tw_kernel, ctx->d_tw_walk,
                   ctx->d_hmc_adapt, ctx->d_hmc_eps, ctx->d_hmc_p};
    for (void *b : p2p)
        if (b) cudaFree(b);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

extern "C" const char *scorus_last_error(const scorus_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

extern "C" int scorus_set_stream(scorus_ctx *ctx, void *s)
{
    if (!ctx) return SCORUS_ERR_INVALID_ARG;
    cudaSetDevice(ctx->device);
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->stream = (cudaStream_t)s;
    return SCORUS_OK;
}
extern "C" int scorus_reset_stream(scorus_ctx *ctx)
{
    if (!ctx) return SCORUS_ERR_INVALID_ARG;
    cudaSetDevice(ctx->device);
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->stream = ctx->own_stream;
    return SCORUS_OK;
}
extern "C" void *scorus_get_stream(scorus_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

extern "C" int scorus_sync(scorus_ctx *ctx)
{
    if (!ctx) return SCORUS_ERR_INVALID_ARG;
    cudaSetDevice(ctx->device);
    CU(cudaStreamSynchronize(ctx->stream));
    return SCORUS_OK;
}

extern "C" int scorus_device_buffers(scorus_ctx *ctx, double **ens, size_t *pitch, double **lp)
{
    if (!ctx) return SCORUS_ERR_INVALID_ARG;
    if (ens) *ens = ctx->d_ens;
    if (pitch) *pitch = ctx->pitch;
    if (lp) *lp = ctx->d_lp;
    return SCORUS_OK;
}

extern "C" int scorus_get_counters(const scorus_ctx *ctx, uint64_t *step, uint64_t *swap_call, uint64_t *onehot)
{
    if (!ctx) return SCORUS_ERR_INVALID_ARG;
    if (step) *step = ctx->step;
    if (swap_call) *swap_call = ctx->swap_call;
    if (onehot) *onehot = ctx->onehot;
    return SCORUS_OK;
}
extern "C" int scorus_set_counters(scorus_ctx *ctx, uint64_t step, uint64_t swap_call, uint64_t onehot)
{
    if (!ctx) return SCORUS_ERR_INVALID_ARG;
    ctx->step = step;
    ctx->swap_call = swap_call;
    ctx->onehot = onehot % ctx->dim;
    return SCORUS_OK;
}

extern "C" uint64_t scorus_launch_count(const scorus_ctx *ctx) { return ctx ? ctx->launches : 0; }

extern "C" int scorus_get_stats(scorus_ctx *ctx, uint64_t *accepted, uint64_t *proposed, uint64_t *swapped,
                                uint64_t *swap_proposed)
{
    if (!ctx) return SCORUS_ERR_INVALID_ARG;
    cudaSetDevice(ctx->device);
    unsigned long long h[4];
    CU(cudaMemcpyAsync(h, ctx->d_stats, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (accepted) *accepted = h[0];
    if (proposed) *proposed = ctx->proposed;
    if (swapped) *swapped = h[2];
    if (swap_proposed) *swap_proposed = ctx->swap_proposed;
    return SCORUS_OK;
}

static int ensure(scorus_ctx *ctx, void **buf, size_t *cap, size_t bytes)
{
    if (*buf && *cap >= bytes) return SCORUS_OK;
    if (*buf) { CU(cudaStreamSynchronize(ctx->stream)); cudaFree(*buf); *buf = nullptr; }
    CU(cudaMalloc(buf, bytes ? bytes : 16));
    *cap = bytes;
    return SCORUS_OK;
}

extern "C" int scorus_set_logprob(scorus_ctx *ctx, int kind, const double *params, size_t nparams)
{
    if (!ctx) return SCORUS_ERR_INVALID_ARG;
    if (kind < 0 || kind >= SCORUS_LP_COUNT) return fail(ctx, SCORUS_ERR_INVALID_ARG, "unknown log-density kind %d", kind);
    if (nparams && !params) return fail(ctx, SCORUS_ERR_INVALID_ARG, "params is NULL");
    const size_t d = ctx->dim;
    if ((kind == SCORUS_LP_BIMODAL2D || kind == SCORUS_LP_STEP2D) && d < 2)
        return fail(ctx, SCORUS_ERR_INVALID_ARG, "2-D log-density needs dim >= 2");
    if (kind == SCORUS_LP_BIMODAL2D && nparams != 0 && nparams != 9)
        return fail(ctx, SCORUS_ERR_INVALID_ARG, "BIMODAL2D takes 0 or 9 params");
    if (kind == SCORUS_LP_CORR_GAUSSIAN && nparams != d + d * d)
        return fail(ctx, SCORUS_ERR_INVALID_ARG, "CORR_GAUSSIAN takes dim + dim*dim params");
    if (kind == SCORUS_LP_MIXTURE && nparams != 2 + d)
        return fail(ctx, SCORUS_ERR_INVALID_ARG, "MIXTURE takes 2 + dim params");
    cudaSetDevice(ctx->device);
    // dense Gaussian: one extra device-side value after {mu, L} tells the tile kernel that L is upper triangular
    // (a Cholesky-type factor), so that the k loop of column block n0 can start at n0 instead of 0
    const bool corr = kind == SCORUS_LP_CORR_GAUSSIAN;
    double upper = 1.0;
    if (corr)
        for (size_t i = 1; i < d && upper != 0.0; ++i)
            for (size_t j = 0; j < i; ++j)
                if (params[d + i * d + j] != 0.0) { upper = 0.0; break; }
    const size_t ndev = nparams + (corr ? 1 : 0);
    int rc = ensure(ctx, (void **)&ctx->d_params, &ctx->params_cap, (ndev ? ndev : 1) * sizeof(double));
    if (rc) return rc;
    CU(cudaStreamSynchronize(ctx->stream));
    if (nparams) CU(cudaMemcpy(ctx->d_params, params, nparams * sizeof(double), cudaMemcpyHostToDevice));
    if (corr) CU(cudaMemcpy(ctx->d_params + nparams, &upper, sizeof(double), cudaMemcpyHostToDevice));
    ctx->lp_kind = kind;
    ctx->nparams = (uint32_t)ndev;
    return SCORUS_OK;
}

// ------------------------------------------------------------------------------------------
// launch geometry
// ------------------------------------------------------------------------------------------

static int env_int(const char *name, int dflt)
{
    const char *v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

static int step_geometry(scorus_ctx *ctx, Geometry *g)
{
    g->row_bytes = ctx->pitch * 8u;
    g->row_stride = g->row_bytes;
    if (((g->row_stride / 16u) & 1u) == 0) g->row_stride += 16u;  // odd number of 16-byte units
    g->flag_words = (ctx->dim + 31u) / 32u;
    const size_t flag_bytes = ((size_t)g->flag_words * 128u + 127u) & ~(size_t)127u;
    const size_t tile = 32u * (size_t)g->row_stride;
    const size_t budget = (size_t)ctx->smem_optin;
    const int want_stages = env_int("SCORUS_STAGES", 0), want_warps = env_int("SCORUS_WARPS", 0);
    const uint32_t ntiles = (uint32_t)((ctx->n + 31u) / 32u);
    const char *force = getenv("SCORUS_KERNEL");

    // kernel choice: sequential sums on request; else register streaming (dim <= 256); else the
    // cooperative kernel that stages old rows by TMA
    g->kernel = ctx->sequential ? KERNEL_SEQ : (ctx->pitch <= 256 ? KERNEL_REG : KERNEL_COOP);
    if (force && !ctx->sequential) {
        if (!strcmp(force, "coop")) g->kernel = KERNEL_COOP;
        else if (!strcmp(force, "reg") && ctx->pitch <= 256) g->kernel = KERNEL_REG;
        else if (!strcmp(force, "seq")) g->kernel = KERNEL_SEQ;
    }

    if (g->kernel == KERNEL_REG) {
        // one tile of proposed rows + the flag words per warp; at most 14 warps (448 threads)
        // leave SCORUS_L1_KB (default 32) of the SM's unified array to L1: the row loads stream through
        // it, and with the carve-out at its maximum the kernel ran 12 % slower (profiles/r1_notes.md)
        const size_t per_warp = tile + flag_bytes;
        const size_t l1_keep = (size_t)env_int("SCORUS_L1_KB", 32) * 1024u;
        // mirrors reg_kernel_max_warps (step_reg_kernel.cuh): dense plugin 8, rows of <= 8 double2 16, else 14
        uint32_t w = ctx->lp_kind == SCORUS_LP_CORR_GAUSSIAN ? 8 : (ctx->pitch / 2 <= 8 ? 16 : 14);
        while (w >= 1 && w * per_warp > budget) --w;
        while (w > 8 && w * per_warp + l1_keep > budget) --w;
        if (w == 0) return fail(ctx, SCORUS_ERR_DIM_TOO_LARGE, "dim %u: a 32-walker tile does not fit in %d bytes of shared memory", ctx->dim, ctx->smem_optin);
        if (want_warps > 0 && (uint32_t)want_warps < w) w = (uint32_t)want_warps;
        g->stages = 1;
        g->warps = w;
        g->smem = w * per_warp;
    } else {
        auto fit = [&](uint32_t s) {
            const size_t per_warp = s * tile + flag_bytes;
            uint32_t w = 16;
            while (w >= 1 && ((8u * w * s + 127u) & ~127u) + w * per_warp > budget) --w;
            return w;
        };
        uint32_t best_s, best_w;
        if (want_stages >= 2 && want_stages <= 4) {
            best_s = (uint32_t)want_stages;
            best_w = fit(best_s);
        } else {
            // measured on B200 (profiles/r1_notes.md): these kernels are bound by warps per SM, not
            // by prefetch depth, so two stages (more warps) beat three
            best_s = 2;
            best_w = fit(2);
        }
        if (best_w == 0) return fail(ctx, SCORUS_ERR_DIM_TOO_LARGE, "dim %u: a 32-walker tile does not fit in %d bytes of shared memory", ctx->dim, ctx->smem_optin);
        if (want_warps > 0 && (uint32_t)want_warps < best_w) best_w = (uint32_t)want_warps;
        g->stages = best_s;
        g->warps = best_w;
        g->smem = ((8u * g->warps * g->stages + 127u) & ~127u) + (size_t)g->warps * (g->stages * tile + flag_bytes);
    }
    uint32_t grid = (ntiles + g->warps - 1) / g->warps;
    if (grid > (uint32_t)ctx->num_sms) grid = (uint32_t)ctx->num_sms;
    g->grid = grid;
    g->pdl = g->kernel == KERNEL_REG && env_int("SCORUS_PDL", 1);
    return SCORUS_OK;
}

extern "C" int scorus_init_logprob(scorus_ctx *ctx)
{
    if (!ctx) return SCORUS_ERR_INVALID_ARG;
    if (ctx->lp_kind < 0) return fail(ctx, SCORUS_ERR_NO_LOGPROB, "call scorus_set_logprob first");
    cudaSetDevice(ctx->device);
    Geometry g;
    int rc = step_geometry(ctx, &g);
    if (rc) return rc;
    StepParams p;
    memset(&p, 0, sizeof(p));
    p.ens = ctx->d_ens; p.lp = ctx->d_lp; p.pparams = ctx->d_params; p.nparams = ctx->nparams;
    p.n = (uint32_t)ctx->n; p.dim = ctx->dim; p.pitch = ctx->pitch;
    p.row_bytes = g.row_bytes; p.row_stride = g.row_stride;
    p.ntiles = (uint32_t)((ctx->n + 31u) / 32u);
    const size_t tile = 32u * (size_t)g.row_stride;
    uint32_t w = 16;
    while (w > 1 && ((8u * w + 127u) & ~127u) + w * tile > (size_t)ctx->smem_optin) --w;
    p.warps = w;
    size_t smem = ((8u * w + 127u) & ~127u) + w * tile;
    uint32_t grid = (p.ntiles + w - 1) / w;
    // several CTAs per SM are fine here: one stage per warp, latency hidden across CTAs
    uint32_t max_grid = (uint32_t)ctx->num_sms * 4u;
    if (grid > max_grid) grid = max_grid;
    cudaError_t e = launch_init(ctx->lp_kind, p, grid, smem, ctx->stream);
    if (e != cudaSuccess) return fail(ctx, SCORUS_ERR_CUDA, "init_logprob launch: %s", cudaGetErrorString(e));
    ++ctx->launches;
    return SCORUS_OK;
}

// ------------------------------------------------------------------------------------------
// host <-> device
// ------------------------------------------------------------------------------------------
extern "C" int scorus_upload(scorus_ctx *ctx, const double *ens, const double *lp)
{
    if (!ctx || !ens) return SCORUS_ERR_INVALID_ARG;
    cudaSetDevice(ctx->device);
    const size_t rb = (size_t)ctx->dim * sizeof(double);
    CU(cudaMemcpy2DAsync(ctx->d_ens, (size_t)ctx->pitch * sizeof(double), ens, rb, rb, ctx->n,
                         cudaMemcpyHostToDevice, ctx->stream));
    ctx->uploaded = true;
    if (lp) {
        CU(cudaMemcpyAsync(ctx->d_lp, lp, ctx->n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    } else {
        int rc = scorus_init_logprob(ctx);
        if (rc) return rc;
    }
    return SCORUS_OK;
}

extern "C" int scorus_download(scorus_ctx *ctx, double *ens, double *lp)
{
    if (!ctx) return SCORUS_ERR_INVALID_ARG;
    if (!ctx->uploaded) return fail(ctx, SCORUS_ERR_NOT_UPLOADED, "nothing uploaded");
    cudaSetDevice(ctx->device);
    const size_t rb = (size_t)ctx->dim * sizeof(double);
    if (ens)
        CU(cudaMemcpy2DAsync(ens, rb, ctx->d_ens, (size_t)ctx->pitch * sizeof(double), rb, ctx->n,
                             cudaMemcpyDeviceToHost, ctx->stream));
    if (lp) CU(cudaMemcpyAsync(lp, ctx->d_lp, ctx->n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return SCORUS_OK;
}

// ------------------------------------------------------------------------------------------
// sample_pt
// ------------------------------------------------------------------------------------------
static int upload_beta(scorus_ctx *ctx, const double *beta, size_t nbeta)
{
    // Device copies of the last few beta lists are kept so that alternating lists (a rank's own rungs
    // for sample_pt, the whole ladder for the swap plan) never force a stream synchronisation.
    for (auto &e : ctx->beta_cache)
        if (e.host.size() == nbeta && memcmp(e.host.data(), beta, nbeta * sizeof(double)) == 0) {
            e.used = ++ctx->beta_clock;
            ctx->d_beta = e.dev;
            return SCORUS_OK;
        }
    scorus_ctx::BetaEntry *slot = nullptr;
    if (ctx->beta_cache.size() < 8) {
        ctx->beta_cache.reserve(8);  // entries must not move: async copies read e.host
        ctx->beta_cache.emplace_back();
        slot = &ctx->beta_cache.back();
    } else {
        slot = &ctx->beta_cache[0];
        for (auto &e : ctx->beta_cache)
            if (e.used < slot->used) slot = &e;
        CU(cudaStreamSynchronize(ctx->stream));  // earlier launches may still read the evicted list
        cudaFree(slot->dev);
        slot->dev = nullptr;
    }
    slot->host.assign(beta, beta + nbeta);
    CU(cudaMalloc(&slot->dev, nbeta * sizeof(double)));
    CU(cudaMemcpyAsync(slot->dev, slot->host.data(), nbeta * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    slot->used = ++ctx->beta_clock;
    ctx->d_beta = slot->dev;
    return SCORUS_OK;
}

// standard (single-GPU layout) launches of the register kernel claim tiles dynamically; the counter
// advances by exactly ntiles per launch (see tile_common.cuh), so the base is known on the host
static void stamp_tiles(scorus_ctx *ctx, StepParams *p, const Geometry &g)
{
    p->tile_counter = nullptr;
    if (g.kernel == KERNEL_REG && !p->n_dev && env_int("SCORUS_DYNAMIC_TILES", 1)) {
        p->tile_counter = ctx->d_tile_counter;
        p->tile_base = ctx->tile_ticket;
        ctx->tile_ticket += p->ntiles;
    }
}

static uint32_t flag_bits_for(double prob)
{
    const uint32_t cand[5] = {1, 2, 4, 8, 16};
    for (uint32_t b : cand) {
        double t = prob * (double)(1u << b);
        if (t == floor(t)) return b;
    }
    return 32;
}

// chain statistics: device counters per rung (accepted moves) and per exchange stage (accepted swaps),
// grown on demand; the proposal counts are known on the host
static int ensure_rung_stats(scorus_ctx *ctx, size_t nbeta)
{
    if (ctx->rung_cap >= nbeta) return SCORUS_OK;
    size_t cap = ctx->rung_cap ? ctx->rung_cap : 64;
    while (cap < nbeta) cap *= 2;
    unsigned long long *nb = nullptr;
    CU(cudaMalloc(&nb, 2 * cap * sizeof(unsigned long long)));
    CU(cudaMemsetAsync(nb, 0, 2 * cap * sizeof(unsigned long long), ctx->stream));
    if (ctx->d_rung_stats) {
        CU(cudaMemcpyAsync(nb, ctx->d_rung_stats, ctx->rung_cap * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, ctx->stream));
        CU(cudaMemcpyAsync(nb + cap, ctx->d_rung_stats + ctx->rung_cap, ctx->rung_cap * sizeof(unsigned long long),
                           cudaMemcpyDeviceToDevice, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->d_rung_stats);
    }
    ctx->d_rung_stats = nb;
    ctx->rung_cap = cap;
    return SCORUS_OK;
}

// walkers [w_lo, w_hi) of an ensemble with npb walkers per rung were proposed `steps` times
static void count_proposed(scorus_ctx *ctx, size_t nbeta, uint64_t npb, uint64_t w_lo, uint64_t w_hi, uint64_t steps)
{
    ctx->proposed += (w_hi - w_lo) * steps;
    if (ctx->rung_proposed.size() < nbeta) ctx->rung_proposed.resize(nbeta, 0);
    for (uint64_t r = w_lo / npb; r < nbeta && r * npb < w_hi; ++r) {
        const uint64_t lo = std::max(w_lo, r * npb), hi = std::min(w_hi, (r + 1) * npb);
        ctx->rung_proposed[r] += (hi - lo) * steps;
    }
}

static void count_swap_proposed(scorus_ctx *ctx, size_t nbeta, uint64_t npb)
{
    ctx->swap_proposed += (nbeta - 1) * npb;
    if (ctx->stage_proposed.size() < nbeta - 1) ctx->stage_proposed.resize(nbeta - 1, 0);
    for (size_t k = 0; k + 1 < nbeta; ++k) ctx->stage_proposed[k] += npb;
}

static int fill_common(scorus_ctx *ctx, StepParams *p, Geometry *g, size_t nbeta)
{
    int rc = step_geometry(ctx, g);
    if (rc) return rc;
    memset(p, 0, sizeof(*p));
    p->ens = ctx->d_ens; p->lp = ctx->d_lp; p->beta = ctx->d_beta; p->pparams = ctx->d_params;
    p->nparams = ctx->nparams; p->stats = ctx->d_stats; p->seed = ctx->seed;
    p->n = (uint32_t)ctx->n; p->npb = (uint32_t)(ctx->n / nbeta); p->dim = ctx->dim; p->pitch = ctx->pitch;
    p->nbeta = (uint32_t)nbeta;
    p->row_bytes = g->row_bytes; p->row_stride = g->row_stride; p->stages = g->stages; p->warps = g->warps;
    p->flag_words = g->flag_words;
    p->ntiles = (uint32_t)((ctx->n + 31u) / 32u);
    p->bij_bits = bij_bits(p->npb);
    p->debug_identity = env_int("SCORUS_DEBUG_IDENTITY", 0) ? 1u : 0u;
    p->w_off = (uint32_t)ctx->w_off; p->rung_off = ctx->rung_off; p->n_global = ctx->n_global;
    p->src_ens = ctx->d_ens; p->src_lp = ctx->d_lp; p->n_dev = nullptr; p->w_lo = 0; p->w_hi = (uint32_t)ctx->n;
    p->dirty = ctx->track_dirty ? ctx->d_dirty : nullptr;
    // measured (profiles/r1_notes.md): +4 % at 32-D (rows of 256 B: too few bytes in flight per warp to cover DRAM
    // latency), -9 % at 64-D (already at 92 % of the roofline), nothing at 10-D (arithmetic-bound)
    const int pf = env_int("SCORUS_L2_PREFETCH", -1);
    p->l2_prefetch = pf < 0 ? (g->row_bytes >= 128u && g->row_bytes <= 256u ? 1u : 0u) : (pf ? 1u : 0u);
    if (nbeta > 1) {
        if ((rc = ensure_rung_stats(ctx, nbeta))) return rc;
        p->rung_acc = ctx->d_rung_stats;
    }
    return SCORUS_OK;
}

static int accumulate_moments(scorus_ctx *ctx);

// after every step: chain capture (scorus_trace_enable; silently stops when the buffer is full) and the
// running moments (scorus_moments_enable)
static int after_step(scorus_ctx *ctx)
{
    if (ctx->tr_ncap && ctx->tr_count < ctx->tr_capacity && ++ctx->tr_tick >= ctx->tr_every) {
        ctx->tr_tick = 0;
        trace_kernel<<<(unsigned)((ctx->tr_ncap + 127) / 128), 128, 0, ctx->stream>>>(
            ctx->d_ens, ctx->pitch, ctx->d_tr_walker, ctx->d_tr_coord, (uint32_t)ctx->tr_ncap,
            ctx->d_trace + ctx->tr_count * ctx->tr_ncap);
        ++ctx->launches;
        ++ctx->tr_count;
    }
    if (ctx->mom_rungs && ++ctx->mom_tick >= ctx->mom_every) {
        ctx->mom_tick = 0;
        return accumulate_moments(ctx);
    }
    return SCORUS_OK;
}

extern "C" int scorus_sample_pt(scorus_ctx *ctx, double a, const scorus_ufs *ufs, const double *beta,
                                size_t nbeta, uint64_t nsteps)
{
    if (!ctx || !ufs || !beta) return SCORUS_ERR_INVALID_ARG;
    int rc = scorus_check_sample_args(ctx->n, ctx->n, nbeta);
    if (rc) return fail(ctx, rc, "%s", scorus_status_string(rc));
    if (ctx->lp_kind < 0) return fail(ctx, SCORUS_ERR_NO_LOGPROB, "call scorus_set_logprob first");
    if (!ctx->uploaded) return fail(ctx, SCORUS_ERR_NOT_UPLOADED, "call scorus_upload first");
    if (!(a > 1.0)) return fail(ctx, SCORUS_ERR_INVALID_ARG, "stretch scale a must be > 1 (empty Uniform range panics in the reference)");
    cudaSetDevice(ctx->device);

    StepParams p;
    Geometry g;
    if ((rc = upload_beta(ctx, beta, nbeta))) return rc;
    if ((rc = fill_common(ctx, &p, &g, nbeta))) return rc;
    const double sqrt_a = sqrt(a);
    p.inv_sqrt_a = 1.0 / sqrt_a;
    p.z_range = 2.0 * (sqrt_a - p.inv_sqrt_a);  // src/mcmc/utils.rs:42
    p.flag_kind = (uint32_t)ufs->kind;
    bool all_flags = false;
    switch (ufs->kind) {
    case SCORUS_UFS_ALL: all_flags = true; break;
    case SCORUS_UFS_PROB:
        if (!(ufs->prob > 0.0)) return fail(ctx, SCORUS_ERR_INVALID_ARG, "Prob(p) needs p > 0 (the reference would loop forever)");
        p.flag_bits = flag_bits_for(ufs->prob);
        p.flag_thr = p.flag_bits == 32 ? (uint64_t)llrint(fmin(ufs->prob, 1.0) * 4294967296.0)
                                       : (uint64_t)(fmin(ufs->prob, 1.0) * (double)(1u << p.flag_bits));
        break;
    case SCORUS_UFS_ONE_HOT_CYCLE: break;
    case SCORUS_UFS_MASK: {
        if (!ufs->mask || ufs->mask_len > ctx->dim) return fail(ctx, SCORUS_ERR_INVALID_ARG, "bad mask");
        if (nsteps != 1) return fail(ctx, SCORUS_ERR_INVALID_ARG, "a host-evaluated mask covers exactly one step");
        for (size_t i = 0; i < ctx->n; ++i) {
            size_t c = 0;
            for (size_t k = 0; k < ufs->mask_len; ++k) c += ufs->mask[i * ufs->mask_len + k] ? 1 : 0;
            if (c == 0) return fail(ctx, SCORUS_ERR_NPHI_ZERO, "walker %zu has no update flag set", i);
        }
        size_t bytes = (size_t)ctx->n * ufs->mask_len;
        if ((rc = ensure(ctx, (void **)&ctx->d_flags, &ctx->flags_cap, bytes))) return rc;
        CU(cudaMemcpyAsync(ctx->d_flags, ufs->mask, bytes, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        p.flags_in = ctx->d_flags;
        p.flag_len = (uint32_t)ufs->mask_len;
        break;
    }
    default: return fail(ctx, SCORUS_ERR_INVALID_ARG, "unknown UpdateFlagSpec kind %d", ufs->kind);
    }

    bool exact = p.npb <= kExactShuffleMax;
    // one-block ensembles: the register kernel shuffles the rungs itself (step_reg_kernel.cuh)
    const size_t fused_bytes = ctx->n * (sizeof(uint64_t) + sizeof(uint32_t));
    if (exact && g.kernel == KERNEL_REG && g.grid == 1 && g.smem + fused_bytes <= (size_t)ctx->smem_optin &&
        env_int("SCORUS_FUSED_ORDER", 1)) {
        p.fused_order = 1;
        g.smem += fused_bytes;
        exact = false;
    }
    if (exact && !ctx->d_order) CU(cudaMalloc(&ctx->d_order, ctx->n * sizeof(uint32_t)));
    for (uint64_t s = 0; s < nsteps; ++s) {
        p.step = ctx->step;
        p.onehot = ctx->onehot;
        if (exact) {
            exact_order_kernel<<<(unsigned)nbeta, (p.npb + 31u) & ~31u, 0, ctx->stream>>>(ctx->seed, ctx->step, p.npb, ctx->d_order, (uint32_t)ctx->w_off);
            ++ctx->launches;
            p.order = ctx->d_order;
        }
        stamp_tiles(ctx, &p, g);
        cudaError_t e = launch_step(ctx->lp_kind, p, g, all_flags, ctx->stream);
        if (e != cudaSuccess) return fail(ctx, SCORUS_ERR_CUDA, "step launch: %s", cudaGetErrorString(e));
        ++ctx->launches;
        ++ctx->step;
        if (ufs->kind == SCORUS_UFS_ONE_HOT_CYCLE) ctx->onehot = (ctx->onehot + ctx->n_global) % ctx->dim;
        if ((rc = after_step(ctx))) return rc;
    }
    count_proposed(ctx, nbeta, p.npb, 0, ctx->n, nsteps);
    return SCORUS_OK;
}

extern "C" int scorus_sample(scorus_ctx *ctx, double a, const scorus_ufs *ufs, uint64_t nsteps)
{
    const double one = 1.0;  // src/mcmc/ensemble_sample.rs:104
    return scorus_sample_pt(ctx, a, ufs, &one, 1, nsteps);
}

extern "C" int scorus_sample_pt_fixed(scorus_ctx *ctx, const double *beta, size_t nbeta, const uint32_t *partner,
                                      const double *z, const uint8_t *flags, size_t flag_len, const double *u,
                                      uint8_t *accepted_out)
{
    if (!ctx || !beta || !partner || !z || !flags || !u) return SCORUS_ERR_INVALID_ARG;
    int rc = scorus_check_sample_args(ctx->n, ctx->n, nbeta);
    if (rc) return fail(ctx, rc, "%s", scorus_status_string(rc));
    if (ctx->lp_kind < 0) return fail(ctx, SCORUS_ERR_NO_LOGPROB, "call scorus_set_logprob first");
    if (!ctx->uploaded) return fail(ctx, SCORUS_ERR_NOT_UPLOADED, "call scorus_upload first");
    if (flag_len > ctx->dim) return fail(ctx, SCORUS_ERR_INVALID_ARG, "flag_len > dim (index panic in the reference)");
    cudaSetDevice(ctx->device);
    const size_t n = ctx->n, npb = n / nbeta;

    // the pairing must be a fixed-point-free involution inside each rung (ensemble_sample.rs:135-148)
    std::vector<uint32_t> order(n);
    size_t k = 0;
    for (size_t i = 0; i < n; ++i) {
        size_t q = partner[i];
        if (q >= n || q == i || q / npb != i / npb || partner[q] != i)
            return fail(ctx, SCORUS_ERR_INVALID_ARG, "partner[] is not a within-rung involution at walker %zu", i);
        if (i < q) { order[k++] = (uint32_t)i; order[k++] = (uint32_t)q; }
    }
    for (size_t i = 0; i < n; ++i) {
        size_t c = 0;
        for (size_t d = 0; d < flag_len; ++d) c += flags[i * flag_len + d] ? 1 : 0;
        if (c == 0) return fail(ctx, SCORUS_ERR_NPHI_ZERO, "walker %zu has no update flag set", i);
    }

    StepParams p;
    Geometry g;
    if ((rc = upload_beta(ctx, beta, nbeta))) return rc;
    if ((rc = fill_common(ctx, &p, &g, nbeta))) return rc;
    if (!ctx->d_order) CU(cudaMalloc(&ctx->d_order, n * sizeof(uint32_t)));
    if (!ctx->d_z) CU(cudaMalloc(&ctx->d_z, n * sizeof(double)));
    if (!ctx->d_u) CU(cudaMalloc(&ctx->d_u, n * sizeof(double)));
    if (!ctx->d_acc) CU(cudaMalloc(&ctx->d_acc, n));
    if ((rc = ensure(ctx, (void **)&ctx->d_flags, &ctx->flags_cap, n * (flag_len ? flag_len : 1)))) return rc;
    CU(cudaMemcpyAsync(ctx->d_order, order.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_z, z, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_u, u, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_flags, flags, n * flag_len, cudaMemcpyHostToDevice, ctx->stream));
    p.order = ctx->d_order; p.z_in = ctx->d_z; p.u_in = ctx->d_u; p.flags_in = ctx->d_flags;
    p.flag_len = (uint32_t)flag_len; p.flag_kind = SCORUS_UFS_MASK; p.accepted_out = ctx->d_acc;
    p.step = ctx->step;
    stamp_tiles(ctx, &p, g);
    cudaError_t e = launch_step(ctx->lp_kind, p, g, false, ctx->stream);
    if (e != cudaSuccess) return fail(ctx, SCORUS_ERR_CUDA, "step launch: %s", cudaGetErrorString(e));
    ++ctx->launches;
    count_proposed(ctx, nbeta, p.npb, 0, n, 1);
    if (accepted_out) CU(cudaMemcpyAsync(accepted_out, ctx->d_acc, n, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));  // `order` and the caller's arrays are released here
    return SCORUS_OK;
}

// ------------------------------------------------------------------------------------------
// draw_z
// ------------------------------------------------------------------------------------------
__global__ void draw_z_kernel(uint64_t seed, uint64_t step, double inv_sqrt_a, double range, uint32_t count,
                              double *out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) {
        U4 r = draw(seed, step, ST_ZU, i, 0);
        out[i] = z_from_uniform(u53(r.x, r.y), inv_sqrt_a, range);
    }
}

extern "C" int scorus_draw_z(scorus_ctx *ctx, double a, size_t count, double *out)
{
    if (!ctx || !out || count > ctx->n) return SCORUS_ERR_INVALID_ARG;
    if (!(a > 1.0)) return fail(ctx, SCORUS_ERR_INVALID_ARG, "stretch scale a must be > 1");
    cudaSetDevice(ctx->device);
    if (!ctx->d_z) CU(cudaMalloc(&ctx->d_z, ctx->n * sizeof(double)));
    const double sqrt_a = sqrt(a), inv = 1.0 / sqrt_a;
    draw_z_kernel<<<(unsigned)((count + 255) / 256), 256, 0, ctx->stream>>>(ctx->seed, ctx->step, inv,
                                                                            2.0 * (sqrt_a - inv), (uint32_t)count, ctx->d_z);
    ++ctx->launches;
    ++ctx->step;
    CU(cudaMemcpyAsync(out, ctx->d_z, count * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return SCORUS_OK;
}

// ------------------------------------------------------------------------------------------
// swap_walkers
// ------------------------------------------------------------------------------------------
static int swap_impl(scorus_ctx *ctx, const double *beta, size_t nbeta, const uint32_t *jvec, const double *r,
                     uint8_t *swapped_out)
{
    const size_t n = ctx->n;
    if (nbeta == 0) return fail(ctx, SCORUS_ERR_INVALID_ARG, "nbeta == 0");
    const size_t npb = n / nbeta;
    if (npb * nbeta != n) return fail(ctx, SCORUS_ERR_NWALKERS_MISMATCHES_NBETA, "NWalkersMismatchesNBeta");  // utils.rs:86
    int rc = scorus_check_beta_order(beta, nbeta);  // utils.rs:91-95, checked before anything moves
    if (rc) return fail(ctx, rc, "beta list must be in decreasing order");
    if (!ctx->uploaded) return fail(ctx, SCORUS_ERR_NOT_UPLOADED, "call scorus_upload first");
    cudaSetDevice(ctx->device);
    if (nbeta == 1) { ++ctx->swap_call; return SCORUS_OK; }  // for i in (1..1).rev() is empty
    if ((rc = upload_beta(ctx, beta, nbeta))) return rc;
    const size_t m = (nbeta - 1) * npb;

    SwapParams sp;
    memset(&sp, 0, sizeof(sp));
    if (!ctx->d_src) CU(cudaMalloc(&ctx->d_src, n * sizeof(uint32_t)));
    if (!ctx->d_scratch) CU(cudaMalloc(&ctx->d_scratch, n * (size_t)ctx->pitch * sizeof(double)));
    sp.lp = ctx->d_lp; sp.beta = ctx->d_beta; sp.src = ctx->d_src; sp.stats = ctx->d_stats;
    if ((rc = ensure_rung_stats(ctx, nbeta))) return rc;
    sp.stage_swapped = ctx->d_rung_stats + ctx->rung_cap;
    sp.seed = ctx->seed; sp.swap_call = ctx->swap_call;
    sp.npb = (uint32_t)npb; sp.nbeta = (uint32_t)nbeta; sp.bij_bits = bij_bits((uint32_t)npb);

    std::vector<uint32_t> jinv;
    if (jvec) {
        // caller-supplied draws: invert each stage's permutation on the host (j1 = jvec[s][j2])
        jinv.assign(m, 0xFFFFFFFFu);
        for (size_t s = 0; s + 1 < nbeta; ++s)
            for (size_t j2 = 0; j2 < npb; ++j2) {
                uint32_t j1 = jvec[s * npb + j2];
                if (j1 >= npb || jinv[s * npb + j1] != 0xFFFFFFFFu)
                    return fail(ctx, SCORUS_ERR_INVALID_ARG, "jvec stage %zu is not a permutation", s);
                jinv[s * npb + j1] = (uint32_t)j2;
            }
        if ((rc = ensure(ctx, (void **)&ctx->d_jinv, &ctx->jinv_cap, m * sizeof(uint32_t)))) return rc;
        if ((rc = ensure(ctx, (void **)&ctx->d_r, &ctx->r_cap, m * sizeof(double)))) return rc;
        CU(cudaMemcpyAsync(ctx->d_jinv, jinv.data(), m * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(ctx->d_r, r, m * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        sp.jinv = ctx->d_jinv;
        sp.r_in = ctx->d_r;
    } else if (npb <= kExactShuffleMax) {
        if ((rc = ensure(ctx, (void **)&ctx->d_jinv, &ctx->jinv_cap, m * sizeof(uint32_t)))) return rc;
        exact_jvec_kernel<<<(unsigned)(nbeta - 1), ((unsigned)npb + 31u) & ~31u, 0, ctx->stream>>>(
            ctx->seed, ctx->swap_call, (uint32_t)npb, (uint32_t)nbeta, ctx->d_jinv);
        ++ctx->launches;
        sp.jinv = ctx->d_jinv;
    }
    if (swapped_out) {
        if ((rc = ensure(ctx, (void **)&ctx->d_swapped, &ctx->swapped_cap, m))) return rc;
        sp.swapped_out = ctx->d_swapped;
    }
    swap_chain_kernel<<<(unsigned)((npb + 127) / 128), 128, 0, ctx->stream>>>(sp);
    ++ctx->launches;
    const uint32_t vpr = ctx->pitch / 2;
    const size_t total = n * (size_t)vpr;
    unsigned blocks = (unsigned)((total + 255) / 256);
    unsigned cap = (unsigned)ctx->num_sms * 16u;
    if (blocks > cap) blocks = cap;
    permute_rows_kernel<<<blocks, 256, 0, ctx->stream>>>((double2 *)ctx->d_ens, (double2 *)ctx->d_scratch, ctx->d_src,
                                                         (uint32_t)n, vpr, 0);
    permute_rows_kernel<<<blocks, 256, 0, ctx->stream>>>((double2 *)ctx->d_ens, (double2 *)ctx->d_scratch, ctx->d_src,
                                                         (uint32_t)n, vpr, 1);
    ctx->launches += 2;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(ctx, SCORUS_ERR_CUDA, "swap launch: %s", cudaGetErrorString(e));
    ++ctx->swap_call;
    count_swap_proposed(ctx, nbeta, npb);
    if (swapped_out) CU(cudaMemcpyAsync(swapped_out, ctx->d_swapped, m, cudaMemcpyDeviceToHost, ctx->stream));
    if (jvec || swapped_out) CU(cudaStreamSynchronize(ctx->stream));
    return SCORUS_OK;
}

extern "C" int scorus_swap_walkers(scorus_ctx *ctx, const double *beta, size_t nbeta)
{
    if (!ctx || !beta) return SCORUS_ERR_INVALID_ARG;
    return swap_impl(ctx, beta, nbeta, nullptr, nullptr, nullptr);
}

extern "C" int scorus_swap_walkers_fixed(scorus_ctx *ctx, const double *beta, size_t nbeta, const uint32_t *jvec,
                                         const double *r, uint8_t *swapped_out)
{
    if (!ctx || !beta) return SCORUS_ERR_INVALID_ARG;
    if (nbeta > 1 && (!jvec || !r)) return SCORUS_ERR_INVALID_ARG;
    return swap_impl(ctx, beta, nbeta, nbeta > 1 ? jvec : nullptr, r, swapped_out);
}

// ------------------------------------------------------------------------------------------
// sharding: one process per GPU
// ------------------------------------------------------------------------------------------
extern "C" int scorus_set_shard(scorus_ctx *ctx, uint64_t walker_offset, uint32_t rung_offset, uint64_t n_walkers_global)
{
    if (!ctx) return SCORUS_ERR_INVALID_ARG;
    if (walker_offset + ctx->n > n_walkers_global || n_walkers_global > 0xFFFFFFFFull)
        return fail(ctx, SCORUS_ERR_INVALID_ARG, "shard [%llu, %llu) does not fit in %llu walkers",
                    (unsigned long long)walker_offset, (unsigned long long)(walker_offset + ctx->n),
                    (unsigned long long)n_walkers_global);
    ctx->w_off = walker_offset;
    ctx->rung_off = rung_offset;
    ctx->n_global = n_walkers_global;
    return SCORUS_OK;
}

// walker at every matching position of the WHOLE ensemble -> pairs with a member in [w_lo, w_hi)
__global__ void local_pairs_kernel(const StepParams p, uint32_t n_global, uint32_t *pairs, uint32_t *count)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (2u * k + 1u >= n_global) return;
    const uint32_t a = walker_at(p, 2u * k), b = walker_at(p, 2u * k + 1u);
    const bool ina = a >= p.w_lo && a < p.w_hi, inb = b >= p.w_lo && b < p.w_hi;
    if (ina || inb) {
        const uint32_t slot = atomicAdd(count, 2u);
        pairs[slot] = a;
        pairs[slot + 1] = b;
    }
}

extern "C" int scorus_sample_pt_global(scorus_ctx *ctx, double a, const scorus_ufs *ufs, const double *beta,
                                       size_t nbeta, const double *ens_all, const double *lp_all)
{
    if (!ctx || !ufs || !beta || !ens_all || !lp_all) return SCORUS_ERR_INVALID_ARG;
    const size_t ng = ctx->n_global;
    int rc = scorus_check_sample_args(ng, ng, nbeta);
    if (rc) return fail(ctx, rc, "%s", scorus_status_string(rc));
    if (ctx->lp_kind < 0) return fail(ctx, SCORUS_ERR_NO_LOGPROB, "call scorus_set_logprob first");
    if (!ctx->uploaded) return fail(ctx, SCORUS_ERR_NOT_UPLOADED, "call scorus_upload first");
    if (!(a > 1.0)) return fail(ctx, SCORUS_ERR_INVALID_ARG, "stretch scale a must be > 1");
    if (ufs->kind == SCORUS_UFS_MASK) return fail(ctx, SCORUS_ERR_INVALID_ARG, "host-evaluated masks are not supported across shards");
    if (ctx->sequential || ctx->pitch > 256) return fail(ctx, SCORUS_ERR_INVALID_ARG, "global matching needs the register kernel (dim <= 256, no SEQUENTIAL_SUM)");
    cudaSetDevice(ctx->device);

    StepParams p;
    Geometry g;
    if ((rc = upload_beta(ctx, beta, nbeta))) return rc;
    if ((rc = fill_common(ctx, &p, &g, nbeta))) return rc;
    if (g.kernel != KERNEL_REG) return fail(ctx, SCORUS_ERR_INVALID_ARG, "global matching needs the register kernel");
    const double sqrt_a = sqrt(a);
    p.inv_sqrt_a = 1.0 / sqrt_a;
    p.z_range = 2.0 * (sqrt_a - p.inv_sqrt_a);
    p.flag_kind = (uint32_t)ufs->kind;
    bool all_flags = ufs->kind == SCORUS_UFS_ALL;
    if (ufs->kind == SCORUS_UFS_PROB) {
        if (!(ufs->prob > 0.0)) return fail(ctx, SCORUS_ERR_INVALID_ARG, "Prob(p) needs p > 0");
        p.flag_bits = flag_bits_for(ufs->prob);
        p.flag_thr = p.flag_bits == 32 ? (uint64_t)llrint(fmin(ufs->prob, 1.0) * 4294967296.0)
                                       : (uint64_t)(fmin(ufs->prob, 1.0) * (double)(1u << p.flag_bits));
    }
    // ids are global in this mode: no offsets on the stream keys
    p.npb = (uint32_t)(ng / nbeta);
    p.bij_bits = bij_bits(p.npb);
    p.w_off = 0; p.rung_off = 0;
    p.w_lo = (uint32_t)ctx->w_off; p.w_hi = (uint32_t)(ctx->w_off + ctx->n);
    p.step = ctx->step; p.onehot = ctx->onehot;
    p.order = nullptr;
    if (p.npb <= kExactShuffleMax) {  // exact shuffle of every rung of the whole ensemble
        if ((rc = ensure(ctx, (void **)&ctx->d_gorder, &ctx->gorder_cap, ng * sizeof(uint32_t)))) return rc;
        exact_order_kernel<<<(unsigned)nbeta, (p.npb + 31u) & ~31u, 0, ctx->stream>>>(ctx->seed, ctx->step, p.npb, ctx->d_gorder, 0u);
        ++ctx->launches;
        p.order = ctx->d_gorder;
    }
    if ((rc = ensure(ctx, (void **)&ctx->d_pairs, &ctx->pairs_cap, 2 * ctx->n * sizeof(uint32_t)))) return rc;
    if (!ctx->d_count) CU(cudaMalloc(&ctx->d_count, sizeof(uint32_t)));
    CU(cudaMemsetAsync(ctx->d_count, 0, sizeof(uint32_t), ctx->stream));
    local_pairs_kernel<<<(unsigned)((ng / 2 + 255) / 256), 256, 0, ctx->stream>>>(p, (uint32_t)ng, ctx->d_pairs, ctx->d_count);
    ++ctx->launches;
    // the step itself: pair list as the matching order, rows from the gathered copies
    p.order = ctx->d_pairs;
    p.n_dev = ctx->d_count;
    p.n = (uint32_t)(2 * ctx->n);
    p.ntiles = (p.n + 31u) / 32u;
    p.src_ens = ens_all;
    p.src_lp = lp_all;
    cudaError_t e = launch_step(ctx->lp_kind, p, g, all_flags, ctx->stream);
    if (e != cudaSuccess) return fail(ctx, SCORUS_ERR_CUDA, "step launch: %s", cudaGetErrorString(e));
    ++ctx->launches;
    ++ctx->step;
    count_proposed(ctx, nbeta, p.npb, p.w_lo, p.w_hi, 1);
    if (ufs->kind == SCORUS_UFS_ONE_HOT_CYCLE) ctx->onehot = (ctx->onehot + ng) % ctx->dim;
    return after_step(ctx);
}

extern "C" int scorus_swap_plan_device(scorus_ctx *ctx, const double *beta, size_t nbeta, double *lp_all,
                                       uint32_t *src_all)
{
    if (!ctx || !beta || !lp_all || !src_all) return SCORUS_ERR_INVALID_ARG;
    const size_t n = ctx->n_global;
    if (nbeta == 0) return fail(ctx, SCORUS_ERR_INVALID_ARG, "nbeta == 0");
    const size_t npb = n / nbeta;
    if (npb * nbeta != n) return fail(ctx, SCORUS_ERR_NWALKERS_MISMATCHES_NBETA, "NWalkersMismatchesNBeta");
    int rc = scorus_check_beta_order(beta, nbeta);
    if (rc) return fail(ctx, rc, "beta list must be in decreasing order");
    cudaSetDevice(ctx->device);
    if ((rc = upload_beta(ctx, beta, nbeta))) return rc;
    SwapParams sp;
    memset(&sp, 0, sizeof(sp));
    sp.lp = lp_all; sp.beta = ctx->d_beta; sp.src = src_all; sp.stats = ctx->d_stats;
    if ((rc = ensure_rung_stats(ctx, nbeta))) return rc;
    sp.stage_swapped = ctx->d_rung_stats + ctx->rung_cap;
    sp.seed = ctx->seed; sp.swap_call = ctx->swap_call;
    sp.npb = (uint32_t)npb; sp.nbeta = (uint32_t)nbeta; sp.bij_bits = bij_bits((uint32_t)npb);
    if (nbeta > 1 && npb <= kExactShuffleMax) {
        const size_t m = (nbeta - 1) * npb;
        if ((rc = ensure(ctx, (void **)&ctx->d_jinv, &ctx->jinv_cap, m * sizeof(uint32_t)))) return rc;
        exact_jvec_kernel<<<(unsigned)(nbeta - 1), ((unsigned)npb + 31u) & ~31u, 0, ctx->stream>>>(
            ctx->seed, ctx->swap_call, (uint32_t)npb, (uint32_t)nbeta, ctx->d_jinv);
        ++ctx->launches;
        sp.jinv = ctx->d_jinv;
    }
    swap_chain_kernel<<<(unsigned)((npb + 127) / 128), 128, 0, ctx->stream>>>(sp);
    ++ctx->launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(ctx, SCORUS_ERR_CUDA, "swap plan launch: %s", cudaGetErrorString(e));
    ++ctx->swap_call;
    count_swap_proposed(ctx, nbeta, npb);  // whole ladder, as counted by the chain kernel
    return SCORUS_OK;
}

// ------------------------------------------------------------------------------------------
// one-sided partner exchange over NVLink
// ------------------------------------------------------------------------------------------
extern "C" int scorus_ipc_export(scorus_ctx *ctx, void *handle_ens, void *handle_lp)
{
    if (!ctx || !handle_ens || !handle_lp) return SCORUS_ERR_INVALID_ARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaSetDevice(ctx->device);
    CU(cudaIpcGetMemHandle((cudaIpcMemHandle_t *)handle_ens, ctx->d_ens));
    CU(cudaIpcGetMemHandle((cudaIpcMemHandle_t *)handle_lp, ctx->d_lp));
    return SCORUS_OK;
}

extern "C" int scorus_ipc_attach(scorus_ctx *ctx, int world, int rank, const void *handles_ens, const void *handles_lp)
{
    if (!ctx || world < 1 || world > 16 || rank < 0 || rank >= world || !handles_ens || !handles_lp)
        return SCORUS_ERR_INVALID_ARG;
    cudaSetDevice(ctx->device);
    const cudaIpcMemHandle_t *he = (const cudaIpcMemHandle_t *)handles_ens, *hl = (const cudaIpcMemHandle_t *)handles_lp;
    for (int q = 0; q < world; ++q) {
        if (q == rank) { ctx->peer_ens[q] = ctx->d_ens; ctx->peer_lp[q] = ctx->d_lp; continue; }
        void *pe = nullptr, *pl = nullptr;
        CU(cudaIpcOpenMemHandle(&pe, he[q], cudaIpcMemLazyEnablePeerAccess));
        CU(cudaIpcOpenMemHandle(&pl, hl[q], cudaIpcMemLazyEnablePeerAccess));
        ctx->peer_ens[q] = (double *)pe;
        ctx->peer_lp[q] = (double *)pl;
    }
    ctx->world = world;
    ctx->rank = rank;
    if (!ctx->d_peer_ens) CU(cudaMalloc(&ctx->d_peer_ens, 16 * sizeof(double *)));
    if (!ctx->d_peer_lp) CU(cudaMalloc(&ctx->d_peer_lp, 16 * sizeof(double *)));
    CU(cudaMemcpy(ctx->d_peer_ens, ctx->peer_ens, 16 * sizeof(double *), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(ctx->d_peer_lp, ctx->peer_lp, 16 * sizeof(double *), cudaMemcpyHostToDevice));
    if (!ctx->d_stage) CU(cudaMalloc(&ctx->d_stage, ctx->n * (size_t)ctx->pitch * sizeof(double)));
    if (!ctx->d_stage_lp) CU(cudaMalloc(&ctx->d_stage_lp, ctx->n * sizeof(double)));
    return SCORUS_OK;
}

// pulls, for every pair with exactly one local member, the remote member's row and log-prob into the
// stage slot of the local member.  One warp per pair, four pairs in flight per warp.
__global__ void __launch_bounds__(256) pull_partners_kernel(const uint32_t *pairs, const uint32_t *count, uint32_t w_lo,
                                                            uint32_t w_hi, uint32_t n_local, uint32_t pitch,
                                                            double *const *peer_ens, double *const *peer_lp,
                                                            double *stage, double *stage_lp)
{
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t npairs = __ldg(count) >> 1, nvec = pitch >> 1;
    for (uint32_t pr = warp; pr < npairs; pr += nwarps) {
        const uint32_t a = __ldg(pairs + 2 * pr), b = __ldg(pairs + 2 * pr + 1);
        const bool ina = a >= w_lo && a < w_hi, inb = b >= w_lo && b < w_hi;
        if (ina == inb) continue;  // both local: nothing to pull
        const uint32_t local = ina ? a : b, remote = ina ? b : a;
        const uint32_t q = remote / n_local, r = remote - q * n_local;
        const double2 *src = reinterpret_cast<const double2 *>(peer_ens[q] + (size_t)r * pitch);
        double2 *dst = reinterpret_cast<double2 *>(stage + (size_t)(local - w_lo) * pitch);
        for (uint32_t c = lane; c < nvec; c += 128) {  // up to four 16-byte peer loads in flight per lane
            double2 v0 = src[c], v1, v2, v3;
            const bool h1 = c + 32 < nvec, h2 = c + 64 < nvec, h3 = c + 96 < nvec;
            if (h1) v1 = src[c + 32];
            if (h2) v2 = src[c + 64];
            if (h3) v3 = src[c + 96];
            dst[c] = v0;
            if (h1) dst[c + 32] = v1;
            if (h2) dst[c + 64] = v2;
            if (h3) dst[c + 96] = v3;
        }
        if (lane == 0) stage_lp[local - w_lo] = peer_lp[q][r];
    }
}

// sample_pt of a walker shard with the whole-ensemble matching, phase 1: build the local pair list
// and pull the remote partners' rows over NVLink.  Every rank must pass a barrier on the stream
// before (peers finished their previous step) and after (nobody writes before everybody has read).
extern "C" int scorus_sample_pt_p2p_begin(scorus_ctx *ctx, double a, const scorus_ufs *ufs, const double *beta,
                                          size_t nbeta)
{
    if (!ctx || !ufs || !beta) return SCORUS_ERR_INVALID_ARG;
    if (!ctx->d_peer_ens || ctx->world < 1) return fail(ctx, SCORUS_ERR_INVALID_ARG, "call scorus_ipc_attach first");
    if (ctx->p2p_pending) return fail(ctx, SCORUS_ERR_INVALID_ARG, "scorus_sample_pt_p2p_end has not been called");
    const size_t ng = ctx->n_global;
    int rc = scorus_check_sample_args(ng, ng, nbeta);
    if (rc) return fail(ctx, rc, "%s", scorus_status_string(rc));
    if (ng != ctx->n * (size_t)ctx->world) return fail(ctx, SCORUS_ERR_INVALID_ARG, "shards must be equal");
    if (ctx->lp_kind < 0) return fail(ctx, SCORUS_ERR_NO_LOGPROB, "call scorus_set_logprob first");
    if (!ctx->uploaded) return fail(ctx, SCORUS_ERR_NOT_UPLOADED, "call scorus_upload first");
    if (!(a > 1.0)) return fail(ctx, SCORUS_ERR_INVALID_ARG, "stretch scale a must be > 1");
    if (ufs->kind == SCORUS_UFS_MASK) return fail(ctx, SCORUS_ERR_INVALID_ARG, "host-evaluated masks are not supported across shards");
    cudaSetDevice(ctx->device);
    StepParams &p = ctx->p2p_params;
    Geometry &g = ctx->p2p_geom;
    if ((rc = upload_beta(ctx, beta, nbeta))) return rc;
    if ((rc = fill_common(ctx, &p, &g, nbeta))) return rc;
    if (g.kernel != KERNEL_REG) return fail(ctx, SCORUS_ERR_INVALID_ARG, "one-sided matching needs the register kernel (dim <= 256, no SEQUENTIAL_SUM)");
    const double sqrt_a = sqrt(a);
    p.inv_sqrt_a = 1.0 / sqrt_a;
    p.z_range = 2.0 * (sqrt_a - p.inv_sqrt_a);
    p.flag_kind = (uint32_t)ufs->kind;
    ctx->p2p_all_flags = ufs->kind == SCORUS_UFS_ALL;
    if (ufs->kind == SCORUS_UFS_PROB) {
        if (!(ufs->prob > 0.0)) return fail(ctx, SCORUS_ERR_INVALID_ARG, "Prob(p) needs p > 0");
        p.flag_bits = flag_bits_for(ufs->prob);
        p.flag_thr = p.flag_bits == 32 ? (uint64_t)llrint(fmin(ufs->prob, 1.0) * 4294967296.0)
                                       : (uint64_t)(fmin(ufs->prob, 1.0) * (double)(1u << p.flag_bits));
    }
    p.npb = (uint32_t)(ng / nbeta);
    p.bij_bits = bij_bits(p.npb);
    p.w_off = 0; p.rung_off = 0;
    p.w_lo = (uint32_t)ctx->w_off; p.w_hi = (uint32_t)(ctx->w_off + ctx->n);
    p.step = ctx->step; p.onehot = ctx->onehot;
    p.order = nullptr;
    if (p.npb <= kExactShuffleMax) {
        if ((rc = ensure(ctx, (void **)&ctx->d_gorder, &ctx->gorder_cap, ng * sizeof(uint32_t)))) return rc;
        exact_order_kernel<<<(unsigned)nbeta, (p.npb + 31u) & ~31u, 0, ctx->stream>>>(ctx->seed, ctx->step, p.npb, ctx->d_gorder, 0u);
        ++ctx->launches;
        p.order = ctx->d_gorder;
    }
    if ((rc = ensure(ctx, (void **)&ctx->d_pairs, &ctx->pairs_cap, 2 * ctx->n * sizeof(uint32_t)))) return rc;
    if (!ctx->d_count) CU(cudaMalloc(&ctx->d_count, sizeof(uint32_t)));
    CU(cudaMemsetAsync(ctx->d_count, 0, sizeof(uint32_t), ctx->stream));
    local_pairs_kernel<<<(unsigned)((ng / 2 + 255) / 256), 256, 0, ctx->stream>>>(p, (uint32_t)ng, ctx->d_pairs, ctx->d_count);
    pull_partners_kernel<<<(unsigned)ctx->num_sms * 8u, 256, 0, ctx->stream>>>(
        ctx->d_pairs, ctx->d_count, p.w_lo, p.w_hi, (uint32_t)ctx->n, ctx->pitch, ctx->d_peer_ens, ctx->d_peer_lp,
        ctx->d_stage, ctx->d_stage_lp);
    ctx->launches += 2;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(ctx, SCORUS_ERR_CUDA, "pull launch: %s", cudaGetErrorString(e));
    p.order = ctx->d_pairs;
    p.n_dev = ctx->d_count;
    p.n = (uint32_t)(2 * ctx->n);
    p.ntiles = (p.n + 31u) / 32u;
    p.stage_ens = ctx->d_stage;
    p.stage_lp = ctx->d_stage_lp;
    ctx->p2p_pending = true;
    return SCORUS_OK;
}

// phase 2: the fused step on local rows + staged partner rows
extern "C" int scorus_sample_pt_p2p_end(scorus_ctx *ctx)
{
    if (!ctx) return SCORUS_ERR_INVALID_ARG;
    if (!ctx->p2p_pending) return fail(ctx, SCORUS_ERR_INVALID_ARG, "no step pending");
    cudaSetDevice(ctx->device);
    ctx->p2p_pending = false;
    cudaError_t e = launch_step(ctx->lp_kind, ctx->p2p_params, ctx->p2p_geom, ctx->p2p_all_flags, ctx->stream);
    if (e != cudaSuccess) return fail(ctx, SCORUS_ERR_CUDA, "step launch: %s", cudaGetErrorString(e));
    ++ctx->launches;
    ++ctx->step;
    count_proposed(ctx, ctx->p2p_params.nbeta, ctx->p2p_params.npb, ctx->p2p_params.w_lo, ctx->p2p_params.w_hi, 1);
    if (ctx->p2p_params.flag_kind == SCORUS_UFS_ONE_HOT_CYCLE) ctx->onehot = (ctx->onehot + ctx->n_global) % ctx->dim;
    return after_step(ctx);
}

// rows of a cross-rank ladder swap, pulled one-sidedly: scratch[slot] <- (peer holding src)[src] for
// every local slot whose walker changes.  One warp per slot.
__global__ void __launch_bounds__(256) swap_pull_rows_kernel(const uint32_t *src_all, uint32_t w_lo, uint32_t n_local,
                                                             uint32_t pitch, double *const *peer_ens, double *scratch)
{
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t nvec = pitch >> 1;
    for (uint32_t i = warp; i < n_local; i += nwarps) {
        const uint32_t s = __ldg(src_all + w_lo + i);
        if (s == w_lo + i) continue;
        const uint32_t q = s / n_local, r = s - q * n_local;
        const double2 *src = reinterpret_cast<const double2 *>(peer_ens[q] + (size_t)r * pitch);
        double2 *dst = reinterpret_cast<double2 *>(scratch + (size_t)i * pitch);
        for (uint32_t c = lane; c < nvec; c += 32) dst[c] = src[c];
    }
}

__global__ void __launch_bounds__(256) swap_commit_rows_kernel(const uint32_t *src_all, const double *lp_all, uint32_t w_lo,
                                                               uint32_t n_local, uint32_t pitch, const double *scratch,
                                                               double *ens, double *lp)
{
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t nvec = pitch >> 1;
    for (uint32_t i = warp; i < n_local; i += nwarps) {
        if (lane == 0) lp[i] = __ldg(lp_all + w_lo + i);
        if (__ldg(src_all + w_lo + i) == w_lo + i) continue;
        const double2 *src = reinterpret_cast<const double2 *>(scratch + (size_t)i * pitch);
        double2 *dst = reinterpret_cast<double2 *>(ens + (size_t)i * pitch);
        for (uint32_t c = lane; c < nvec; c += 32) dst[c] = src[c];
    }
}

// Applies a swap plan (scorus_swap_plan_device) to this rank's slots with one-sided reads of the peers
// mapped by scorus_ipc_attach.  phase 0 pulls every incoming row into scratch (all ranks still hold
// their old rows); the caller then inserts a stream-ordered barrier; phase 1 writes rows and log-probs.
extern "C" int scorus_swap_apply_p2p(scorus_ctx *ctx, const uint32_t *src_all, const double *lp_all, int phase)
{
    if (!ctx || !src_all || !lp_all) return SCORUS_ERR_INVALID_ARG;
    if (!ctx->d_peer_ens) return fail(ctx, SCORUS_ERR_INVALID_ARG, "call scorus_ipc_attach first");
    cudaSetDevice(ctx->device);
    if (!ctx->d_scratch) CU(cudaMalloc(&ctx->d_scratch, ctx->n * (size_t)ctx->pitch * sizeof(double)));
    const unsigned blocks = (unsigned)ctx->num_sms * 8u;
    if (phase == 0)
        swap_pull_rows_kernel<<<blocks, 256, 0, ctx->stream>>>(src_all, (uint32_t)ctx->w_off, (uint32_t)ctx->n, ctx->pitch,
                                                              ctx->d_peer_ens, ctx->d_scratch);
    else
        swap_commit_rows_kernel<<<blocks, 256, 0, ctx->stream>>>(src_all, lp_all, (uint32_t)ctx->w_off, (uint32_t)ctx->n,
                                                                ctx->pitch, ctx->d_scratch, ctx->d_ens, ctx->d_lp);
    ++ctx->launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(ctx, SCORUS_ERR_CUDA, "swap apply launch: %s", cudaGetErrorString(e));
    return SCORUS_OK;
}

// ------------------------------------------------------------------------------------------
// periphery: box init, ensemble statistics, chain trace (stats_kernel.cuh)
// ------------------------------------------------------------------------------------------
extern "C" int scorus_init_ensemble(scorus_ctx *ctx, const double *lo, const double *hi)
{
    if (!ctx || !lo || !hi) return SCORUS_ERR_INVALID_ARG;
    for (uint32_t d = 0; d < ctx->dim; ++d)
        if (!(lo[d] < hi[d])) return fail(ctx, SCORUS_ERR_INVALID_ARG, "empty range in coordinate %u (Uniform::new panics)", d);
    cudaSetDevice(ctx->device);
    int rc = ensure(ctx, (void **)&ctx->d_stat, &ctx->stat_cap, 2 * (size_t)ctx->dim * sizeof(double));
    if (rc) return rc;
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaMemcpy(ctx->d_stat, lo, ctx->dim * sizeof(double), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(ctx->d_stat + ctx->dim, hi, ctx->dim * sizeof(double), cudaMemcpyHostToDevice));
    init_ensemble_kernel<<<(unsigned)ctx->num_sms * 8u, 256, 0, ctx->stream>>>(
        ctx->d_ens, (uint32_t)ctx->n, ctx->dim, ctx->pitch, ctx->d_stat, ctx->d_stat + ctx->dim, ctx->seed, (uint32_t)ctx->w_off);
    ++ctx->launches;
    ctx->uploaded = true;
    if (ctx->lp_kind >= 0) return scorus_init_logprob(ctx);
    return SCORUS_OK;
}

static int stats_impl(scorus_ctx *ctx, size_t first, size_t count, double *mean_out, double *cov_out)
{
    if (!ctx->uploaded) return fail(ctx, SCORUS_ERR_NOT_UPLOADED, "call scorus_upload first");
    if (count == 0 || first + count > ctx->n) return fail(ctx, SCORUS_ERR_INVALID_ARG, "walker range outside the ensemble");
    const uint32_t dim = ctx->dim, len = dim * dim;
    if (cov_out && dim > (uint32_t)kCovMaxDim) return fail(ctx, SCORUS_ERR_DIM_TOO_LARGE, "cov supports dim <= %d", kCovMaxDim);
    cudaSetDevice(ctx->device);
    uint32_t blocks = (uint32_t)ctx->num_sms * 2u;
    if (blocks > count) blocks = (uint32_t)count;
    int rc = ensure(ctx, (void **)&ctx->d_part, &ctx->part_cap, (size_t)blocks * (cov_out ? len : dim) * sizeof(double));
    if (rc) return rc;
    if ((rc = ensure(ctx, (void **)&ctx->d_stat, &ctx->stat_cap, ((size_t)dim + len) * sizeof(double)))) return rc;
    double *d_mean = ctx->d_stat, *d_cov = ctx->d_stat + dim;
    colsum_partial_kernel<<<blocks, 256, 0, ctx->stream>>>(ctx->d_ens, (uint32_t)first, (uint32_t)count, dim, ctx->pitch, nullptr, ctx->d_part);
    reduce_partials_kernel<<<(dim + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_part, blocks, dim, 1.0 / (double)count, 0, 0, d_mean);
    ctx->launches += 2;
    if (mean_out) CU(cudaMemcpyAsync(mean_out, d_mean, dim * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (cov_out) {
        uint32_t cblocks = (uint32_t)ctx->num_sms;
        if (cblocks > (count + 31) / 32) cblocks = (uint32_t)((count + 31) / 32);
        const size_t smem = 32u * (size_t)dim * sizeof(double);
        cudaError_t e = cudaFuncSetAttribute(cov_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail(ctx, SCORUS_ERR_CUDA, "cov smem: %s", cudaGetErrorString(e));
        cov_partial_kernel<<<cblocks, 256, smem, ctx->stream>>>(ctx->d_ens, (uint32_t)first, (uint32_t)count, dim, ctx->pitch,
                                                               d_mean, ctx->d_part);
        reduce_partials_kernel<<<(len + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_part, cblocks, len, (double)count, 1, 0, d_cov);
        ctx->launches += 2;
        CU(cudaMemcpyAsync(cov_out, d_cov, len * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(ctx, SCORUS_ERR_CUDA, "stats launch: %s", cudaGetErrorString(e));
    CU(cudaStreamSynchronize(ctx->stream));
    return SCORUS_OK;
}

extern "C" int scorus_mean(scorus_ctx *ctx, size_t first, size_t count, double *mean_out)
{
    if (!ctx || !mean_out) return SCORUS_ERR_INVALID_ARG;
    return stats_impl(ctx, first, count, mean_out, nullptr);
}

extern "C" int scorus_cov(scorus_ctx *ctx, size_t first, size_t count, double *cov_out)
{
    if (!ctx || !cov_out) return SCORUS_ERR_INVALID_ARG;
    return stats_impl(ctx, first, count, nullptr, cov_out);
}

extern "C" int scorus_trace_enable(scorus_ctx *ctx, const uint32_t *walkers, const uint32_t *coords, size_t ncapture,
                                   size_t capacity_steps)
{
    if (!ctx) return SCORUS_ERR_INVALID_ARG;
    cudaSetDevice(ctx->device);
    CU(cudaStreamSynchronize(ctx->stream));
    void *old[] = {ctx->d_tr_walker, ctx->d_tr_coord, ctx->d_trace};
    for (void *b : old)
        if (b) cudaFree(b);
    ctx->d_tr_walker = ctx->d_tr_coord = nullptr;
    ctx->d_trace = nullptr;
    ctx->tr_ncap = ctx->tr_capacity = ctx->tr_count = 0;
    ctx->tr_tick = 0;
    if (ncapture == 0 || capacity_steps == 0) return SCORUS_OK;  // disabled
    if (!walkers || !coords) return SCORUS_ERR_INVALID_ARG;
    for (size_t c = 0; c < ncapture; ++c)
        if (walkers[c] >= ctx->n || coords[c] >= ctx->dim) return fail(ctx, SCORUS_ERR_INVALID_ARG, "trace entry %zu out of range", c);
    CU(cudaMalloc(&ctx->d_tr_walker, ncapture * sizeof(uint32_t)));
    CU(cudaMalloc(&ctx->d_tr_coord, ncapture * sizeof(uint32_t)));
    CU(cudaMalloc(&ctx->d_trace, ncapture * capacity_steps * sizeof(double)));
    CU(cudaMemcpy(ctx->d_tr_walker, walkers, ncapture * sizeof(uint32_t), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(ctx->d_tr_coord, coords, ncapture * sizeof(uint32_t), cudaMemcpyHostToDevice));
    ctx->tr_ncap = ncapture;
    ctx->tr_capacity = capacity_steps;
    return SCORUS_OK;
}

extern "C" int scorus_trace_download(scorus_ctx *ctx, double *out, size_t max_steps, size_t *nrecorded)
{
    if (!ctx || !nrecorded) return SCORUS_ERR_INVALID_ARG;
    cudaSetDevice(ctx->device);
    size_t k = ctx->tr_count < max_steps ? ctx->tr_count : max_steps;
    if (k && out) CU(cudaMemcpyAsync(out, ctx->d_trace, k * ctx->tr_ncap * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    *nrecorded = ctx->tr_count;
    ctx->tr_count = 0;
    return SCORUS_OK;
}

extern "C" int scorus_trace_stride(scorus_ctx *ctx, uint64_t every)
{
    if (!ctx || every == 0) return SCORUS_ERR_INVALID_ARG;
    ctx->tr_every = every;
    ctx->tr_tick = 0;
    return SCORUS_OK;
}

// ------------------------------------------------------------------------------------------
// chain statistics on the device
// ------------------------------------------------------------------------------------------
extern "C" int scorus_get_rung_stats(scorus_ctx *ctx, size_t nrungs, uint64_t *accepted, uint64_t *proposed,
                                     uint64_t *swapped, uint64_t *swap_proposed)
{
    if (!ctx || nrungs == 0) return SCORUS_ERR_INVALID_ARG;
    cudaSetDevice(ctx->device);
    std::vector<unsigned long long> h(2 * ctx->rung_cap + 4, 0);
    CU(cudaMemcpyAsync(h.data(), ctx->d_stats, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    if (ctx->rung_cap)
        CU(cudaMemcpyAsync(h.data() + 4, ctx->d_rung_stats, 2 * ctx->rung_cap * sizeof(unsigned long long),
                           cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    for (size_t r = 0; r < nrungs; ++r) {
        if (accepted) accepted[r] = r < ctx->rung_cap ? h[4 + r] : 0;
        if (proposed) proposed[r] = r < ctx->rung_proposed.size() ? ctx->rung_proposed[r] : 0;
        if (r + 1 < nrungs) {
            if (swapped) swapped[r] = r < ctx->rung_cap ? h[4 + ctx->rung_cap + r] : 0;
            if (swap_proposed) swap_proposed[r] = r < ctx->stage_proposed.size() ? ctx->stage_proposed[r] : 0;
        }
    }
    // single-rung calls are counted in the totals only (rung 0 of a one-rung ensemble)
    if (accepted) {
        unsigned long long in_rungs = 0;
        for (size_t r = 0; r < ctx->rung_cap; ++r) in_rungs += h[4 + r];
        accepted[0] += h[0] - in_rungs;
    }
    return SCORUS_OK;
}

static uint32_t moment_blocks(const scorus_ctx *ctx, uint64_t npb, uint32_t rungs, uint32_t rows_per_block)
{
    uint64_t b = ((uint64_t)ctx->num_sms * 4u + rungs - 1) / rungs;
    const uint64_t need = (npb + rows_per_block - 1) / rows_per_block;
    if (b > need) b = need;
    return (uint32_t)(b ? b : 1);
}

extern "C" int scorus_moments_enable(scorus_ctx *ctx, size_t nrungs, uint64_t every, int with_cov)
{
    if (!ctx) return SCORUS_ERR_INVALID_ARG;
    cudaSetDevice(ctx->device);
    CU(cudaStreamSynchronize(ctx->stream));
    if (ctx->d_mom) { cudaFree(ctx->d_mom); ctx->d_mom = nullptr; }
    ctx->mom_rungs = 0; ctx->mom_count = 0; ctx->mom_tick = 0;
    if (nrungs == 0) return SCORUS_OK;  // disabled
    if (every == 0) return fail(ctx, SCORUS_ERR_INVALID_ARG, "every must be >= 1");
    if (ctx->n % nrungs) return fail(ctx, SCORUS_ERR_NWALKERS_MISMATCHES_NBETA, "NWalkersMismatchesNBeta");
    const size_t dim = ctx->dim;
    if (with_cov && dim > (size_t)kCovMaxDim) return fail(ctx, SCORUS_ERR_DIM_TOO_LARGE, "cov supports dim <= %d", kCovMaxDim);
    const size_t words = nrungs * (2 * dim + (with_cov ? dim * dim : 0));
    CU(cudaMalloc(&ctx->d_mom, words * sizeof(double)));
    CU(cudaMemsetAsync(ctx->d_mom, 0, words * sizeof(double), ctx->stream));
    ctx->mom_rungs = (uint32_t)nrungs;
    ctx->mom_every = every;
    ctx->mom_cov = with_cov != 0;
    return SCORUS_OK;
}

// S1 += sum (x - shift), S2 += sum (x - shift)(x - shift)^T per rung; the shift is the rung mean at the
// first accumulation, so that the final cov = S2/M - (S1/M)(S1/M)^T does not cancel
static int accumulate_moments(scorus_ctx *ctx)
{
    const uint32_t R = ctx->mom_rungs, dim = ctx->dim, len = dim * dim;
    const uint64_t npb = ctx->n / R;
    double *shift = ctx->d_mom, *s1 = shift + (size_t)R * dim, *s2 = s1 + (size_t)R * dim;
    const uint32_t bx = moment_blocks(ctx, npb, R, 1), cbx = moment_blocks(ctx, npb, R, 32);
    int rc = ensure(ctx, (void **)&ctx->d_part, &ctx->part_cap,
                    (size_t)R * std::max((size_t)bx * dim, ctx->mom_cov ? (size_t)cbx * len : 0) * sizeof(double));
    if (rc) return rc;
    if (ctx->mom_count == 0) {
        colsum_partial_kernel<<<dim3(bx, R), 256, 0, ctx->stream>>>(ctx->d_ens, 0u, (uint32_t)npb, dim, ctx->pitch, nullptr, ctx->d_part);
        reduce_partials_kernel<<<dim3((dim + 255) / 256, R), 256, 0, ctx->stream>>>(ctx->d_part, bx, dim, 1.0 / (double)npb, 0, 0, shift);
        ctx->launches += 2;
    }
    colsum_partial_kernel<<<dim3(bx, R), 256, 0, ctx->stream>>>(ctx->d_ens, 0u, (uint32_t)npb, dim, ctx->pitch, shift, ctx->d_part);
    reduce_partials_kernel<<<dim3((dim + 255) / 256, R), 256, 0, ctx->stream>>>(ctx->d_part, bx, dim, 1.0, 0, 1, s1);
    ctx->launches += 2;
    if (ctx->mom_cov) {
        const size_t smem = 32u * (size_t)dim * sizeof(double);
        cudaError_t e = cudaFuncSetAttribute(cov_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail(ctx, SCORUS_ERR_CUDA, "cov smem: %s", cudaGetErrorString(e));
        cov_partial_kernel<<<dim3(cbx, R), 256, smem, ctx->stream>>>(ctx->d_ens, 0u, (uint32_t)npb, dim, ctx->pitch, shift, ctx->d_part);
        reduce_partials_kernel<<<dim3((len + 255) / 256, R), 256, 0, ctx->stream>>>(ctx->d_part, cbx, len, 1.0, 0, 1, s2);
        ctx->launches += 2;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(ctx, SCORUS_ERR_CUDA, "moments launch: %s", cudaGetErrorString(e));
    ++ctx->mom_count;
    return SCORUS_OK;
}

extern "C" int scorus_moments_download(scorus_ctx *ctx, size_t rung, double *mean_out, double *cov_out, uint64_t *nsamples)
{
    if (!ctx) return SCORUS_ERR_INVALID_ARG;
    if (!ctx->mom_rungs) return fail(ctx, SCORUS_ERR_INVALID_ARG, "call scorus_moments_enable first");
    if (rung >= ctx->mom_rungs) return fail(ctx, SCORUS_ERR_INVALID_ARG, "rung %zu of %u", rung, ctx->mom_rungs);
    if (cov_out && !ctx->mom_cov) return fail(ctx, SCORUS_ERR_INVALID_ARG, "moments were enabled without the covariance");
    cudaSetDevice(ctx->device);
    const size_t R = ctx->mom_rungs, dim = ctx->dim, len = dim * dim;
    const uint64_t m = ctx->mom_count * (ctx->n / R);
    if (nsamples) *nsamples = m;
    if (m == 0) return fail(ctx, SCORUS_ERR_INVALID_ARG, "no step has been accumulated yet");
    std::vector<double> shift(dim), s1(dim), s2(cov_out ? len : 0);
    const double *d_shift = ctx->d_mom, *d_s1 = d_shift + R * dim, *d_s2 = d_s1 + R * dim;
    CU(cudaMemcpyAsync(shift.data(), d_shift + rung * dim, dim * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(s1.data(), d_s1 + rung * dim, dim * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (cov_out) CU(cudaMemcpyAsync(s2.data(), d_s2 + rung * len, len * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    for (size_t d = 0; d < dim; ++d) s1[d] /= (double)m;
    if (mean_out)
        for (size_t d = 0; d < dim; ++d) mean_out[d] = shift[d] + s1[d];
    if (cov_out)
        for (size_t i = 0; i < dim; ++i)
            for (size_t j = 0; j < dim; ++j) cov_out[i * dim + j] = s2[i * dim + j] / (double)m - s1[i] * s1[j];
    return SCORUS_OK;
}

// integrated autocorrelation time of every captured series, from the rows recorded so far (not rewound)
extern "C" int scorus_trace_iat(scorus_ctx *ctx, double c_window, size_t max_lag, double *tau_out, uint32_t *window_out,
                                double *mean_out, double *var_out)
{
    if (!ctx || !tau_out) return SCORUS_ERR_INVALID_ARG;
    if (!ctx->tr_ncap) return fail(ctx, SCORUS_ERR_INVALID_ARG, "call scorus_trace_enable first");
    const size_t T = ctx->tr_count, C = ctx->tr_ncap;
    if (T < 2) return fail(ctx, SCORUS_ERR_INVALID_ARG, "need at least two recorded steps");
    if (!(c_window > 0.0)) return fail(ctx, SCORUS_ERR_INVALID_ARG, "window factor must be > 0");
    const size_t L = max_lag == 0 || max_lag > T - 1 ? T - 1 : max_lag;
    cudaSetDevice(ctx->device);
    // [xc: C*T][acov: C*(L+1)][mean: C][tau: C][window: C (u32)]
    const size_t words = C * T + C * (L + 1) + 3 * C;
    int rc = ensure(ctx, (void **)&ctx->d_iat, &ctx->iat_cap, words * sizeof(double));
    if (rc) return rc;
    double *xc = ctx->d_iat, *acov = xc + C * T, *mean = acov + C * (L + 1), *tau = mean + C;
    uint32_t *window = reinterpret_cast<uint32_t *>(tau + C);
    iat_center_kernel<<<(unsigned)C, 256, 0, ctx->stream>>>(ctx->d_trace, (uint32_t)C, (uint32_t)T, xc, mean);
    iat_autocov_kernel<<<dim3((unsigned)((L + 128) / 128), (unsigned)C), 128, 0, ctx->stream>>>(xc, (uint32_t)T, (uint32_t)L, acov);
    iat_window_kernel<<<(unsigned)((C + 63) / 64), 64, 0, ctx->stream>>>(acov, (uint32_t)C, (uint32_t)L, c_window, tau, window);
    ctx->launches += 3;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(ctx, SCORUS_ERR_CUDA, "iat launch: %s", cudaGetErrorString(e));
    CU(cudaMemcpyAsync(tau_out, tau, C * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (window_out) CU(cudaMemcpyAsync(window_out, window, C * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (mean_out) CU(cudaMemcpyAsync(mean_out, mean, C * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (var_out) CU(cudaMemcpy2DAsync(var_out, sizeof(double), acov, (L + 1) * sizeof(double), sizeof(double), C,
                                      cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return SCORUS_OK;
}

// ------------------------------------------------------------------------------------------
// t-walk: src/mcmc/twalk.rs
// ------------------------------------------------------------------------------------------
extern "C" int scorus_twalk_params_default(size_t dim, scorus_twalk_params *out)
{
    if (!out || dim == 0) return SCORUS_ERR_INVALID_ARG;
    const double n = (double)dim, n1phi = 4.0;  // TWalkParams::new, twalk.rs:89-110
    out->aw = 1.5;
    out->at = 6.0;
    out->pphi = (n < n1phi ? n : n1phi) / n;
    out->fw[0] = 0.0 + 0.5;
    out->fw[1] = out->fw[0] + 0.5;
    return SCORUS_OK;
}

static int twalk_common(scorus_ctx *ctx, const scorus_twalk_params *tw, const double *beta, size_t nbeta, TwalkParams *tp,
                        Geometry *g)
{
    int rc = scorus_check_sample_args(ctx->n, ctx->n, nbeta);  // n = B * n/B (:675), pairs need an even rung (:610)
    if (rc) return fail(ctx, rc, "%s", scorus_status_string(rc));
    if (ctx->lp_kind < 0) return fail(ctx, SCORUS_ERR_NO_LOGPROB, "call scorus_set_logprob first");
    if (!ctx->uploaded) return fail(ctx, SCORUS_ERR_NOT_UPLOADED, "call scorus_upload first");
    if (!(tw->pphi > 0.0)) return fail(ctx, SCORUS_ERR_INVALID_ARG, "pphi must be > 0 (gen_update_flags would loop forever)");
    if (!(tw->fw[1] > 0.0) || !(tw->fw[0] >= 0.0) || tw->fw[0] > tw->fw[1])
        return fail(ctx, SCORUS_ERR_INVALID_ARG, "fw must be cumulative weights, 0 <= fw[0] <= fw[1], fw[1] > 0");
    cudaSetDevice(ctx->device);
    if ((rc = upload_beta(ctx, beta, nbeta))) return rc;
    if ((rc = fill_common(ctx, &tp->s, g, nbeta))) return rc;
    if (g->kernel != KERNEL_REG)
        return fail(ctx, SCORUS_ERR_INVALID_ARG, "the t-walk needs the register kernel (dim <= 256, no SEQUENTIAL_SUM)");
    tp->aw = tw->aw;
    tp->aw_ratio = tw->aw / (1.0 + tw->aw);  // twalk.rs:292
    tp->at = tw->at;
    tp->fw0 = tw->fw[0];
    tp->fw1 = tw->fw[1];
    tp->kernel_in = nullptr; tp->b_in = nullptr; tp->walk_u_in = nullptr;
    tp->l2_keep = (uint32_t)env_int("SCORUS_TWALK_L2", 1);
    StepParams &p = tp->s;
    p.flag_kind = SCORUS_UFS_PROB;
    p.flag_bits = flag_bits_for(tw->pphi);
    p.flag_thr = p.flag_bits == 32 ? (uint64_t)llrint(fmin(tw->pphi, 1.0) * 4294967296.0)
                                   : (uint64_t)(fmin(tw->pphi, 1.0) * (double)(1u << p.flag_bits));
    return SCORUS_OK;
}

// twalk::sample (:586-648): one matching, two half-passes (sample1 :650-762) with the roles exchanged
extern "C" int scorus_twalk_sample(scorus_ctx *ctx, const scorus_twalk_params *tw, const double *beta, size_t nbeta,
                                   uint64_t nsteps)
{
    if (!ctx || !tw || !beta) return SCORUS_ERR_INVALID_ARG;
    TwalkParams tp;
    Geometry g;
    memset(&tp, 0, sizeof(tp));
    int rc = twalk_common(ctx, tw, beta, nbeta, &tp, &g);
    if (rc) return rc;
    StepParams &p = tp.s;
    const bool exact = p.npb <= kExactShuffleMax;
    const bool fused = env_int("SCORUS_TWALK_FUSED", 1) != 0;
    if (exact && !ctx->d_order) CU(cudaMalloc(&ctx->d_order, ctx->n * sizeof(uint32_t)));
    for (uint64_t s = 0; s < nsteps; ++s) {
        tp.order_step = ctx->step;
        if (exact) {
            exact_order_kernel<<<(unsigned)nbeta, (p.npb + 31u) & ~31u, 0, ctx->stream>>>(ctx->seed, ctx->step, p.npb, ctx->d_order, (uint32_t)ctx->w_off);
            ++ctx->launches;
            p.order = ctx->d_order;
        }
        if (fused) {  // both half-passes of every pair in one launch (twalk_kernel.cuh); counters step, step + 1
            tp.half = 2;
            p.step = ctx->step;
            stamp_tiles(ctx, &p, g);
            cudaError_t e = launch_twalk(ctx->lp_kind, tp, g, ctx->stream);
            if (e != cudaSuccess) return fail(ctx, SCORUS_ERR_CUDA, "t-walk launch: %s", cudaGetErrorString(e));
            ++ctx->launches;
            ctx->step += 2;
        } else {
            for (uint32_t half = 0; half < 2; ++half) {
                tp.half = half;
                p.step = ctx->step;
                stamp_tiles(ctx, &p, g);
                cudaError_t e = launch_twalk(ctx->lp_kind, tp, g, ctx->stream);
                if (e != cudaSuccess) return fail(ctx, SCORUS_ERR_CUDA, "t-walk launch: %s", cudaGetErrorString(e));
                ++ctx->launches;
                ++ctx->step;
            }
        }
        if ((rc = after_step(ctx))) return rc;
    }
    count_proposed(ctx, nbeta, p.npb, 0, ctx->n, nsteps);  // every walker is proposed once per t-walk step
    return SCORUS_OK;
}

extern "C" int scorus_twalk_half_fixed(scorus_ctx *ctx, const scorus_twalk_params *tw, const double *beta, size_t nbeta,
                                       const uint32_t *order, int half, const uint8_t *kernel, const double *b,
                                       const uint8_t *flags, const double *walk_u, const double *u, uint8_t *accepted_out)
{
    if (!ctx || !tw || !beta || !order || !kernel || !b || !flags || !walk_u || !u) return SCORUS_ERR_INVALID_ARG;
    if (half != 0 && half != 1) return fail(ctx, SCORUS_ERR_INVALID_ARG, "half must be 0 or 1");
    TwalkParams tp;
    Geometry g;
    memset(&tp, 0, sizeof(tp));
    int rc = twalk_common(ctx, tw, beta, nbeta, &tp, &g);
    if (rc) return rc;
    StepParams &p = tp.s;
    const size_t n = ctx->n, dim = ctx->dim, npb = p.npb, npair = n / 2;
    // the matching: a permutation whose pairs (2k, 2k+1) stay inside one rung (:605-612)
    std::vector<uint8_t> seen(n, 0);
    for (size_t k = 0; k < n; ++k) {
        if (order[k] >= n || seen[order[k]] || order[k] / npb != order[k ^ 1] / npb || order[k] / npb != k / npb)
            return fail(ctx, SCORUS_ERR_INVALID_ARG, "order[] is not a rung-by-rung permutation at position %zu", k);
        seen[order[k]] = 1;
    }
    // per-pair draws -> per-walker arrays (the kernel indexes its streams by walker)
    std::vector<uint8_t> kw(n, 0), fw(n * dim, 0);
    std::vector<double> bw(n, 1.0), uw(n, 1.0), ww(n * dim, 0.0);
    for (size_t k = 0; k < npair; ++k) {
        if (kernel[k] > 1) return fail(ctx, SCORUS_ERR_INVALID_ARG, "kernel[%zu] must be 0 (Walk) or 1 (Traverse)", k);
        const size_t w = order[2 * k + half];
        kw[w] = kernel[k]; bw[w] = b[k]; uw[w] = u[k];
        memcpy(&fw[w * dim], flags + k * dim, dim);
        memcpy(&ww[w * dim], walk_u + k * dim, dim * sizeof(double));
    }
    if (!ctx->d_order) CU(cudaMalloc(&ctx->d_order, n * sizeof(uint32_t)));
    if (!ctx->d_z) CU(cudaMalloc(&ctx->d_z, n * sizeof(double)));
    if (!ctx->d_u) CU(cudaMalloc(&ctx->d_u, n * sizeof(double)));
    if (!ctx->d_acc) CU(cudaMalloc(&ctx->d_acc, n));
    if ((rc = ensure(ctx, (void **)&ctx->d_flags, &ctx->flags_cap, n * dim))) return rc;
    if ((rc = ensure(ctx, (void **)&ctx->d_tw_kernel, &ctx->tw_kernel_cap, n))) return rc;
    if ((rc = ensure(ctx, (void **)&ctx->d_tw_walk, &ctx->tw_walk_cap, n * dim * sizeof(double)))) return rc;
    CU(cudaMemcpyAsync(ctx->d_order, order, n * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_z, bw.data(), n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_u, uw.data(), n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_flags, fw.data(), n * dim, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_tw_kernel, kw.data(), n, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_tw_walk, ww.data(), n * dim * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemsetAsync(ctx->d_acc, 0, n, ctx->stream));
    p.order = ctx->d_order; p.u_in = ctx->d_u; p.flags_in = ctx->d_flags; p.flag_len = (uint32_t)dim;
    p.flag_kind = SCORUS_UFS_MASK; p.accepted_out = ctx->d_acc;
    tp.kernel_in = ctx->d_tw_kernel; tp.b_in = ctx->d_z; tp.walk_u_in = ctx->d_tw_walk;
    tp.half = (uint32_t)half;
    tp.order_step = ctx->step;
    p.step = ctx->step;
    stamp_tiles(ctx, &p, g);
    cudaError_t e = launch_twalk(ctx->lp_kind, tp, g, ctx->stream);
    if (e != cudaSuccess) return fail(ctx, SCORUS_ERR_CUDA, "t-walk launch: %s", cudaGetErrorString(e));
    ++ctx->launches;
    // half a step: the movers of this half-pass were proposed
    ctx->proposed += npair;
    if (ctx->rung_proposed.size() < nbeta) ctx->rung_proposed.resize(nbeta, 0);
    for (size_t r = 0; r < nbeta; ++r) ctx->rung_proposed[r] += npb / 2;
    std::vector<uint8_t> acc_w(accepted_out ? n : 0);
    if (accepted_out) CU(cudaMemcpyAsync(acc_w.data(), ctx->d_acc, n, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));  // the staging vectors are released here
    if (accepted_out)
        for (size_t k = 0; k < npair; ++k) accepted_out[k] = acc_w[order[2 * k + half]];
    return SCORUS_OK;
}

// the device's sim_b value (:306-322) of walkers [0, count) at step counter `step`: the one draw of the t-walk
// that goes through pow(), exposed so that a device-stream run can be replayed exactly
__global__ void twalk_b_kernel(uint64_t seed, uint64_t step, uint32_t w_off, double at, uint32_t count, double *out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const U4 q = draw(seed, step, ST_TW_B, i + w_off, 0);
    out[i] = twalk_b(u53(q.x, q.y), u53(q.z, q.w), at);
}

extern "C" int scorus_twalk_draw_b(scorus_ctx *ctx, double at, uint64_t step, size_t count, double *out)
{
    if (!ctx || !out || count > ctx->n) return SCORUS_ERR_INVALID_ARG;
    cudaSetDevice(ctx->device);
    if (count == 0) return SCORUS_OK;
    if (!ctx->d_z) CU(cudaMalloc(&ctx->d_z, ctx->n * sizeof(double)));
    twalk_b_kernel<<<(unsigned)((count + 127) / 128), 128, 0, ctx->stream>>>(ctx->seed, step, (uint32_t)ctx->w_off, at,
                                                                           (uint32_t)count, ctx->d_z);
    ++ctx->launches;
    CU(cudaMemcpyAsync(out, ctx->d_z, count * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return SCORUS_OK;
}

// ------------------------------------------------------------------------------------------
// ensemble / parallel-tempering HMC: src/mcmc/hmc/naive.rs:122-292
// ------------------------------------------------------------------------------------------
static int hmc_impl(scorus_ctx *ctx, double *epsilon, const double *beta, size_t nbeta, size_t l, const scorus_hmc_param *param,
                    const double *p0, const double *u1, const double *u2, uint8_t *accepted_out, uint64_t *accepted_cnt,
                    uint64_t nsteps)
{
    if (nbeta == 0) return fail(ctx, SCORUS_ERR_INVALID_ARG, "nbeta == 0");
    const size_t n = ctx->n, npb = n / nbeta, dim = ctx->dim;
    // beta_list[i / n_per_beta] (:222) indexes past the list when the rungs do not divide the ensemble
    if (npb * nbeta != n) return fail(ctx, SCORUS_ERR_NWALKERS_MISMATCHES_NBETA, "NWalkersMismatchesNBeta");
    if (ctx->lp_kind < 0) return fail(ctx, SCORUS_ERR_NO_LOGPROB, "call scorus_set_logprob first");
    if (!ctx->uploaded) return fail(ctx, SCORUS_ERR_NOT_UPLOADED, "call scorus_upload first");
    if (ctx->pitch > 256) return fail(ctx, SCORUS_ERR_DIM_TOO_LARGE, "HMC keeps a walker's q, p and force in registers: dim <= 256");
    cudaSetDevice(ctx->device);
    int rc = upload_beta(ctx, beta, nbeta);
    if (rc) return rc;
    HmcParams hp;
    memset(&hp, 0, sizeof(hp));
    Geometry g;
    if ((rc = fill_common(ctx, &hp.s, &g, nbeta))) return rc;
    if (!ctx->d_hmc_adapt) CU(cudaMalloc(&ctx->d_hmc_adapt, n));
    if ((rc = ensure(ctx, (void **)&ctx->d_hmc_eps, &ctx->hmc_eps_cap, nbeta * sizeof(double)))) return rc;
    CU(cudaMemcpyAsync(ctx->d_hmc_eps, epsilon, nbeta * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    hp.eps = ctx->d_hmc_eps;
    hp.adapt_out = ctx->d_hmc_adapt;
    hp.target_accept_ratio = param->target_accept_ratio;
    hp.l = (uint32_t)l;
    StepParams &p = hp.s;
    if (p0) {  // caller-supplied draws (one step)
        if (!ctx->d_u) CU(cudaMalloc(&ctx->d_u, n * sizeof(double)));
        if (!ctx->d_z) CU(cudaMalloc(&ctx->d_z, n * sizeof(double)));
        if (!ctx->d_acc) CU(cudaMalloc(&ctx->d_acc, n));
        if ((rc = ensure(ctx, (void **)&ctx->d_hmc_p, &ctx->hmc_p_cap, n * dim * sizeof(double)))) return rc;
        CU(cudaMemcpyAsync(ctx->d_hmc_p, p0, n * dim * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(ctx->d_u, u1, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(ctx->d_z, u2, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        hp.p_in = ctx->d_hmc_p; p.u_in = ctx->d_u; hp.u2_in = ctx->d_z; p.accepted_out = ctx->d_acc;
    }
    // accepted_cnt: the per-rung acceptance counters before and after (single rung: the total)
    std::vector<unsigned long long> before(nbeta, 0), after(nbeta, 0);
    unsigned long long *cnt_src = nbeta > 1 ? ctx->d_rung_stats : ctx->d_stats;
    if (accepted_cnt) {
        CU(cudaMemcpyAsync(before.data(), cnt_src, nbeta * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    for (uint64_t s = 0; s < nsteps; ++s) {
        p.step = ctx->step;
        cudaError_t e = launch_hmc(ctx->lp_kind, hp, (uint32_t)ctx->num_sms, ctx->stream);
        if (e == cudaErrorNotSupported) return fail(ctx, SCORUS_ERR_INVALID_ARG, "this log-density plugin has no gradient");
        if (e != cudaSuccess) return fail(ctx, SCORUS_ERR_CUDA, "hmc launch: %s", cudaGetErrorString(e));
        ++ctx->launches;
        if (param->adj_factor != 0.0) {  // HmcParam::fixed: epsilon * 1 and epsilon / 1 change nothing
            if (npb <= kHmcExactAdapt)
                hmc_adapt_kernel<<<(unsigned)((nbeta + 63) / 64), 64, 0, ctx->stream>>>(ctx->d_hmc_adapt, (uint32_t)npb, (uint32_t)nbeta,
                                                                                      param->adj_factor, ctx->d_hmc_eps);
            else
                hmc_adapt_count_kernel<<<(unsigned)nbeta, 256, 0, ctx->stream>>>(ctx->d_hmc_adapt, (uint32_t)npb, param->adj_factor,
                                                                               ctx->d_hmc_eps);
            ++ctx->launches;
        }
        ++ctx->step;
        if ((rc = after_step(ctx))) return rc;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(ctx, SCORUS_ERR_CUDA, "hmc launch: %s", cudaGetErrorString(e));
    count_proposed(ctx, nbeta, npb, 0, n, nsteps);
    CU(cudaMemcpyAsync(epsilon, ctx->d_hmc_eps, nbeta * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (accepted_out) CU(cudaMemcpyAsync(accepted_out, ctx->d_acc, n, cudaMemcpyDeviceToHost, ctx->stream));
    if (accepted_cnt) CU(cudaMemcpyAsync(after.data(), cnt_src, nbeta * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));  // epsilon is an in/out host array: the call is synchronous
    if (accepted_cnt)
        for (size_t r = 0; r < nbeta; ++r) accepted_cnt[r] = after[r] - before[r];
    return SCORUS_OK;
}

extern "C" int scorus_hmc_sample_ensemble_pt(scorus_ctx *ctx, double *epsilon, const double *beta, size_t nbeta, size_t l,
                                             const scorus_hmc_param *param, uint64_t *accepted_cnt, uint64_t nsteps)
{
    if (!ctx || !epsilon || !beta || !param) return SCORUS_ERR_INVALID_ARG;
    return hmc_impl(ctx, epsilon, beta, nbeta, l, param, nullptr, nullptr, nullptr, nullptr, accepted_cnt, nsteps);
}

extern "C" int scorus_hmc_sample_ensemble_pt_fixed(scorus_ctx *ctx, double *epsilon, const double *beta, size_t nbeta, size_t l,
                                                   const scorus_hmc_param *param, const double *p0, const double *u1,
                                                   const double *u2, uint8_t *accepted_out, uint64_t *accepted_cnt)
{
    if (!ctx || !epsilon || !beta || !param || !p0 || !u1 || !u2) return SCORUS_ERR_INVALID_ARG;
    return hmc_impl(ctx, epsilon, beta, nbeta, l, param, p0, u1, u2, accepted_out, accepted_cnt, 1);
}

// sample_ensemble_pt on caller-owned host buffers (q0: &mut [V], lp: &mut [T], :125-126)
extern "C" int scorus_hmc_sample_ensemble_pt_host(scorus_ctx *ctx, double *q0, double *lp, size_t lp_len, double *epsilon,
                                                  const double *beta, size_t nbeta, size_t l, const scorus_hmc_param *param,
                                                  uint64_t *accepted_cnt)
{
    if (!ctx || !q0 || !lp || !epsilon || !beta || !param) return SCORUS_ERR_INVALID_ARG;
    if (lp_len != ctx->n) return fail(ctx, SCORUS_ERR_LENGTH_MISMATCH, "assert_eq!(q0.len(), lp.len())");  // :177
    int rc = scorus_upload(ctx, q0, lp);
    if (rc) return rc;
    if ((rc = hmc_impl(ctx, epsilon, beta, nbeta, l, param, nullptr, nullptr, nullptr, nullptr, accepted_cnt, 1))) return rc;
    return scorus_download(ctx, q0, lp);
}

// the device stream's momenta of walkers [0, count) at step counter `step` (they go through log / sqrt / sincos;
// exposed so that a device-stream run can be replayed exactly through scorus_hmc_sample_ensemble_pt_fixed)
__global__ void hmc_momenta_kernel(uint64_t seed, uint64_t step, uint32_t w_off, uint32_t count, uint32_t dim, double *out)
{
    const uint32_t half = (dim + 1u) / 2u;
    const size_t total = (size_t)count * half;
    for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (size_t)gridDim.x * blockDim.x) {
        const uint32_t w = (uint32_t)(g / half), c = (uint32_t)(g - (size_t)w * half);
        const double2 m = hmc_momentum(seed, step, w + w_off, c);
        out[(size_t)w * dim + 2u * c] = m.x;
        if (2u * c + 1u < dim) out[(size_t)w * dim + 2u * c + 1u] = m.y;
    }
}

extern "C" int scorus_hmc_draw_momenta(scorus_ctx *ctx, uint64_t step, size_t count, double *out)
{
    if (!ctx || !out || count > ctx->n) return SCORUS_ERR_INVALID_ARG;
    cudaSetDevice(ctx->device);
    if (count == 0) return SCORUS_OK;
    int rc = ensure(ctx, (void **)&ctx->d_hmc_p, &ctx->hmc_p_cap, ctx->n * (size_t)ctx->dim * sizeof(double));
    if (rc) return rc;
    hmc_momenta_kernel<<<(unsigned)ctx->num_sms * 4u, 256, 0, ctx->stream>>>(ctx->seed, step, (uint32_t)ctx->w_off, (uint32_t)count,
                                                                           ctx->dim, ctx->d_hmc_p);
    ++ctx->launches;
    CU(cudaMemcpyAsync(out, ctx->d_hmc_p, count * ctx->dim * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return SCORUS_OK;
}

// ------------------------------------------------------------------------------------------
// literal drop-ins on host buffers
// ------------------------------------------------------------------------------------------
// upload -> run(ctx) -> download, shared by the host-buffer entry points
template <class Run>
static int host_call(scorus_ctx *ctx, double *ens, double *lp, size_t lp_len, size_t nbeta, Run run)
{
    int rc = scorus_check_sample_args(ctx->n, lp_len, nbeta);
    if (rc) return fail(ctx, rc, "%s", scorus_status_string(rc));
    if ((rc = scorus_upload(ctx, ens, lp))) return rc;
    // Pinned (device-mapped) host buffers: after the steps only the accepted walkers differ from what
    // was just uploaded, so only those rows travel back, written in place by scatter_download_kernel.
    // Pageable buffers take the full copy.
    cudaPointerAttributes ae{}, al{};
    const bool pinned = env_int("SCORUS_SCATTER_DOWNLOAD", 1) &&
                        cudaPointerGetAttributes(&ae, ens) == cudaSuccess && ae.type == cudaMemoryTypeHost &&
                        cudaPointerGetAttributes(&al, lp) == cudaSuccess && al.type == cudaMemoryTypeHost;
    cudaGetLastError();
    if (!pinned) {
        if ((rc = run())) return rc;
        return scorus_download(ctx, ens, lp);
    }
    if (!ctx->d_dirty) CU(cudaMalloc(&ctx->d_dirty, ctx->n));
    CU(cudaMemsetAsync(ctx->d_dirty, 0, ctx->n, ctx->stream));
    ctx->track_dirty = true;
    rc = run();
    ctx->track_dirty = false;
    if (rc) return rc;
    scatter_download_kernel<<<(unsigned)ctx->num_sms * 8u, 256, 0, ctx->stream>>>(
        ctx->d_ens, ctx->d_lp, ctx->d_dirty, (uint32_t)ctx->n, ctx->dim, ctx->pitch, (double *)ae.devicePointer,
        (double *)al.devicePointer);
    ++ctx->launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(ctx, SCORUS_ERR_CUDA, "scatter download launch: %s", cudaGetErrorString(e));
    CU(cudaStreamSynchronize(ctx->stream));
    return SCORUS_OK;
}

extern "C" int scorus_sample_pt_host_n(scorus_ctx *ctx, double *ens, double *lp, size_t lp_len, double a,
                                       const scorus_ufs *ufs, const double *beta, size_t nbeta, uint64_t nsteps)
{
    if (!ctx || !ens || !lp) return SCORUS_ERR_INVALID_ARG;
    return host_call(ctx, ens, lp, lp_len, nbeta, [&] { return scorus_sample_pt(ctx, a, ufs, beta, nbeta, nsteps); });
}

// twalk::sample on host buffers (the `ensemble_logprob: &mut (W, X)` argument, twalk.rs:588)
extern "C" int scorus_twalk_sample_host(scorus_ctx *ctx, double *ens, double *lp, size_t lp_len,
                                        const scorus_twalk_params *tw, const double *beta, size_t nbeta, uint64_t nsteps)
{
    if (!ctx || !ens || !lp || !tw) return SCORUS_ERR_INVALID_ARG;
    return host_call(ctx, ens, lp, lp_len, nbeta, [&] { return scorus_twalk_sample(ctx, tw, beta, nbeta, nsteps); });
}

extern "C" int scorus_sample_pt_host(scorus_ctx *ctx, double *ens, double *lp, size_t lp_len, double a,
                                     const scorus_ufs *ufs, const double *beta, size_t nbeta)
{
    return scorus_sample_pt_host_n(ctx, ens, lp, lp_len, a, ufs, beta, nbeta, 1);
}

extern "C" int scorus_swap_walkers_host(scorus_ctx *ctx, double *ens, double *lp, size_t lp_len,
                                        const double *beta, size_t nbeta)
{
    if (!ctx || !ens || !lp) return SCORUS_ERR_INVALID_ARG;
    if (nbeta == 0) return SCORUS_ERR_INVALID_ARG;
    if ((ctx->n / nbeta) * nbeta != ctx->n) return fail(ctx, SCORUS_ERR_NWALKERS_MISMATCHES_NBETA, "NWalkersMismatchesNBeta");
    if (lp_len != ctx->n) return SCORUS_OK;  // src/mcmc/utils.rs:88: silently does nothing
    int rc = scorus_check_beta_order(beta, nbeta);
    if (rc) return fail(ctx, rc, "beta list must be in decreasing order");
    if ((rc = scorus_upload(ctx, ens, lp))) return rc;
    if ((rc = scorus_swap_walkers(ctx, beta, nbeta))) return rc;
    return scorus_download(ctx, ens, lp);
}
