// api_core.cu -- C ABI of libscorus_b200.so, part 1 (see include/scorus_b200.h): library queries, the
// context, plugin registration, host <-> device copies, and the hot path itself -- sample / sample_pt
// (src/mcmc/ensemble_sample.rs:87-190) and swap_walkers (src/mcmc/utils.rs:72-112).  Validates the
// reference's preconditions and launches the kernels of step_*_kernel.cuh and swap_kernel.cuh.
#include "host_common.cuh"
#include "step_small_kernel.cuh"
#include "swap_kernel.cuh"

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
    ctx->f32 = (cfg->flags & SCORUS_CFG_F32) != 0;
    if (ctx->f32 && (cfg->flags & SCORUS_CFG_SEQUENTIAL_SUM)) { delete ctx; return SCORUS_ERR_INVALID_ARG; }
    ctx->elem = ctx->f32 ? 4u : 8u;
    // rows are 16-byte multiples for cp.async.bulk: 2 doubles or 4 floats
    ctx->pitch = ctx->f32 ? ((cfg->dim + 3u) & ~3u) : cfg->dim + (cfg->dim & 1u);
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
    size_t ens_bytes = (size_t)ctx->n * ctx->pitch * ctx->elem;
    if ((e = cudaMalloc(&ctx->d_ens, ens_bytes)) != cudaSuccess) return bail(e, "cudaMalloc ensemble");
    // 32 extra slots behind the log-probs: the flag row of the inter-rank barrier (api_shard.cu), reachable by the peers
    // through the same IPC mapping as the log-probs
    // ... and one byte per walker behind those: "accepted since the upload" marks of the host-buffer calls, in the same
    // allocation so that the owner of a pair on another GPU can mark its partner there (one-sided mode)
    if ((e = cudaMalloc(&ctx->d_lp, (ctx->n + 32) * sizeof(double) + ((ctx->n + 15) & ~(uint64_t)15))) != cudaSuccess)
        return bail(e, "cudaMalloc logprob");
    cudaMemsetAsync(ctx->d_lp + ctx->n, 0, 32 * sizeof(double), ctx->stream);
    ctx->d_dirty = reinterpret_cast<uint8_t *>(ctx->d_lp + ctx->n + 32);
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
                    ctx->d_beta_f32, ctx->d_flag_cdf, ctx->d_jinv, ctx->d_r, ctx->d_swapped, ctx->d_chain, ctx->d_swap_out, ctx->d_lp_all, ctx->d_src_all, ctx->d_gorder, ctx->d_pairs, ctx->d_count, ctx->d_tile_counter};
    for (void *b : bufs)
        if (b) cudaFree(b);
    for (auto &e : ctx->beta_cache)
        if (e.dev) cudaFree(e.dev);
    for (int q = 0; q < ctx->world; ++q)
        if (q != ctx->rank) {
            if (ctx->peer_ens[q]) cudaIpcCloseMemHandle(ctx->peer_ens[q]);
            if (ctx->peer_lp[q]) cudaIpcCloseMemHandle(ctx->peer_lp[q]);
        }
    void *p2p[] = {ctx->d_peer_ens, ctx->d_peer_lp, ctx->d_part, ctx->d_stat,
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
    F64_ONLY(ctx);
    int rc = settle_swap(ctx);  // the pointers handed out are of the buffers as the caller expects them: rows in their slots
    if (rc) return rc;
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

int scorus::host::ensure(scorus_ctx *ctx, void **buf, size_t *cap, size_t bytes)
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
    if (ctx->f32) {  // parameters as float on the device (api_f32.cu)
        const int old = ctx->lp_kind;
        ctx->lp_kind = kind;
        int rc32 = f32_set_params(ctx, params, nparams);
        if (rc32) { ctx->lp_kind = old; return rc32; }
        ctx->nparams = (uint32_t)nparams;
        return SCORUS_OK;
    }
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

int scorus::host::env_int(const char *name, int dflt)
{
    const char *v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

int scorus::host::step_geometry(scorus_ctx *ctx, Geometry *g, bool tile_in_smem)
{
    g->row_bytes = ctx->pitch * ctx->elem;
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
        // per warp: the flag words, plus -- for the kernels that stage proposals in shared memory (the dense target's
        // tensor-core evaluation, the t-walk) -- one tile of 32 rows.  Those leave SCORUS_L1_KB (default 32) of the
        // SM's unified array to L1: the row loads stream through it, and with the carve-out at its maximum the
        // kernel ran 12 % slower (profiles/r1_notes.md).  The stretch move on the other targets keeps its proposals
        // in registers (step_reg_kernel.cuh, kDirect) and is bound by registers only: 16 warps.
        const bool dense = ctx->lp_kind == SCORUS_LP_CORR_GAUSSIAN;
        const uint32_t nvec = ctx->pitch / 2;
        // mirrors reg_kernel_direct (step_reg_kernel.cuh): rows of 9..16 double2 (16-lane groups) keep the tile
        const bool tiled = tile_in_smem || dense || (!SCORUS_G16_DIRECT && nvec > 8 && nvec <= 16);
        const size_t per_warp = (tiled ? tile : 0) + flag_bytes;
        const size_t l1_keep = tiled ? (size_t)env_int("SCORUS_L1_KB", 32) * 1024u : 0;
        // mirrors reg_kernel_max_warps / twalk_max_warps (step_reg_kernel.cuh): dense plugin 8, the t-walk's tile kernel
        // 16, the stretch kernels SCORUS_REG_WARPS
        uint32_t w = dense ? 8 : (tile_in_smem ? 16 : (nvec > (SCORUS_G16_WIDE ? 8u : 16u) && nvec <= 64 ? SCORUS_REG_WARPS_WIDE : SCORUS_REG_WARPS));
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
    if (ctx->f32) return f32_init_logprob(ctx);
    Geometry g;
    int rc = settle_swap(ctx);
    if (rc) return rc;
    if ((rc = step_geometry(ctx, &g, false))) return rc;
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
    F64_ONLY(ctx);
    cudaSetDevice(ctx->device);
    ctx->swap_pending = false;  // everything a pending swap would have moved is overwritten here
    const size_t rb = (size_t)ctx->dim * sizeof(double);
    // (rows without padding: one flat copy -- a 2-D copy of 80-byte rows ran at 2 GB/s, 48 ms for the 84 MB of c4)
    if (ctx->pitch == ctx->dim)
        CU(cudaMemcpyAsync(ctx->d_ens, ens, rb * ctx->n, cudaMemcpyHostToDevice, ctx->stream));
    else
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
    F64_ONLY(ctx);
    if (!ctx->uploaded) return fail(ctx, SCORUS_ERR_NOT_UPLOADED, "nothing uploaded");
    cudaSetDevice(ctx->device);
    int rc = settle_swap(ctx);
    if (rc) return rc;
    const size_t rb = (size_t)ctx->dim * sizeof(double);
    if (ens && ctx->pitch == ctx->dim)
        CU(cudaMemcpyAsync(ens, ctx->d_ens, rb * ctx->n, cudaMemcpyDeviceToHost, ctx->stream));
    else if (ens)
        CU(cudaMemcpy2DAsync(ens, rb, ctx->d_ens, (size_t)ctx->pitch * sizeof(double), rb, ctx->n,
                             cudaMemcpyDeviceToHost, ctx->stream));
    if (lp) CU(cudaMemcpyAsync(lp, ctx->d_lp, ctx->n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return SCORUS_OK;
}

// ------------------------------------------------------------------------------------------
// sample_pt
// ------------------------------------------------------------------------------------------
int scorus::host::upload_beta(scorus_ctx *ctx, const double *beta, size_t nbeta)
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
void scorus::host::stamp_tiles(scorus_ctx *ctx, StepParams *p, const Geometry &g)
{
    p->tile_counter = nullptr;
    if (g.kernel == KERNEL_REG && !p->n_dev && env_int("SCORUS_DYNAMIC_TILES", 1)) {
        p->tile_counter = ctx->d_tile_counter;
        p->tile_base = ctx->tile_ticket;
    }
}

// the device counter moves by ntiles per launch that RAN: advance the host's copy only after a successful launch
void scorus::host::commit_tiles(scorus_ctx *ctx, const StepParams &p)
{
    if (p.tile_counter == ctx->d_tile_counter) ctx->tile_ticket += p.ntiles;
}

void scorus::host::set_matching(const scorus_ctx *ctx, StepParams *p, uint64_t step)
{
    const bool whole = ctx->mblocks <= 1 || (ctx->global_every && step % ctx->global_every == 0);
    p->mblk = whole ? p->npb : p->npb / ctx->mblocks;
    p->mblk_bits = bij_bits(p->mblk);
    p->mblk_off = ctx->mblk_off;
}

// Blocked matchings: on "local" steps every rung is cut into `nblocks` contiguous blocks that pair among themselves;
// with global_every = k >= 1 the steps whose index is a multiple of k use the reference's whole-rung matching
// (ensemble_sample.rs:135-148).  The schedule depends on the step index only, so every step's matching is independent
// of the state: each step is a valid stretch move.  This is what a sharded run does when it exchanges partners only
// every k-th step, stated for one GPU so that the two can be compared bit for bit.
// block_offset shifts the stream keys of the blocks: a shard that IS block c of a larger ensemble passes c.
extern "C" int scorus_set_matching_blocks(scorus_ctx *ctx, uint32_t nblocks, uint32_t block_offset, uint64_t global_every)
{
    if (!ctx || nblocks == 0) return SCORUS_ERR_INVALID_ARG;
    ctx->mblocks = nblocks;
    ctx->mblk_off = block_offset;
    ctx->global_every = global_every;
    return SCORUS_OK;
}

uint32_t scorus::host::flag_bits_for(double prob)
{
    const uint32_t cand[5] = {1, 2, 4, 8, 16};
    for (uint32_t b : cand) {
        double t = prob * (double)(1u << b);
        if (t == floor(t)) return b;
    }
    return 32;
}

// Prob(p): flag d of a walker is (value < thr) on flag_bits random bits.  The threshold is clamped to >= 1 so that
// a tiny p (below 2^-33, which would round to 0) still sets flags -- with thr == 0 no flag could ever be set and
// the redraw loop of ensemble_sample.rs:43-50 would never end.  The device loop is bounded as well
// (kMaxFlagAttempts, tile_common.cuh).
int scorus::host::set_prob_flags(scorus_ctx *ctx, StepParams *p, double prob)
{
    if (!(prob > 0.0)) return fail(ctx, SCORUS_ERR_INVALID_ARG, "Prob(p) needs p > 0 (the reference would loop forever)");
    const double q = fmin(prob, 1.0);
    p->flag_bits = flag_bits_for(q);
    p->flag_thr = p->flag_bits == 32 ? (uint64_t)llrint(q * 4294967296.0) : (uint64_t)(q * (double)(1u << p->flag_bits));
    if (p->flag_thr == 0) p->flag_thr = 1;
    p->flag_cdf = nullptr;
    if (p->flag_bits == 32) {
        // thresholds of the first set flag's position (tile_common.cuh: gen_flags_first), cached per (p, dim):
        // cdf[j] = (1 - (1-p)^(j+1)) / (1 - (1-p)^D) in 32-bit fixed point, via expm1 / log1p so that a tiny p keeps its digits
        if (ctx->flag_cdf_prob != q || !ctx->d_flag_cdf) {
            const uint32_t D = ctx->dim;
            std::vector<uint32_t> cdf(D);
            const double lq = log1p(-q), den = expm1((double)D * lq);
            for (uint32_t j = 0; j < D; ++j) {
                const double t = expm1((double)(j + 1u) * lq) / den * 4294967296.0;
                const long long v = llrint(t);
                cdf[j] = v < 0 ? 0u : (v > 0xFFFFFFFFll ? 0xFFFFFFFFu : (uint32_t)v);
            }
            if (!ctx->d_flag_cdf) CU(cudaMalloc(&ctx->d_flag_cdf, D * sizeof(uint32_t)));
            CU(cudaStreamSynchronize(ctx->stream));  // earlier launches may still read the old table
            CU(cudaMemcpy(ctx->d_flag_cdf, cdf.data(), D * sizeof(uint32_t), cudaMemcpyHostToDevice));
            ctx->flag_cdf_prob = q;
        }
        p->flag_cdf = ctx->d_flag_cdf;
    }
    return SCORUS_OK;
}

// chain statistics: device counters per rung (accepted moves) and per exchange stage (accepted swaps),
// grown on demand; the proposal counts are known on the host
int scorus::host::ensure_rung_stats(scorus_ctx *ctx, size_t nbeta)
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
void scorus::host::count_proposed(scorus_ctx *ctx, size_t nbeta, uint64_t npb, uint64_t w_lo, uint64_t w_hi, uint64_t steps)
{
    ctx->proposed += (w_hi - w_lo) * steps;
    if (ctx->rung_proposed.size() < nbeta) ctx->rung_proposed.resize(nbeta, 0);
    for (uint64_t r = w_lo / npb; r < nbeta && r * npb < w_hi; ++r) {
        const uint64_t lo = std::max(w_lo, r * npb), hi = std::min(w_hi, (r + 1) * npb);
        ctx->rung_proposed[r] += (hi - lo) * steps;
    }
}

void scorus::host::count_swap_proposed(scorus_ctx *ctx, size_t nbeta, uint64_t npb)
{
    ctx->swap_proposed += (nbeta - 1) * npb;
    if (ctx->stage_proposed.size() < nbeta - 1) ctx->stage_proposed.resize(nbeta - 1, 0);
    for (size_t k = 0; k + 1 < nbeta; ++k) ctx->stage_proposed[k] += npb;
}

int scorus::host::fill_common(scorus_ctx *ctx, StepParams *p, Geometry *g, size_t nbeta, bool tile_in_smem,
                              bool keep_pending_swap)
{
    int rc;
    if (!keep_pending_swap && (rc = settle_swap(ctx))) return rc;
    if ((rc = step_geometry(ctx, g, tile_in_smem))) return rc;
    memset(p, 0, sizeof(*p));
    p->ens = ctx->d_ens; p->lp = ctx->d_lp; p->beta = ctx->d_beta; p->pparams = ctx->d_params;
    p->nparams = ctx->nparams; p->stats = ctx->d_stats; p->seed = ctx->seed;
    p->n = (uint32_t)ctx->n; p->npb = (uint32_t)(ctx->n / nbeta); p->dim = ctx->dim; p->pitch = ctx->pitch;
    p->nbeta = (uint32_t)nbeta;
    p->row_bytes = g->row_bytes; p->row_stride = g->row_stride; p->stages = g->stages; p->warps = g->warps;
    p->flag_words = g->flag_words;
    p->ntiles = (uint32_t)((ctx->n + 31u) / 32u);
    p->bij_bits = bij_bits(p->npb);
    p->mblk = p->npb; p->mblk_bits = p->bij_bits; p->mblk_off = ctx->mblk_off;
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

static bool small_geometry(const scorus_ctx *ctx, Geometry *g);

// kernels of swap_kernel.cuh that other translation units launch too
namespace scorus {
namespace host {
void launch_exact_order(scorus_ctx *ctx, uint64_t step, size_t nbeta, uint32_t npb, uint32_t *order, uint32_t w_off)
{
    exact_order_kernel<<<(unsigned)nbeta, (npb + 31u) & ~31u, 0, ctx->stream>>>(ctx->seed, step, npb, order, w_off);
    ++ctx->launches;
}
void launch_exact_jvec(scorus_ctx *ctx, size_t nbeta, size_t npb, uint32_t *jinv)
{
    exact_jvec_kernel<<<(unsigned)(nbeta - 1), ((unsigned)npb + 31u) & ~31u, 0, ctx->stream>>>(
        ctx->seed, ctx->swap_call, (uint32_t)npb, (uint32_t)nbeta, jinv);
    ++ctx->launches;
}
int launch_swap_chain(scorus_ctx *ctx, SwapParams &sp)
{
    // the slot walk over (segment, slot) with the uniforms, their logarithms and the cold log-probs, then the carry
    // sweep, one thread per chain (swap_kernel.cuh)
    const unsigned nseg = (unsigned)((sp.nbeta - 1 + kSwapSeg - 1) / kSwapSeg);
    if (nseg > 65535u) return fail(ctx, SCORUS_ERR_INVALID_ARG, "too many rungs for the swap kernels (%zu)", (size_t)sp.nbeta);
    int rc = ensure(ctx, (void **)&ctx->d_chain, &ctx->chain_cap, (size_t)nseg * sp.npb * sizeof(SwapSeg));
    if (rc) return rc;
    sp.tbl_seg = reinterpret_cast<SwapSeg *>(ctx->d_chain);
    sp.debug_direct_store = env_int("SCORUS_SWAP_DIRECT_STORE", 0) ? 1u : 0u;
    swap_paths_kernel<<<dim3((unsigned)((sp.npb + 127) / 128), nseg), 256, 0, ctx->stream>>>(sp);
    const unsigned T = 128;
    // per-stage exchange counters of a block in shared memory when the ladder is short enough
    if (sp.nbeta - 1 <= kSwapSmemStages)
        swap_walk_kernel<true><<<(unsigned)((sp.npb + T - 1) / T), T, (size_t)(sp.nbeta - 1) * sizeof(unsigned int), ctx->stream>>>(sp);
    else
        swap_walk_kernel<false><<<(unsigned)((sp.npb + T - 1) / T), T, 0, ctx->stream>>>(sp);
    ctx->launches += 2;
    return SCORUS_OK;
}
bool stretch_step_moves_rows(scorus_ctx *ctx, size_t nbeta)
{
    if (ctx->f32 || ctx->sequential || ctx->d_peer_ens || ctx->lp_kind < 0 || nbeta == 0 || ctx->n % nbeta) return false;
    Geometry g, gs;
    if (step_geometry(ctx, &g, false) != SCORUS_OK) return false;
    gs = g;
    return !small_geometry(ctx, &gs) && step_can_remap(ctx->lp_kind, ctx->pitch, g) && g.grid > 1;
}
int settle_swap(scorus_ctx *ctx)
{
    if (!ctx->swap_pending) return SCORUS_OK;
    cudaSetDevice(ctx->device);
    const uint32_t vpr = ctx->pitch / 2;
    const size_t total = ctx->n * (size_t)vpr;
    const unsigned blocks = (unsigned)std::min<size_t>((total + 255) / 256, (size_t)ctx->num_sms * 16u);
    apply_swap_kernel<<<blocks, 256, 0, ctx->stream>>>((const double2 *)ctx->d_ens, (double2 *)ctx->d_scratch,
                                                      (const SwapOut *)ctx->d_swap_out, ctx->d_lp, (uint32_t)ctx->n, vpr);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(ctx, SCORUS_ERR_CUDA, "swap row move: %s", cudaGetErrorString(e));
    std::swap(ctx->d_ens, ctx->d_scratch);  // the second buffer becomes the ensemble
    ctx->swap_pending = false;
    ++ctx->launches;
    return SCORUS_OK;
}
}  // namespace host
}  // namespace scorus

// One-block ensembles (step_small_kernel.cuh): the whole ensemble in shared memory, any number of steps per launch.
// Chosen by size alone -- never by the number of steps of a call -- so that k calls of one step and one call of k
// steps are the same arithmetic.  SCORUS_KERNEL=reg|seq|coop keeps the streaming kernels (tests, measurements).
static bool small_geometry(const scorus_ctx *ctx, Geometry *g)
{
    const char *force = getenv("SCORUS_KERNEL");
    if (ctx->sequential || ctx->n > 1024 || ctx->mblocks > 1 || !env_int("SCORUS_SMALL", 1)) return false;
    if (force && *force && strcmp(force, "small")) return false;
    const uint32_t warps = (uint32_t)((ctx->n + 31u) / 32u);
    const SmallLayout L = small_layout((uint32_t)ctx->n, ctx->pitch, g->flag_words, warps);
    if (L.total > (size_t)ctx->smem_optin) return false;
    g->kernel = KERNEL_SMALL;
    g->warps = warps;
    g->grid = 1;
    g->smem = L.total;
    g->pdl = 0;
    return true;
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
    if (ctx->f32) {
        if (ctx->mblocks > 1) return fail(ctx, SCORUS_ERR_INVALID_ARG, "blocked matchings are not available on an f32 context");
        if (ctx->tr_ncap != 0 || ctx->mom_rungs != 0) return fail(ctx, SCORUS_ERR_INVALID_ARG, "chain capture is not available on an f32 context");
        return f32_sample_pt(ctx, a, ufs, beta, nbeta, nsteps);
    }

    StepParams p;
    Geometry g;
    if ((rc = upload_beta(ctx, beta, nbeta))) return rc;
    if ((rc = fill_common(ctx, &p, &g, nbeta, false, true))) return rc;
    // a ladder swap left its rows to this call (swap_impl): the first step moves them itself when its kernel can
    // (single launch over the whole ensemble, register kernel in direct mode), otherwise they move now
    bool move_rows = false;
    // host-buffer call on pinned buffers: the first step streams the rows from the caller's ensemble (api_host.cu
    // checked stretch_step_moves_rows and that nothing is pending)
    double *const host_ens = ctx->host_ens, *const host_lp = ctx->host_lp;
    ctx->host_ens = ctx->host_lp = nullptr;
    if (host_ens && (ctx->swap_pending || nsteps == 0 || !stretch_step_moves_rows(ctx, nbeta)))
        return fail(ctx, SCORUS_ERR_INVALID_ARG, "internal: this call cannot stream its rows from the host");
    if (ctx->swap_pending) {
        Geometry gs = g;
        move_rows = !small_geometry(ctx, &gs) && step_can_remap(ctx->lp_kind, ctx->pitch, g) && g.grid > 1 &&
                    !ctx->track_dirty && nsteps > 0 && env_int("SCORUS_SWAP_FUSE", 1);
        if (!move_rows) {
            if ((rc = settle_swap(ctx))) return rc;
            p.ens = ctx->d_ens; p.src_ens = ctx->d_ens;
        }
    }
    const double sqrt_a = sqrt(a);
    p.inv_sqrt_a = 1.0 / sqrt_a;
    p.z_range = 2.0 * (sqrt_a - p.inv_sqrt_a);  // src/mcmc/utils.rs:42
    p.flag_kind = (uint32_t)ufs->kind;
    bool all_flags = false;
    switch (ufs->kind) {
    case SCORUS_UFS_ALL: all_flags = true; break;
    case SCORUS_UFS_PROB:
        if ((rc = set_prob_flags(ctx, &p, ufs->prob))) return rc;
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

    if (small_geometry(ctx, &g)) {
        // without per-step hooks (chain capture, running moments) all nsteps go down in one launch
        const bool hooks = ctx->tr_ncap != 0 || ctx->mom_rungs != 0;
        uint64_t done = 0;
        while (done < nsteps) {
            const uint32_t inner = hooks ? 1u : (uint32_t)std::min<uint64_t>(nsteps - done, 1u << 20);
            p.step = ctx->step;
            p.onehot = ctx->onehot;
            cudaError_t e = launch_small(ctx->lp_kind, p, g, inner, ctx->stream);
            if (e != cudaSuccess) return fail(ctx, SCORUS_ERR_CUDA, "step launch: %s", cudaGetErrorString(e));
            ++ctx->launches;
            ctx->step += inner;
            if (ufs->kind == SCORUS_UFS_ONE_HOT_CYCLE)
                ctx->onehot = (ctx->onehot + (uint64_t)inner * (ctx->n_global % ctx->dim)) % ctx->dim;
            done += inner;
            if (hooks && (rc = after_step(ctx))) return rc;
        }
        count_proposed(ctx, nbeta, p.npb, 0, ctx->n, nsteps);
        return SCORUS_OK;
    }

    if (ctx->mblocks > 1 && (p.npb % ctx->mblocks || (p.npb / ctx->mblocks) % 2))
        return fail(ctx, SCORUS_ERR_NWALKERS_IS_NOT_EVEN, "a rung of %u walkers does not split into %u even matching blocks", p.npb, ctx->mblocks);
    // one-block ensembles: the register kernel shuffles the rungs itself (step_reg_kernel.cuh)
    const size_t fused_bytes = ctx->n * (sizeof(uint64_t) + sizeof(uint32_t));
    if (p.npb <= kExactShuffleMax && ctx->mblocks == 1 && g.kernel == KERNEL_REG && g.grid == 1 &&
        g.smem + fused_bytes <= (size_t)ctx->smem_optin && env_int("SCORUS_FUSED_ORDER", 1)) {
        p.fused_order = 1;
        g.smem += fused_bytes;
    }
    for (uint64_t s = 0; s < nsteps; ++s) {
        p.step = ctx->step;
        p.onehot = ctx->onehot;
        set_matching(ctx, &p, ctx->step);
        p.order = nullptr;
        if (p.mblk <= kExactShuffleMax && !p.fused_order) {  // exact uniform shuffle of every matching block
            if (!ctx->d_order) CU(cudaMalloc(&ctx->d_order, ctx->n * sizeof(uint32_t)));
            launch_exact_order(ctx, ctx->step, ctx->n / p.mblk, p.mblk, ctx->d_order, (uint32_t)ctx->w_off);
            p.order = ctx->d_order;
        }
        if (move_rows) {  // read through the swap's result, write every row to the second buffer
            p.remap = ctx->d_swap_out;
            p.src_ens = ctx->d_ens;
            p.ens = ctx->d_scratch;
        }
        const uint32_t l2_prefetch = p.l2_prefetch;
        if (host_ens && s == 0) {  // rows straight from the caller's pinned buffer, accepted rows straight back
            p.src_ens = host_ens;
            p.host_ens = host_ens;
            p.host_lp = host_lp;
            p.l2_prefetch = 0;
        }
        stamp_tiles(ctx, &p, g);
        cudaError_t e = launch_step(ctx->lp_kind, p, g, all_flags, ctx->stream);
        if (e != cudaSuccess) return fail(ctx, SCORUS_ERR_CUDA, "step launch: %s", cudaGetErrorString(e));
        commit_tiles(ctx, p);
        if (move_rows) {  // ... which is the ensemble from here on
            std::swap(ctx->d_ens, ctx->d_scratch);
            ctx->swap_pending = false;
            move_rows = false;
            p.remap = nullptr;
            p.ens = ctx->d_ens; p.src_ens = ctx->d_ens;
        }
        if (host_ens && s == 0) {  // the device copy is complete from here on
            p.src_ens = ctx->d_ens;
            p.host_ens = p.host_lp = nullptr;
            p.l2_prefetch = l2_prefetch;
        }
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
    F64_ONLY(ctx);
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
    cudaError_t e;
    if (small_geometry(ctx, &g)) {
        e = launch_small(ctx->lp_kind, p, g, 1, ctx->stream);
    } else {
        stamp_tiles(ctx, &p, g);
        e = launch_step(ctx->lp_kind, p, g, false, ctx->stream);
        if (e == cudaSuccess) commit_tiles(ctx, p);
    }
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
    F64_ONLY(ctx);
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

    if ((rc = settle_swap(ctx))) return rc;  // a swap right after a swap
    SwapParams sp;
    memset(&sp, 0, sizeof(sp));
    if (!ctx->d_scratch) CU(cudaMalloc(&ctx->d_scratch, n * (size_t)ctx->pitch * sizeof(double)));
    // buffers mapped by peers (scorus_ipc_attach) stay put: the rows move in place, through scratch; otherwise the
    // sweep's result stays interleaved and the rows move with the next step (or when anything else looks, settle_swap)
    const bool in_place = ctx->d_peer_ens != nullptr;
    if (in_place) {
        if (!ctx->d_src) CU(cudaMalloc(&ctx->d_src, n * sizeof(uint32_t)));
        sp.lp = ctx->d_lp; sp.src = ctx->d_src;
    } else {
        if ((rc = ensure(ctx, &ctx->d_swap_out, &ctx->swap_out_cap, n * sizeof(SwapOut)))) return rc;
        sp.out = (SwapOut *)ctx->d_swap_out;
    }
    sp.lp_src = ctx->d_lp; sp.beta = ctx->d_beta; sp.stats = ctx->d_stats;
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
        launch_exact_jvec(ctx, nbeta, npb, ctx->d_jinv);
        sp.jinv = ctx->d_jinv;
    }
    if (swapped_out) {
        if ((rc = ensure(ctx, (void **)&ctx->d_swapped, &ctx->swapped_cap, m))) return rc;
        sp.swapped_out = ctx->d_swapped;
    }
    if ((rc = launch_swap_chain(ctx, sp))) return rc;
    if (in_place) {
        const uint32_t vpr = ctx->pitch / 2;
        const size_t total = n * (size_t)vpr;
        const unsigned blocks = (unsigned)std::min<size_t>((total + 255) / 256, (size_t)ctx->num_sms * 16u);
        permute_rows_kernel<<<blocks, 256, 0, ctx->stream>>>((double2 *)ctx->d_ens, (double2 *)ctx->d_scratch, ctx->d_src,
                                                             (uint32_t)n, vpr, 0);
        permute_rows_kernel<<<blocks, 256, 0, ctx->stream>>>((double2 *)ctx->d_ens, (double2 *)ctx->d_scratch, ctx->d_src,
                                                             (uint32_t)n, vpr, 1);
        ctx->launches += 2;
    } else {
        // the rows go to their new slots in the second buffer, which then becomes the ensemble -- one pass, by the next
        // step itself (SCORUS_SWAP_FUSE=0: right away; the first version gathered the moved rows into scratch and
        // copied them back: two passes -- at c4, where 87 % of the exchanges are accepted, nearly every row twice)
        ctx->swap_pending = true;
        if (!env_int("SCORUS_SWAP_FUSE", 1) && (rc = settle_swap(ctx))) return rc;
    }
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
    if (ctx->f32) return f32_swap_walkers(ctx, beta, nbeta);
    return swap_impl(ctx, beta, nbeta, nullptr, nullptr, nullptr);
}

extern "C" int scorus_swap_walkers_fixed(scorus_ctx *ctx, const double *beta, size_t nbeta, const uint32_t *jvec,
                                         const double *r, uint8_t *swapped_out)
{
    if (!ctx || !beta) return SCORUS_ERR_INVALID_ARG;
    F64_ONLY(ctx);
    if (nbeta > 1 && (!jvec || !r)) return SCORUS_ERR_INVALID_ARG;
    return swap_impl(ctx, beta, nbeta, nbeta > 1 ? jvec : nullptr, r, swapped_out);
}
