// api_f32.cu -- C ABI, part 6: T = f32 contexts (SCORUS_CFG_F32).  The reference is generic over the walkers' scalar
// type (`T: Float`, src/mcmc/ensemble_sample.rs:95, src/mcmc/utils.rs:72-85); the kernels here are the f32 instantiation of
// K1 (the fused register step), K2 (init_logprob) and K5 (the ladder swap), produced from the same sources as the f64
// ones (gen_f32.py -> namespace scorus::f32).  This file is their host side: parameters filled in float, host arrays in
// float, everything else (validation, counters, streams, matching orders, statistics counters) shared with f64.
#include "host_common.cuh"

#include "gen_f32/swap_kernel.cuh"   // scorus::f32::{SwapParams, swap_paths_kernel, swap_walk_kernel, gather_rows_kernel}
#include "gen_f32/tile_common.cuh"   // scorus::f32::StepParams

namespace scorus {
namespace f32 {
#define SCORUS_F32_DECLARE(NAME)                                                                              \
    cudaError_t launch_step_##NAME(const StepParams &p, const Geometry &g, bool all_flags, cudaStream_t st);  \
    cudaError_t launch_init_##NAME(const StepParams &p, uint32_t grid, size_t smem, cudaStream_t st);
SCORUS_F32_DECLARE(LpGaussian)
SCORUS_F32_DECLARE(LpBoundedGaussian)
SCORUS_F32_DECLARE(LpRosenbrock)
SCORUS_F32_DECLARE(LpRastrigin)
SCORUS_F32_DECLARE(LpBimodal2d)
SCORUS_F32_DECLARE(LpStep2d)
SCORUS_F32_DECLARE(LpMixture)

static cudaError_t launch_step(int kind, const StepParams &p, const Geometry &g, bool all_flags, cudaStream_t st)
{
    switch (kind) {
    case SCORUS_LP_GAUSSIAN: return launch_step_LpGaussian(p, g, all_flags, st);
    case SCORUS_LP_BOUNDED_GAUSSIAN: return launch_step_LpBoundedGaussian(p, g, all_flags, st);
    case SCORUS_LP_ROSENBROCK: return launch_step_LpRosenbrock(p, g, all_flags, st);
    case SCORUS_LP_RASTRIGIN: return launch_step_LpRastrigin(p, g, all_flags, st);
    case SCORUS_LP_BIMODAL2D: return launch_step_LpBimodal2d(p, g, all_flags, st);
    case SCORUS_LP_STEP2D: return launch_step_LpStep2d(p, g, all_flags, st);
    case SCORUS_LP_MIXTURE: return launch_step_LpMixture(p, g, all_flags, st);
    default: return cudaErrorNotSupported;  // the dense Gaussian is a tensor-core fp64 kernel
    }
}
static cudaError_t launch_init(int kind, const StepParams &p, uint32_t grid, size_t smem, cudaStream_t st)
{
    switch (kind) {
    case SCORUS_LP_GAUSSIAN: return launch_init_LpGaussian(p, grid, smem, st);
    case SCORUS_LP_BOUNDED_GAUSSIAN: return launch_init_LpBoundedGaussian(p, grid, smem, st);
    case SCORUS_LP_ROSENBROCK: return launch_init_LpRosenbrock(p, grid, smem, st);
    case SCORUS_LP_RASTRIGIN: return launch_init_LpRastrigin(p, grid, smem, st);
    case SCORUS_LP_BIMODAL2D: return launch_init_LpBimodal2d(p, grid, smem, st);
    case SCORUS_LP_STEP2D: return launch_init_LpStep2d(p, grid, smem, st);
    case SCORUS_LP_MIXTURE: return launch_init_LpMixture(p, grid, smem, st);
    default: return cudaErrorNotSupported;
    }
}
}  // namespace f32
}  // namespace scorus

namespace F = scorus::f32;

// plugin parameters: double in the ABI, float on the device (called by scorus_set_logprob)
int scorus::host::f32_set_params(scorus_ctx *ctx, const double *params, size_t nparams)
{
    if (ctx->lp_kind == SCORUS_LP_CORR_GAUSSIAN)
        return fail(ctx, SCORUS_ERR_INVALID_ARG, "the dense Gaussian runs on the fp64 tensor cores: no f32 instantiation");
    std::vector<float> pf(nparams ? nparams : 1);
    for (size_t i = 0; i < nparams; ++i) pf[i] = (float)params[i];
    int rc = ensure(ctx, (void **)&ctx->d_params, &ctx->params_cap, pf.size() * sizeof(double));
    if (rc) return rc;
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaMemcpy(ctx->d_params, pf.data(), pf.size() * sizeof(float), cudaMemcpyHostToDevice));
    return SCORUS_OK;
}

// the beta list as float on the device (small; re-uploaded when it changes)
static int f32_upload_beta(scorus_ctx *ctx, const double *beta, size_t nbeta)
{
    if (ctx->beta_f32_host.size() == nbeta && memcmp(ctx->beta_f32_src.data(), beta, nbeta * sizeof(double)) == 0) return SCORUS_OK;
    CU(cudaStreamSynchronize(ctx->stream));  // earlier launches may still read the old list
    ctx->beta_f32_src.assign(beta, beta + nbeta);
    ctx->beta_f32_host.resize(nbeta);
    for (size_t i = 0; i < nbeta; ++i) ctx->beta_f32_host[i] = (float)beta[i];
    int rc = ensure(ctx, (void **)&ctx->d_beta_f32, &ctx->beta_f32_cap, nbeta * sizeof(float));
    if (rc) return rc;
    CU(cudaMemcpy(ctx->d_beta_f32, ctx->beta_f32_host.data(), nbeta * sizeof(float), cudaMemcpyHostToDevice));
    return SCORUS_OK;
}

static int f32_fill_common(scorus_ctx *ctx, F::StepParams *p, Geometry *g, size_t nbeta)
{
    int rc = step_geometry(ctx, g, false);
    if (rc) return rc;
    if (g->kernel != KERNEL_REG) return fail(ctx, SCORUS_ERR_DIM_TOO_LARGE, "f32 contexts run the register kernel: dim <= 256");
    memset(p, 0, sizeof(*p));
    p->ens = reinterpret_cast<float *>(ctx->d_ens);
    p->lp = reinterpret_cast<float *>(ctx->d_lp);
    p->beta = ctx->d_beta_f32;
    p->pparams = reinterpret_cast<const float *>(ctx->d_params);
    p->nparams = ctx->nparams; p->stats = ctx->d_stats; p->seed = ctx->seed;
    p->n = (uint32_t)ctx->n; p->npb = (uint32_t)(ctx->n / nbeta); p->dim = ctx->dim; p->pitch = ctx->pitch;
    p->nbeta = (uint32_t)nbeta;
    p->row_bytes = g->row_bytes; p->row_stride = g->row_stride; p->stages = g->stages; p->warps = g->warps;
    p->flag_words = g->flag_words;
    p->ntiles = (uint32_t)((ctx->n + 31u) / 32u);
    p->bij_bits = bij_bits(p->npb);
    p->mblk = p->npb; p->mblk_bits = p->bij_bits; p->mblk_off = 0;
    p->n_global = ctx->n;
    p->src_ens = p->ens; p->src_lp = p->lp; p->w_lo = 0; p->w_hi = (uint32_t)ctx->n;
    const int pf = env_int("SCORUS_L2_PREFETCH", -1);
    p->l2_prefetch = pf < 0 ? (g->row_bytes >= 128u && g->row_bytes <= 256u ? 1u : 0u) : (pf ? 1u : 0u);
    if (nbeta > 1) {
        if ((rc = ensure_rung_stats(ctx, nbeta))) return rc;
        p->rung_acc = ctx->d_rung_stats;
    }
    return SCORUS_OK;
}

static void f32_stamp_tiles(scorus_ctx *ctx, F::StepParams *p)
{
    p->tile_counter = nullptr;
    if (env_int("SCORUS_DYNAMIC_TILES", 1)) {
        p->tile_counter = ctx->d_tile_counter;
        p->tile_base = ctx->tile_ticket;
    }
}

int scorus::host::f32_init_logprob(scorus_ctx *ctx)
{
    Geometry g;
    int rc = step_geometry(ctx, &g, false);
    if (rc) return rc;
    F::StepParams p;
    memset(&p, 0, sizeof(p));
    p.ens = reinterpret_cast<float *>(ctx->d_ens); p.lp = reinterpret_cast<float *>(ctx->d_lp);
    p.pparams = reinterpret_cast<const float *>(ctx->d_params); p.nparams = ctx->nparams;
    p.n = (uint32_t)ctx->n; p.dim = ctx->dim; p.pitch = ctx->pitch;
    p.row_bytes = g.row_bytes; p.row_stride = g.row_stride;
    p.ntiles = (uint32_t)((ctx->n + 31u) / 32u);
    const size_t tile = 32u * (size_t)g.row_stride;
    uint32_t w = 16;
    while (w > 1 && ((8u * w + 127u) & ~127u) + w * tile > (size_t)ctx->smem_optin) --w;
    p.warps = w;
    const size_t smem = ((8u * w + 127u) & ~127u) + w * tile;
    uint32_t grid = (p.ntiles + w - 1) / w;
    const uint32_t max_grid = (uint32_t)ctx->num_sms * 4u;
    if (grid > max_grid) grid = max_grid;
    cudaError_t e = F::launch_init(ctx->lp_kind, p, grid, smem, ctx->stream);
    if (e != cudaSuccess) return fail(ctx, SCORUS_ERR_CUDA, "init_logprob launch (f32): %s", cudaGetErrorString(e));
    ++ctx->launches;
    return SCORUS_OK;
}

// z range in T = f32 arithmetic, as `draw_z::<f32>` computes it (src/mcmc/utils.rs:39-44)
static void f32_z_constants(double a, F::StepParams *p)
{
    const float sa = sqrtf((float)a);
    p->inv_sqrt_a = 1.0f / sa;
    p->z_range = 2.0f * (sa - p->inv_sqrt_a);
}

int scorus::host::f32_sample_pt(scorus_ctx *ctx, double a, const scorus_ufs *ufs, const double *beta, size_t nbeta,
                                uint64_t nsteps)
{
    int rc;
    F::StepParams p;
    Geometry g;
    if ((rc = f32_upload_beta(ctx, beta, nbeta))) return rc;
    if ((rc = f32_fill_common(ctx, &p, &g, nbeta))) return rc;
    f32_z_constants(a, &p);
    p.flag_kind = (uint32_t)ufs->kind;
    bool all_flags = false;
    switch (ufs->kind) {
    case SCORUS_UFS_ALL: all_flags = true; break;
    case SCORUS_UFS_PROB: {
        StepParams q;  // the threshold is the same quantisation of p for both scalar types
        if ((rc = set_prob_flags(ctx, &q, ufs->prob))) return rc;
        p.flag_bits = q.flag_bits; p.flag_thr = q.flag_thr; p.flag_cdf = q.flag_cdf;
        break;
    }
    case SCORUS_UFS_ONE_HOT_CYCLE: break;
    case SCORUS_UFS_MASK: {
        if (!ufs->mask || ufs->mask_len > ctx->dim) return fail(ctx, SCORUS_ERR_INVALID_ARG, "bad mask");
        if (nsteps != 1) return fail(ctx, SCORUS_ERR_INVALID_ARG, "a host-evaluated mask covers exactly one step");
        for (size_t i = 0; i < ctx->n; ++i) {
            size_t c = 0;
            for (size_t k = 0; k < ufs->mask_len; ++k) c += ufs->mask[i * ufs->mask_len + k] ? 1 : 0;
            if (c == 0) return fail(ctx, SCORUS_ERR_NPHI_ZERO, "walker %zu has no update flag set", i);
        }
        const size_t bytes = (size_t)ctx->n * ufs->mask_len;
        if ((rc = ensure(ctx, (void **)&ctx->d_flags, &ctx->flags_cap, bytes))) return rc;
        CU(cudaMemcpyAsync(ctx->d_flags, ufs->mask, bytes, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        p.flags_in = ctx->d_flags;
        p.flag_len = (uint32_t)ufs->mask_len;
        break;
    }
    default: return fail(ctx, SCORUS_ERR_INVALID_ARG, "unknown UpdateFlagSpec kind %d", ufs->kind);
    }
    const bool exact = p.npb <= kExactShuffleMax;
    if (exact && !ctx->d_order) CU(cudaMalloc(&ctx->d_order, ctx->n * sizeof(uint32_t)));
    for (uint64_t s = 0; s < nsteps; ++s) {
        p.step = ctx->step;
        p.onehot = ctx->onehot;
        if (exact) {
            launch_exact_order(ctx, ctx->step, nbeta, p.npb, ctx->d_order, 0u);
            p.order = ctx->d_order;
        }
        f32_stamp_tiles(ctx, &p);
        cudaError_t e = F::launch_step(ctx->lp_kind, p, g, all_flags, ctx->stream);
        if (e != cudaSuccess) return fail(ctx, SCORUS_ERR_CUDA, "step launch (f32): %s", cudaGetErrorString(e));
        if (p.tile_counter) ctx->tile_ticket += p.ntiles;
        ++ctx->launches;
        ++ctx->step;
        if (ufs->kind == SCORUS_UFS_ONE_HOT_CYCLE) ctx->onehot = (ctx->onehot + ctx->n) % ctx->dim;
    }
    count_proposed(ctx, nbeta, p.npb, 0, ctx->n, nsteps);
    return SCORUS_OK;
}

extern "C" int scorus_upload_f32(scorus_ctx *ctx, const float *ens, const float *lp)
{
    if (!ctx || !ens) return SCORUS_ERR_INVALID_ARG;
    if (!ctx->f32) return fail(ctx, SCORUS_ERR_INVALID_ARG, "scorus_upload_f32 needs a SCORUS_CFG_F32 context");
    cudaSetDevice(ctx->device);
    const size_t rb = (size_t)ctx->dim * sizeof(float);
    if (ctx->pitch == ctx->dim)  // rows without padding: one flat copy
        CU(cudaMemcpyAsync(ctx->d_ens, ens, rb * ctx->n, cudaMemcpyHostToDevice, ctx->stream));
    else
        CU(cudaMemcpy2DAsync(ctx->d_ens, (size_t)ctx->pitch * sizeof(float), ens, rb, rb, ctx->n, cudaMemcpyHostToDevice, ctx->stream));
    ctx->uploaded = true;
    if (lp) {
        CU(cudaMemcpyAsync(ctx->d_lp, lp, ctx->n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
        return SCORUS_OK;
    }
    return scorus_init_logprob(ctx);
}

extern "C" int scorus_download_f32(scorus_ctx *ctx, float *ens, float *lp)
{
    if (!ctx) return SCORUS_ERR_INVALID_ARG;
    if (!ctx->f32) return fail(ctx, SCORUS_ERR_INVALID_ARG, "scorus_download_f32 needs a SCORUS_CFG_F32 context");
    if (!ctx->uploaded) return fail(ctx, SCORUS_ERR_NOT_UPLOADED, "nothing uploaded");
    cudaSetDevice(ctx->device);
    const size_t rb = (size_t)ctx->dim * sizeof(float);
    if (ens)
    {
        if (ctx->pitch == ctx->dim)
            CU(cudaMemcpyAsync(ens, ctx->d_ens, rb * ctx->n, cudaMemcpyDeviceToHost, ctx->stream));
        else
            CU(cudaMemcpy2DAsync(ens, rb, ctx->d_ens, (size_t)ctx->pitch * sizeof(float), rb, ctx->n, cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (lp) CU(cudaMemcpyAsync(lp, ctx->d_lp, ctx->n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return SCORUS_OK;
}

extern "C" int scorus_sample_pt_fixed_f32(scorus_ctx *ctx, const double *beta, size_t nbeta, const uint32_t *partner,
                                          const float *z, const uint8_t *flags, size_t flag_len, const float *u,
                                          uint8_t *accepted_out)
{
    if (!ctx || !beta || !partner || !z || !flags || !u) return SCORUS_ERR_INVALID_ARG;
    if (!ctx->f32) return fail(ctx, SCORUS_ERR_INVALID_ARG, "scorus_sample_pt_fixed_f32 needs a SCORUS_CFG_F32 context");
    int rc = scorus_check_sample_args(ctx->n, ctx->n, nbeta);
    if (rc) return fail(ctx, rc, "%s", scorus_status_string(rc));
    if (ctx->lp_kind < 0) return fail(ctx, SCORUS_ERR_NO_LOGPROB, "call scorus_set_logprob first");
    if (!ctx->uploaded) return fail(ctx, SCORUS_ERR_NOT_UPLOADED, "call scorus_upload first");
    if (flag_len > ctx->dim) return fail(ctx, SCORUS_ERR_INVALID_ARG, "flag_len > dim (index panic in the reference)");
    cudaSetDevice(ctx->device);
    const size_t n = ctx->n, npb = n / nbeta;
    std::vector<uint32_t> order(n);
    size_t k = 0;
    for (size_t i = 0; i < n; ++i) {  // a fixed-point-free involution inside each rung (ensemble_sample.rs:135-148)
        const size_t q = partner[i];
        if (q >= n || q == i || q / npb != i / npb || partner[q] != i)
            return fail(ctx, SCORUS_ERR_INVALID_ARG, "partner[] is not a within-rung involution at walker %zu", i);
        if (i < q) { order[k++] = (uint32_t)i; order[k++] = (uint32_t)q; }
    }
    for (size_t i = 0; i < n; ++i) {
        size_t c = 0;
        for (size_t d = 0; d < flag_len; ++d) c += flags[i * flag_len + d] ? 1 : 0;
        if (c == 0) return fail(ctx, SCORUS_ERR_NPHI_ZERO, "walker %zu has no update flag set", i);
    }
    F::StepParams p;
    Geometry g;
    if ((rc = f32_upload_beta(ctx, beta, nbeta))) return rc;
    if ((rc = f32_fill_common(ctx, &p, &g, nbeta))) return rc;
    if (!ctx->d_order) CU(cudaMalloc(&ctx->d_order, n * sizeof(uint32_t)));
    if (!ctx->d_z) CU(cudaMalloc(&ctx->d_z, n * sizeof(double)));
    if (!ctx->d_u) CU(cudaMalloc(&ctx->d_u, n * sizeof(double)));
    if (!ctx->d_acc) CU(cudaMalloc(&ctx->d_acc, n));
    if ((rc = ensure(ctx, (void **)&ctx->d_flags, &ctx->flags_cap, n * (flag_len ? flag_len : 1)))) return rc;
    CU(cudaMemcpyAsync(ctx->d_order, order.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_z, z, n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_u, u, n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_flags, flags, n * flag_len, cudaMemcpyHostToDevice, ctx->stream));
    p.order = ctx->d_order;
    p.z_in = reinterpret_cast<const float *>(ctx->d_z);
    p.u_in = reinterpret_cast<const float *>(ctx->d_u);
    p.flags_in = ctx->d_flags;
    p.flag_len = (uint32_t)flag_len; p.flag_kind = SCORUS_UFS_MASK; p.accepted_out = ctx->d_acc;
    p.step = ctx->step;
    f32_stamp_tiles(ctx, &p);
    cudaError_t e = F::launch_step(ctx->lp_kind, p, g, false, ctx->stream);
    if (e != cudaSuccess) return fail(ctx, SCORUS_ERR_CUDA, "step launch (f32): %s", cudaGetErrorString(e));
    if (p.tile_counter) ctx->tile_ticket += p.ntiles;
    ++ctx->launches;
    count_proposed(ctx, nbeta, p.npb, 0, n, 1);
    if (accepted_out) CU(cudaMemcpyAsync(accepted_out, ctx->d_acc, n, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return SCORUS_OK;
}

// ------------------------------------------------------------------------------------------
// swap_walkers at T = f32
// ------------------------------------------------------------------------------------------
static int f32_swap_impl(scorus_ctx *ctx, const double *beta, size_t nbeta, const uint32_t *jvec, const float *r,
                         uint8_t *swapped_out)
{
    const size_t n = ctx->n;
    if (nbeta == 0) return fail(ctx, SCORUS_ERR_INVALID_ARG, "nbeta == 0");
    const size_t npb = n / nbeta;
    if (npb * nbeta != n) return fail(ctx, SCORUS_ERR_NWALKERS_MISMATCHES_NBETA, "NWalkersMismatchesNBeta");
    int rc = scorus_check_beta_order(beta, nbeta);
    if (rc) return fail(ctx, rc, "beta list must be in decreasing order");
    if (!ctx->uploaded) return fail(ctx, SCORUS_ERR_NOT_UPLOADED, "call scorus_upload first");
    cudaSetDevice(ctx->device);
    if (nbeta == 1) { ++ctx->swap_call; return SCORUS_OK; }
    if ((rc = f32_upload_beta(ctx, beta, nbeta))) return rc;
    const size_t m = (nbeta - 1) * npb;
    F::SwapParams sp;
    memset(&sp, 0, sizeof(sp));
    if (!ctx->d_src) CU(cudaMalloc(&ctx->d_src, n * sizeof(uint32_t)));
    if (!ctx->d_scratch) CU(cudaMalloc(&ctx->d_scratch, n * (size_t)ctx->pitch * sizeof(float)));
    sp.lp = reinterpret_cast<float *>(ctx->d_lp); sp.lp_src = sp.lp; sp.beta = ctx->d_beta_f32; sp.src = ctx->d_src;
    sp.stats = ctx->d_stats;
    if ((rc = ensure_rung_stats(ctx, nbeta))) return rc;
    sp.stage_swapped = ctx->d_rung_stats + ctx->rung_cap;
    sp.seed = ctx->seed; sp.swap_call = ctx->swap_call;
    sp.npb = (uint32_t)npb; sp.nbeta = (uint32_t)nbeta; sp.bij_bits = bij_bits((uint32_t)npb);
    std::vector<uint32_t> jinv;
    if (jvec) {
        jinv.assign(m, 0xFFFFFFFFu);
        for (size_t s = 0; s + 1 < nbeta; ++s)
            for (size_t j2 = 0; j2 < npb; ++j2) {
                const uint32_t j1 = jvec[s * npb + j2];
                if (j1 >= npb || jinv[s * npb + j1] != 0xFFFFFFFFu)
                    return fail(ctx, SCORUS_ERR_INVALID_ARG, "jvec stage %zu is not a permutation", s);
                jinv[s * npb + j1] = (uint32_t)j2;
            }
        if ((rc = ensure(ctx, (void **)&ctx->d_jinv, &ctx->jinv_cap, m * sizeof(uint32_t)))) return rc;
        if ((rc = ensure(ctx, (void **)&ctx->d_r, &ctx->r_cap, m * sizeof(double)))) return rc;
        CU(cudaMemcpyAsync(ctx->d_jinv, jinv.data(), m * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(ctx->d_r, r, m * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
        sp.jinv = ctx->d_jinv;
        sp.r_in = reinterpret_cast<const float *>(ctx->d_r);
    } else if (npb <= kExactShuffleMax) {
        if ((rc = ensure(ctx, (void **)&ctx->d_jinv, &ctx->jinv_cap, m * sizeof(uint32_t)))) return rc;
        launch_exact_jvec(ctx, nbeta, npb, ctx->d_jinv);
        sp.jinv = ctx->d_jinv;
    }
    if (swapped_out) {
        if ((rc = ensure(ctx, (void **)&ctx->d_swapped, &ctx->swapped_cap, m))) return rc;
        sp.swapped_out = ctx->d_swapped;
    }
    const unsigned nseg = (unsigned)((nbeta - 1 + F::kSwapSeg - 1) / F::kSwapSeg);
    if (nseg > 65535u) return fail(ctx, SCORUS_ERR_INVALID_ARG, "too many rungs for the swap kernels (%zu)", (size_t)nbeta);
    rc = ensure(ctx, (void **)&ctx->d_chain, &ctx->chain_cap, (size_t)nseg * npb * sizeof(F::SwapSeg));
    if (rc) return rc;
    sp.tbl_seg = reinterpret_cast<F::SwapSeg *>(ctx->d_chain);
    F::swap_paths_kernel<<<dim3((unsigned)((npb + 127) / 128), nseg), 256, 0, ctx->stream>>>(sp);
    const unsigned T = 128;
    // per-stage exchange counters of a block in shared memory when the ladder is short enough
    if (nbeta - 1 <= F::kSwapSmemStages)
        F::swap_walk_kernel<true><<<(unsigned)((npb + T - 1) / T), T, (size_t)(nbeta - 1) * sizeof(unsigned int), ctx->stream>>>(sp);
    else
        F::swap_walk_kernel<false><<<(unsigned)((npb + T - 1) / T), T, 0, ctx->stream>>>(sp);
    const uint32_t vpr = ctx->pitch / 2;  // float2 chunks per row
    const size_t total = n * (size_t)vpr;
    const unsigned blocks = (unsigned)std::min<size_t>((total + 255) / 256, (size_t)ctx->num_sms * 16u);
    F::gather_rows_kernel<<<blocks, 256, 0, ctx->stream>>>((const float2 *)ctx->d_ens, (float2 *)ctx->d_scratch, ctx->d_src,
                                                           (uint32_t)n, vpr);
    std::swap(ctx->d_ens, ctx->d_scratch);
    ctx->launches += 3;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(ctx, SCORUS_ERR_CUDA, "swap launch (f32): %s", cudaGetErrorString(e));
    ++ctx->swap_call;
    count_swap_proposed(ctx, nbeta, npb);
    if (swapped_out) CU(cudaMemcpyAsync(swapped_out, ctx->d_swapped, m, cudaMemcpyDeviceToHost, ctx->stream));
    if (jvec || swapped_out) CU(cudaStreamSynchronize(ctx->stream));
    return SCORUS_OK;
}

int scorus::host::f32_swap_walkers(scorus_ctx *ctx, const double *beta, size_t nbeta)
{
    return f32_swap_impl(ctx, beta, nbeta, nullptr, nullptr, nullptr);
}

extern "C" int scorus_swap_walkers_fixed_f32(scorus_ctx *ctx, const double *beta, size_t nbeta, const uint32_t *jvec,
                                             const float *r, uint8_t *swapped_out)
{
    if (!ctx || !beta) return SCORUS_ERR_INVALID_ARG;
    if (!ctx->f32) return fail(ctx, SCORUS_ERR_INVALID_ARG, "scorus_swap_walkers_fixed_f32 needs a SCORUS_CFG_F32 context");
    if (nbeta > 1 && (!jvec || !r)) return SCORUS_ERR_INVALID_ARG;
    return f32_swap_impl(ctx, beta, nbeta, nbeta > 1 ? jvec : nullptr, r, swapped_out);
}
