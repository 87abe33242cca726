// api_stats.cu -- C ABI, part 3: the periphery of the path (stats_kernel.cuh) -- box initialisation
// (src/mcmc/init_ensemble.rs:19-32), mean / cov (src/linear_space/utils.rs:4-41), the chain trace, and the
// chain statistics of SURVEY 8(f) rank 1 (per-rung counts, running moments, autocorrelation time).
#include "host_common.cuh"
#include "stats_kernel.cuh"

static int accumulate_moments(scorus_ctx *ctx);

// after every step: chain capture (scorus_trace_enable; silently stops when the buffer is full) and the
// running moments (scorus_moments_enable)
int scorus::host::after_step(scorus_ctx *ctx)
{
    if (ctx->swap_pending) {  // (a step leaves none; kept so that a new caller cannot read rows that have not moved)
        int rc = settle_swap(ctx);
        if (rc) return rc;
    }
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

// the download half of the host-buffer calls (api_host.cu): accepted walkers only, written in place
int scorus::host::launch_scatter_download(scorus_ctx *ctx, double *host_ens, double *host_lp)
{
    int rc = settle_swap(ctx);
    if (rc) return rc;
    scatter_download_kernel<<<(unsigned)ctx->num_sms * 8u, 256, 0, ctx->stream>>>(
        ctx->d_ens, ctx->d_lp, ctx->d_dirty, (uint32_t)ctx->n, ctx->dim, ctx->pitch, host_ens, host_lp);
    ++ctx->launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(ctx, SCORUS_ERR_CUDA, "scatter download launch: %s", cudaGetErrorString(e));
    return SCORUS_OK;
}

// ------------------------------------------------------------------------------------------
// periphery: box init, ensemble statistics, chain trace (stats_kernel.cuh)
// ------------------------------------------------------------------------------------------
extern "C" int scorus_init_ensemble(scorus_ctx *ctx, const double *lo, const double *hi)
{
    if (!ctx || !lo || !hi) return SCORUS_ERR_INVALID_ARG;
    F64_ONLY(ctx);
    for (uint32_t d = 0; d < ctx->dim; ++d)
        if (!(lo[d] < hi[d])) return fail(ctx, SCORUS_ERR_INVALID_ARG, "empty range in coordinate %u (Uniform::new panics)", d);
    cudaSetDevice(ctx->device);
    ctx->swap_pending = false;  // every row is overwritten below
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
    int rc = settle_swap(ctx);
    if (rc) return rc;
    uint32_t blocks = (uint32_t)ctx->num_sms * 2u;
    if (blocks > count) blocks = (uint32_t)count;
    rc = ensure(ctx, (void **)&ctx->d_part, &ctx->part_cap, (size_t)blocks * (cov_out ? len : dim) * sizeof(double));
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
    F64_ONLY(ctx);
    return stats_impl(ctx, first, count, mean_out, nullptr);
}

extern "C" int scorus_cov(scorus_ctx *ctx, size_t first, size_t count, double *cov_out)
{
    if (!ctx || !cov_out) return SCORUS_ERR_INVALID_ARG;
    F64_ONLY(ctx);
    return stats_impl(ctx, first, count, nullptr, cov_out);
}

extern "C" int scorus_trace_enable(scorus_ctx *ctx, const uint32_t *walkers, const uint32_t *coords, size_t ncapture,
                                   size_t capacity_steps)
{
    if (!ctx) return SCORUS_ERR_INVALID_ARG;
    F64_ONLY(ctx);
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
    F64_ONLY(ctx);
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
    F64_ONLY(ctx);
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
    F64_ONLY(ctx);
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
    F64_ONLY(ctx);
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
    F64_ONLY(ctx);
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
