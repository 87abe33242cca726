// api_moves.cu -- C ABI, part 4: the other two move families behind the same context:
// the ensemble / parallel-tempering t-walk (src/mcmc/twalk.rs:586-762, twalk_kernel.cuh) and ensemble-PT HMC
// (src/mcmc/hmc/naive.rs:122-292, hmc_kernel.cuh).
#include "host_common.cuh"
#include "hmc_kernel.cuh"

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
    if ((rc = fill_common(ctx, &tp->s, g, nbeta, true))) return rc;
    if (g->kernel != KERNEL_REG)
        return fail(ctx, SCORUS_ERR_INVALID_ARG, "the t-walk needs the register kernel (dim <= 256, no SEQUENTIAL_SUM)");
    tp->aw = tw->aw;
    tp->aw_ratio = tw->aw / (1.0 + tw->aw);  // twalk.rs:292
    tp->at = tw->at;
    tp->fw0 = tw->fw[0];
    tp->fw1 = tw->fw[1];
    tp->kernel_in = nullptr; tp->b_in = nullptr; tp->walk_u_in = nullptr;
    tp->l2_keep = (uint32_t)env_int("SCORUS_TWALK_L2", 1);
    // twalk_direct_kernel.cuh: no shared-memory tile (flag words + a short z list per walker), 16 warps per SM.
    // Measured: 2.28 -> 2.00 ms per step at 64-D (the tile held the SM at 11 warps); at 10-D, where the tile is small
    // and a one-block tile cannot prefetch its own re-read, 0.201 -> 0.209 ms: rows of up to 16 double2 keep the tile.
    // SCORUS_TWALK_DIRECT=0 / 2 forces the tile / the direct kernel wherever it exists.
    const uint32_t nvec = ctx->pitch / 2;
    const int want_direct = env_int("SCORUS_TWALK_DIRECT", 1);
    if (want_direct && ctx->lp_kind != SCORUS_LP_CORR_GAUSSIAN && (nvec > 16 || (want_direct == 2 && !(nvec > 8 && nvec <= 16)))) {
        tp->direct = 1;
        const size_t per_warp = (((size_t)g->flag_words * 128u + 127u) & ~(size_t)127u) + 32u * 8u * sizeof(double);
        const uint32_t ntiles = (uint32_t)((ctx->n + 31u) / 32u);
        g->warps = 16;
        g->smem = g->warps * per_warp;
        g->grid = std::min<uint32_t>((uint32_t)ctx->num_sms, (ntiles + g->warps - 1) / g->warps);
        tp->s.warps = g->warps;
    }
    StepParams &p = tp->s;
    p.flag_kind = SCORUS_UFS_PROB;
    return set_prob_flags(ctx, &p, tw->pphi);
}

// twalk::sample (:586-648): one matching, two half-passes (sample1 :650-762) with the roles exchanged
extern "C" int scorus_twalk_sample(scorus_ctx *ctx, const scorus_twalk_params *tw, const double *beta, size_t nbeta,
                                   uint64_t nsteps)
{
    if (!ctx || !tw || !beta) return SCORUS_ERR_INVALID_ARG;
    F64_ONLY(ctx);
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
            launch_exact_order(ctx, ctx->step, nbeta, p.npb, ctx->d_order, (uint32_t)ctx->w_off);
            p.order = ctx->d_order;
        }
        if (fused) {  // both half-passes of every pair in one launch (twalk_kernel.cuh); counters step, step + 1
            tp.half = 2;
            p.step = ctx->step;
            stamp_tiles(ctx, &p, g);
            cudaError_t e = launch_twalk(ctx->lp_kind, tp, g, ctx->stream);
            if (e != cudaSuccess) return fail(ctx, SCORUS_ERR_CUDA, "t-walk launch: %s", cudaGetErrorString(e));
            commit_tiles(ctx, p);
            ++ctx->launches;
            ctx->step += 2;
        } else {
            for (uint32_t half = 0; half < 2; ++half) {
                tp.half = half;
                p.step = ctx->step;
                stamp_tiles(ctx, &p, g);
                cudaError_t e = launch_twalk(ctx->lp_kind, tp, g, ctx->stream);
                if (e != cudaSuccess) return fail(ctx, SCORUS_ERR_CUDA, "t-walk launch: %s", cudaGetErrorString(e));
            commit_tiles(ctx, p);
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
    F64_ONLY(ctx);
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
            commit_tiles(ctx, p);
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
    F64_ONLY(ctx);
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
    F64_ONLY(ctx);
    return hmc_impl(ctx, epsilon, beta, nbeta, l, param, nullptr, nullptr, nullptr, nullptr, accepted_cnt, nsteps);
}

extern "C" int scorus_hmc_sample_ensemble_pt_fixed(scorus_ctx *ctx, double *epsilon, const double *beta, size_t nbeta, size_t l,
                                                   const scorus_hmc_param *param, const double *p0, const double *u1,
                                                   const double *u2, uint8_t *accepted_out, uint64_t *accepted_cnt)
{
    if (!ctx || !epsilon || !beta || !param || !p0 || !u1 || !u2) return SCORUS_ERR_INVALID_ARG;
    F64_ONLY(ctx);
    return hmc_impl(ctx, epsilon, beta, nbeta, l, param, p0, u1, u2, accepted_out, accepted_cnt, 1);
}

// Net step-size adaptation of the LAST HMC step over this context's walkers: #grow - #shrink of the codes the kernel
// wrote (naive.rs:274-283).  For walker shards of one rung: every rank runs the step with HmcParam::fixed, the nets are
// summed over the ranks and epsilon (1 + adj)^net is applied once -- what the single-GPU path does for rungs above 4096
// walkers (hmc_adapt_count_kernel), so the sharded epsilon is the single-GPU epsilon bit for bit.
static __global__ void __launch_bounds__(256) hmc_net_kernel(const int8_t *adapt, uint32_t n, long long *net)
{
    long long v = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) v += adapt[i];
    for (int off = 16; off; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if ((threadIdx.x & 31u) == 0 && v) atomicAdd(reinterpret_cast<unsigned long long *>(net), (unsigned long long)v);
}

extern "C" int scorus_hmc_adapt_net(scorus_ctx *ctx, int64_t *net_out)
{
    if (!ctx || !net_out) return SCORUS_ERR_INVALID_ARG;
    F64_ONLY(ctx);
    if (!ctx->d_hmc_adapt) return fail(ctx, SCORUS_ERR_INVALID_ARG, "no HMC step has run on this context");
    cudaSetDevice(ctx->device);
    int rc = ensure(ctx, (void **)&ctx->d_hmc_eps, &ctx->hmc_eps_cap, sizeof(double));
    if (rc) return rc;
    long long *d_net = reinterpret_cast<long long *>(ctx->d_hmc_eps);  // scratch: epsilon went back to the host already
    CU(cudaMemsetAsync(d_net, 0, sizeof(long long), ctx->stream));
    hmc_net_kernel<<<(unsigned)std::min<size_t>((ctx->n + 255) / 256, (size_t)ctx->num_sms * 8u), 256, 0, ctx->stream>>>(
        ctx->d_hmc_adapt, (uint32_t)ctx->n, d_net);
    ++ctx->launches;
    long long h = 0;
    CU(cudaMemcpyAsync(&h, d_net, sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    *net_out = (int64_t)h;
    return SCORUS_OK;
}

// sample_ensemble_pt on caller-owned host buffers (q0: &mut [V], lp: &mut [T], :125-126)
extern "C" int scorus_hmc_sample_ensemble_pt_host(scorus_ctx *ctx, double *q0, double *lp, size_t lp_len, double *epsilon,
                                                  const double *beta, size_t nbeta, size_t l, const scorus_hmc_param *param,
                                                  uint64_t *accepted_cnt)
{
    if (!ctx || !q0 || !lp || !epsilon || !beta || !param) return SCORUS_ERR_INVALID_ARG;
    F64_ONLY(ctx);
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
    F64_ONLY(ctx);
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
