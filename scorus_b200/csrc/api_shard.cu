// api_shard.cu -- C ABI, part 2: one process per GPU.  Shard bookkeeping, the
// whole-ensemble matching over gathered copies, the redundant swap plan, and the one-sided NVLink paths
// (CUDA IPC mappings of the peers' buffers, partner pulls, swap row pulls).
#include "host_common.cuh"

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
        launch_exact_order(ctx, ctx->step, nbeta, p.npb, ctx->d_gorder, 0u);
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
        launch_exact_jvec(ctx, nbeta, npb, ctx->d_jinv);
        sp.jinv = ctx->d_jinv;
    }
    launch_swap_chain(ctx, sp);
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
        launch_exact_order(ctx, ctx->step, nbeta, p.npb, ctx->d_gorder, 0u);
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
