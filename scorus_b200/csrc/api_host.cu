// api_host.cu -- C ABI, part 5: the literal drop-ins on caller-owned host buffers -- what a reference call site
// (`&mut [V]`, `&mut [T]` in, updated in place) maps to: upload, the move on the device, download.
#include "host_common.cuh"

// ------------------------------------------------------------------------------------------
// literal drop-ins on host buffers
// ------------------------------------------------------------------------------------------
// upload -> run(ctx) -> download, shared by the host-buffer entry points
template <class Run>
static int host_call(scorus_ctx *ctx, double *ens, double *lp, size_t lp_len, size_t nbeta, Run run)
{
    int rc = nbeta ? scorus_check_sample_args(ctx->n, lp_len, nbeta) : SCORUS_OK;  // nbeta == 0: checked by the callee
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
    CU(cudaMemsetAsync(ctx->d_dirty, 0, ctx->n, ctx->stream));
    ctx->track_dirty = true;
    rc = run();
    ctx->track_dirty = false;
    if (rc) return rc;
    if ((rc = launch_scatter_download(ctx, (double *)ae.devicePointer, (double *)al.devicePointer))) return rc;
    CU(cudaStreamSynchronize(ctx->stream));
    return SCORUS_OK;
}
// The stretch move on pinned buffers without an upload: the log-probs go up (8 bytes per walker), the FIRST step reads
// the rows straight out of the caller's ensemble over PCIe -- every row is read exactly once by a step anyway -- writes
// all of them to the device copy and every accepted row and log-prob straight back into the caller's arrays as it is
// decided (step_reg_kernel.cuh, kRemap variant with host_ens set).  Upload, step and write-back of the serial form
// (39 + 0.7 + 9 ms at 2^22 x 64-D) become one kernel in which both directions of the link run at once.  Further steps
// of the same call run on the device copy and their accepted walkers leave through scatter_download_kernel as before.
static int host_call_streamed(scorus_ctx *ctx, double *ens_dev, double *lp_host, double *lp_dev, double a,
                              const scorus_ufs *ufs, const double *beta, size_t nbeta, uint64_t nsteps)
{
    cudaSetDevice(ctx->device);
    ctx->swap_pending = false;  // everything a pending swap would have moved is replaced by the caller's state
    CU(cudaMemcpyAsync(ctx->d_lp, lp_host, ctx->n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    ctx->uploaded = true;
    CU(cudaMemsetAsync(ctx->d_dirty, 0, ctx->n, ctx->stream));
    ctx->host_ens = ens_dev;
    ctx->host_lp = lp_dev;
    int rc = scorus_sample_pt(ctx, a, ufs, beta, nbeta, 1);
    ctx->host_ens = ctx->host_lp = nullptr;
    if (rc) { ctx->uploaded = false; return rc; }  // (the device copy may be incomplete: the caller's buffers are the state)
    if (nsteps > 1) {
        ctx->track_dirty = true;
        rc = scorus_sample_pt(ctx, a, ufs, beta, nbeta, nsteps - 1);
        ctx->track_dirty = false;
        if (rc) return rc;
        if ((rc = launch_scatter_download(ctx, ens_dev, lp_dev))) return rc;
    }
    CU(cudaStreamSynchronize(ctx->stream));
    return SCORUS_OK;
}

extern "C" int scorus_sample_pt_host_n(scorus_ctx *ctx, double *ens, double *lp, size_t lp_len, double a,
                                       const scorus_ufs *ufs, const double *beta, size_t nbeta, uint64_t nsteps)
{
    if (!ctx || !ens || !lp || !ufs || !beta) return SCORUS_ERR_INVALID_ARG;
    F64_ONLY(ctx);
    // rows of 512 bytes and more: the step's own reads reach 48 GB/s over PCIe (the copy engine: 55) and the write-back
    // overlaps them -- 45 against 51 ms at 2^22 x 64-D; rows of 80 bytes travel at 11 GB/s that way and stay with the
    // flat upload (3.9 against 7.9 ms at c4)
    if (nsteps >= 1 && ctx->pitch == ctx->dim && (size_t)ctx->dim * sizeof(double) >= (size_t)env_int("SCORUS_HOST_STREAM_MIN_ROW", 512) &&
        env_int("SCORUS_HOST_STREAM", 1) && env_int("SCORUS_SCATTER_DOWNLOAD", 1) &&
        scorus_check_sample_args(ctx->n, lp_len, nbeta) == SCORUS_OK && a > 1.0 && ufs->kind != SCORUS_UFS_MASK &&
        !(ufs->kind == SCORUS_UFS_PROB && !(ufs->prob > 0.0)) && stretch_step_moves_rows(ctx, nbeta)) {
        cudaPointerAttributes ae{}, al{};
        const bool pinned = cudaPointerGetAttributes(&ae, ens) == cudaSuccess && ae.type == cudaMemoryTypeHost &&
                            cudaPointerGetAttributes(&al, lp) == cudaSuccess && al.type == cudaMemoryTypeHost;
        cudaGetLastError();
        if (pinned)
            return host_call_streamed(ctx, (double *)ae.devicePointer, lp, (double *)al.devicePointer, a, ufs, beta, nbeta, nsteps);
    }
    return host_call(ctx, ens, lp, lp_len, nbeta, [&] { return scorus_sample_pt(ctx, a, ufs, beta, nbeta, nsteps); });
}

// the same call for a walker shard of an ensemble spread over several GPUs, whole-ensemble matching, one-sided
// (scorus_sample_pt_peer): every rank calls it at the same time with its own shard's host buffers.  The owner of a
// pair marks an accepted partner on the partner's home GPU, so each rank still writes back only the walkers of ITS
// shard that were accepted, whoever computed them.
extern "C" int scorus_sample_pt_peer_host(scorus_ctx *ctx, double *ens, double *lp, size_t lp_len, double a,
                                          const scorus_ufs *ufs, const double *beta, size_t nbeta, uint64_t nsteps)
{
    if (!ctx || !ens || !lp) return SCORUS_ERR_INVALID_ARG;
    F64_ONLY(ctx);
    if (lp_len != ctx->n) return fail(ctx, SCORUS_ERR_LENGTH_MISMATCH, "ensemble and logprob lengths differ");
    // the rung arithmetic belongs to the whole ensemble (checked by scorus_sample_pt_peer): only the lengths here
    return host_call(ctx, ens, lp, lp_len, 0, [&] { return scorus_sample_pt_peer(ctx, a, ufs, beta, nbeta, nsteps); });
}

// twalk::sample on host buffers (the `ensemble_logprob: &mut (W, X)` argument, twalk.rs:588)
extern "C" int scorus_twalk_sample_host(scorus_ctx *ctx, double *ens, double *lp, size_t lp_len,
                                        const scorus_twalk_params *tw, const double *beta, size_t nbeta, uint64_t nsteps)
{
    if (!ctx || !ens || !lp || !tw) return SCORUS_ERR_INVALID_ARG;
    F64_ONLY(ctx);
    return host_call(ctx, ens, lp, lp_len, nbeta, [&] { return scorus_twalk_sample(ctx, tw, beta, nbeta, nsteps); });
}

extern "C" int scorus_sample_pt_host(scorus_ctx *ctx, double *ens, double *lp, size_t lp_len, double a,
                                     const scorus_ufs *ufs, const double *beta, size_t nbeta)
{
    F64_ONLY(ctx);
    return scorus_sample_pt_host_n(ctx, ens, lp, lp_len, a, ufs, beta, nbeta, 1);
}

extern "C" int scorus_swap_walkers_host(scorus_ctx *ctx, double *ens, double *lp, size_t lp_len,
                                        const double *beta, size_t nbeta)
{
    if (!ctx || !ens || !lp) return SCORUS_ERR_INVALID_ARG;
    F64_ONLY(ctx);
    if (nbeta == 0) return SCORUS_ERR_INVALID_ARG;
    if ((ctx->n / nbeta) * nbeta != ctx->n) return fail(ctx, SCORUS_ERR_NWALKERS_MISMATCHES_NBETA, "NWalkersMismatchesNBeta");
    if (lp_len != ctx->n) return SCORUS_OK;  // src/mcmc/utils.rs:88: silently does nothing
    int rc = scorus_check_beta_order(beta, nbeta);
    if (rc) return fail(ctx, rc, "beta list must be in decreasing order");
    if ((rc = scorus_upload(ctx, ens, lp))) return rc;
    if ((rc = scorus_swap_walkers(ctx, beta, nbeta))) return rc;
    return scorus_download(ctx, ens, lp);
}
