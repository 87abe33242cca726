// api_shard.cu -- C ABI, part 2: one process per GPU.  Shard bookkeeping, the
// whole-ensemble matching over gathered copies, the redundant swap plan, and the one-sided NVLink paths
// (CUDA IPC mappings of the peers' buffers, partner pulls, swap row pulls).
#include "host_common.cuh"
#include "step_reg_kernel.cuh"  // ld_stream

// ------------------------------------------------------------------------------------------
// sharding: one process per GPU
// ------------------------------------------------------------------------------------------
extern "C" int scorus_set_shard(scorus_ctx *ctx, uint64_t walker_offset, uint32_t rung_offset, uint64_t n_walkers_global)
{
    if (!ctx) return SCORUS_ERR_INVALID_ARG;
    F64_ONLY(ctx);
    if (walker_offset + ctx->n > n_walkers_global || n_walkers_global > 0xFFFFFFFFull)
        return fail(ctx, SCORUS_ERR_INVALID_ARG, "shard [%llu, %llu) does not fit in %llu walkers",
                    (unsigned long long)walker_offset, (unsigned long long)(walker_offset + ctx->n),
                    (unsigned long long)n_walkers_global);
    ctx->w_off = walker_offset;
    ctx->rung_off = rung_offset;
    ctx->n_global = n_walkers_global;
    return SCORUS_OK;
}

// ------------------------------------------------------------------------------------------
// pair lists of a shard under the whole-ensemble matching
// ------------------------------------------------------------------------------------------
// PAIRS_EITHER: every pair with a member in [w_lo, w_hi) (all-gather mode: both ranks of a cross pair compute it,
// each keeps its own walker).  PAIRS_OWNER: every pair whose EVEN-position member is in [w_lo, w_hi) -- each pair on
// exactly one rank, which updates both walkers (one-sided mode).  The positions are a keyed bijection of the walkers,
// so ownership is balanced to O(sqrt(n)) without any exchange.
enum : int { PAIRS_EITHER = 0, PAIRS_OWNER = 1 };

// stream-ordered barrier across the ranks of scorus_ipc_attach, by flags in peer memory: rank r stores `epoch` into
// slot r of every peer's flag row (the 32 doubles behind its log-probs) and waits until all slots of its own row
// have reached it.  Threads 0..world-1 of ONE block.  Everything this GPU wrote before (kernel boundary +
// fence.sys) is visible to a peer that has seen the flag.  A peer that never arrives traps after 60 s instead of
// hanging the GPU.
struct PeerTable { double *lp[16]; };
__device__ __forceinline__ void peer_barrier_device(const PeerTable &t, uint32_t n_local, int world, int rank,
                                                    unsigned long long epoch)
{
    const int q = (int)threadIdx.x;
    if (q >= world) return;
    unsigned long long *theirs = reinterpret_cast<unsigned long long *>(t.lp[q] + n_local) + rank;
    const unsigned long long *mine = reinterpret_cast<const unsigned long long *>(t.lp[rank] + n_local) + q;
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(theirs), "l"(epoch) : "memory");
    unsigned long long v, t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine) : "memory");
        if (v >= epoch) break;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > 60000000000ull) __trap();
    }
    __threadfence_system();
}

__global__ void __launch_bounds__(32) peer_barrier_kernel(const PeerTable t, uint32_t n_local, int world, int rank,
                                                          unsigned long long epoch)
{
    peer_barrier_device(t, n_local, world, rank, epoch);
}

// large rungs: each local walker finds its matching position with the inverse bijection (rng.cuh) -- O(n_local) per
// rank, no pass over the whole matching.  Block 0 first runs the inter-rank barrier when epoch != 0 (the pair list
// depends on the step index only, so building it overlaps the wait); thread 0 also resets the tile counter of the
// step kernel that follows.
__global__ void __launch_bounds__(256) shard_pairs_kernel(const StepParams p, int mode, uint32_t *pairs, uint32_t *count,
                                                          unsigned long long *tile_counter, const PeerTable t, int world,
                                                          int rank, unsigned long long epoch)
{
    if (blockIdx.x == 0) {
        if (threadIdx.x == 0 && tile_counter) *tile_counter = 0ull;
        if (epoch) peer_barrier_device(t, p.w_hi - p.w_lo, world, rank, epoch);
    }
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31u;
    bool emit = false;
    uint32_t a = 0, b = 0;
    if (i < p.w_hi - p.w_lo) {
        const uint32_t w = p.w_lo + i, r = w / p.npb, k = w - r * p.npb;
        const BijKeys K = bij_keys(p.seed, p.step, ST_PERM, r + p.rung_off);
        const uint32_t pos = bij_inv(k, p.npb, p.bij_bits, K);
        const uint32_t mate = r * p.npb + bij(pos ^ 1u, p.npb, p.bij_bits, K);
        const bool even = (pos & 1u) == 0u;
        const bool mate_local = mate >= p.w_lo && mate < p.w_hi;
        emit = mode == PAIRS_OWNER ? even : (even || !mate_local);
        a = even ? w : mate;  // list order = matching order: even position first
        b = even ? mate : w;
    }
    // one atomic per BLOCK on the list length (one per warp was 131072 same-address atomics at 2^22 walkers: most of
    // the kernel's 111 us, profiles/r2_c5_peer_single_launch_list_summary.csv)
    __shared__ uint32_t warp_cnt[8], block_base;
    const unsigned m = __ballot_sync(0xffffffffu, emit);
    if (lane == 0) warp_cnt[threadIdx.x >> 5] = (uint32_t)__popc(m);
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t tot = 0;
        for (int wq = 0; wq < 8; ++wq) { const uint32_t c = warp_cnt[wq]; warp_cnt[wq] = tot; tot += c; }
        block_base = tot ? atomicAdd(count, 2u * tot) : 0u;
    }
    __syncthreads();
    if (emit) {
        const uint32_t slot = block_base + 2u * (warp_cnt[threadIdx.x >> 5] + (uint32_t)__popc(m & ((1u << lane) - 1u)));
        pairs[slot] = a;
        pairs[slot + 1] = b;
    }
}

// small rungs (exact shuffle: the order of the whole ensemble is in p.order): scan the matching
__global__ void local_pairs_kernel(const StepParams p, uint32_t n_global, int mode, uint32_t *pairs, uint32_t *count)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (2u * k + 1u >= n_global) return;
    const uint32_t a = walker_at(p, 2u * k), b = walker_at(p, 2u * k + 1u);
    const bool ina = a >= p.w_lo && a < p.w_hi, inb = b >= p.w_lo && b < p.w_hi;
    if (mode == PAIRS_OWNER ? ina : (ina || inb)) {
        const uint32_t slot = atomicAdd(count, 2u);
        pairs[slot] = a;
        pairs[slot + 1] = b;
    }
}

// common front end of the two whole-ensemble-matching entry points: validates, fills the launch parameters with
// GLOBAL ids and builds this rank's pair list for step ctx->step.  epoch != 0: the list kernel starts with the
// inter-rank barrier.
static int global_matching_setup(scorus_ctx *ctx, double a, const scorus_ufs *ufs, const double *beta, size_t nbeta,
                                 int mode, unsigned long long epoch, StepParams *pp, Geometry *gp, bool *all_flags)
{
    const size_t ng = ctx->n_global;
    int rc = scorus_check_sample_args(ng, ng, nbeta);
    if (rc) return fail(ctx, rc, "%s", scorus_status_string(rc));
    if (ctx->lp_kind < 0) return fail(ctx, SCORUS_ERR_NO_LOGPROB, "call scorus_set_logprob first");
    if (!ctx->uploaded) return fail(ctx, SCORUS_ERR_NOT_UPLOADED, "call scorus_upload first");
    if (!(a > 1.0)) return fail(ctx, SCORUS_ERR_INVALID_ARG, "stretch scale a must be > 1");
    // UpdateFlagSpec::Func (ensemble_sample.rs:31,52,150-153): the closure is called once per walker of the WHOLE
    // ensemble in order, so over shards the mask covers all n_global walkers (every rank evaluates the same closure
    // n_global times, as the single-GPU run does) and is indexed by global walker id
    if (ufs->kind == SCORUS_UFS_MASK) {
        if (!ufs->mask || ufs->mask_len > ctx->dim) return fail(ctx, SCORUS_ERR_INVALID_ARG, "bad mask");
        for (size_t i = 0; i < ng; ++i) {
            size_t c = 0;
            for (size_t k = 0; k < ufs->mask_len; ++k) c += ufs->mask[i * ufs->mask_len + k] ? 1 : 0;
            if (c == 0) return fail(ctx, SCORUS_ERR_NPHI_ZERO, "walker %zu has no update flag set", i);
        }
    }
    if (ctx->sequential || ctx->pitch > 256) return fail(ctx, SCORUS_ERR_INVALID_ARG, "whole-ensemble matching over shards needs the register kernel (dim <= 256, no SEQUENTIAL_SUM)");
    cudaSetDevice(ctx->device);
    StepParams &p = *pp;
    Geometry &g = *gp;
    if ((rc = upload_beta(ctx, beta, nbeta))) return rc;
    if ((rc = fill_common(ctx, &p, &g, nbeta))) return rc;
    if (g.kernel != KERNEL_REG) return fail(ctx, SCORUS_ERR_INVALID_ARG, "whole-ensemble matching over shards needs the register kernel");
    const double sqrt_a = sqrt(a);
    p.inv_sqrt_a = 1.0 / sqrt_a;
    p.z_range = 2.0 * (sqrt_a - p.inv_sqrt_a);
    p.flag_kind = (uint32_t)ufs->kind;
    *all_flags = ufs->kind == SCORUS_UFS_ALL;
    if (ufs->kind == SCORUS_UFS_PROB && (rc = set_prob_flags(ctx, &p, ufs->prob))) return rc;
    if (ufs->kind == SCORUS_UFS_MASK) {
        const size_t bytes = ng * ufs->mask_len;
        if ((rc = ensure(ctx, (void **)&ctx->d_flags, &ctx->flags_cap, bytes ? bytes : 1))) return rc;
        CU(cudaMemcpyAsync(ctx->d_flags, ufs->mask, bytes, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));  // the caller's mask is released on return
        p.flags_in = ctx->d_flags;
        p.flag_len = (uint32_t)ufs->mask_len;
    }
    // ids are global in this mode: no offsets on the stream keys, whole-rung matching
    p.npb = (uint32_t)(ng / nbeta);
    p.bij_bits = bij_bits(p.npb);
    p.mblk = p.npb; p.mblk_bits = p.bij_bits; p.mblk_off = 0;
    p.w_off = 0; p.rung_off = 0;
    p.w_lo = (uint32_t)ctx->w_off; p.w_hi = (uint32_t)(ctx->w_off + ctx->n);
    p.step = ctx->step; p.onehot = ctx->onehot;
    p.order = nullptr;
    if (mode != PAIRS_OWNER) p.dirty = nullptr;  // one-sided mode: the marks live behind every rank's log-probs
    if ((rc = ensure(ctx, (void **)&ctx->d_pairs, &ctx->pairs_cap, 2 * ctx->n * sizeof(uint32_t)))) return rc;
    if (!ctx->d_count) CU(cudaMalloc(&ctx->d_count, sizeof(uint32_t) + 2 * sizeof(unsigned long long)));
    unsigned long long *tile_counter = reinterpret_cast<unsigned long long *>(ctx->d_count) + 1;
    CU(cudaMemsetAsync(ctx->d_count, 0, sizeof(uint32_t), ctx->stream));
    PeerTable t;
    for (int q = 0; q < 16; ++q) t.lp[q] = ctx->peer_lp[q];
    if (p.npb <= kExactShuffleMax) {
        // exact shuffle of every rung of the whole ensemble (small rungs), then a scan of the matching
        if ((rc = ensure(ctx, (void **)&ctx->d_gorder, &ctx->gorder_cap, ng * sizeof(uint32_t)))) return rc;
        launch_exact_order(ctx, ctx->step, nbeta, p.npb, ctx->d_gorder, 0u);
        p.order = ctx->d_gorder;
        if (epoch) peer_barrier_kernel<<<1, 32, 0, ctx->stream>>>(t, (uint32_t)ctx->n, ctx->world, ctx->rank, epoch);
        CU(cudaMemsetAsync(tile_counter, 0, sizeof(unsigned long long), ctx->stream));
        local_pairs_kernel<<<(unsigned)((ng / 2 + 255) / 256), 256, 0, ctx->stream>>>(p, (uint32_t)ng, mode, ctx->d_pairs, ctx->d_count);
        ctx->launches += epoch ? 2 : 1;
    } else {
        shard_pairs_kernel<<<(unsigned)((ctx->n + 255) / 256), 256, 0, ctx->stream>>>(p, mode, ctx->d_pairs, ctx->d_count, tile_counter,
                                                                                   t, ctx->world, ctx->rank, epoch);
        ++ctx->launches;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(ctx, SCORUS_ERR_CUDA, "pair list launch: %s", cudaGetErrorString(e));
    // the step itself: pair list as the matching order, tiles claimed dynamically from the counter the list kernel reset
    p.order = ctx->d_pairs;
    p.n_dev = ctx->d_count;
    p.n = (uint32_t)(2 * ctx->n);
    p.ntiles = (p.n + 31u) / 32u;
    p.tile_counter = env_int("SCORUS_DYNAMIC_TILES", 1) ? tile_counter : nullptr;
    p.tile_base = 0;
    return SCORUS_OK;
}

extern "C" int scorus_sample_pt_global(scorus_ctx *ctx, double a, const scorus_ufs *ufs, const double *beta,
                                       size_t nbeta, const double *ens_all, const double *lp_all)
{
    if (!ctx || !ufs || !beta || !ens_all || !lp_all) return SCORUS_ERR_INVALID_ARG;
    F64_ONLY(ctx);
    StepParams p;
    Geometry g;
    bool all_flags = false;
    int rc = global_matching_setup(ctx, a, ufs, beta, nbeta, PAIRS_EITHER, 0ull, &p, &g, &all_flags);
    if (rc) return rc;
    p.src_ens = ens_all;  // rows from the gathered copies
    p.src_lp = lp_all;
    cudaError_t e = launch_step(ctx->lp_kind, p, g, all_flags, ctx->stream);
    if (e != cudaSuccess) return fail(ctx, SCORUS_ERR_CUDA, "step launch: %s", cudaGetErrorString(e));
    ++ctx->launches;
    ++ctx->step;
    count_proposed(ctx, nbeta, p.npb, p.w_lo, p.w_hi, 1);
    if (ufs->kind == SCORUS_UFS_ONE_HOT_CYCLE) ctx->onehot = (ctx->onehot + ctx->n_global) % ctx->dim;
    return after_step(ctx);
}

extern "C" int scorus_swap_plan_device(scorus_ctx *ctx, const double *beta, size_t nbeta, double *lp_all,
                                       uint32_t *src_all)
{
    if (!ctx || !beta || !lp_all || !src_all) return SCORUS_ERR_INVALID_ARG;
    F64_ONLY(ctx);
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
    sp.lp = lp_all; sp.lp_src = lp_all; sp.beta = ctx->d_beta; sp.src = src_all; sp.stats = ctx->d_stats;
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
    if ((rc = launch_swap_chain(ctx, sp))) return rc;
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
    F64_ONLY(ctx);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaSetDevice(ctx->device);
    int rc = settle_swap(ctx);  // the handles are of the buffers that hold the ensemble from here on
    if (rc) return rc;
    CU(cudaStreamSynchronize(ctx->stream));  // the barrier flags behind the log-probs are zeroed before a peer can write them
    CU(cudaIpcGetMemHandle((cudaIpcMemHandle_t *)handle_ens, ctx->d_ens));
    CU(cudaIpcGetMemHandle((cudaIpcMemHandle_t *)handle_lp, ctx->d_lp));
    return SCORUS_OK;
}

extern "C" int scorus_ipc_attach(scorus_ctx *ctx, int world, int rank, const void *handles_ens, const void *handles_lp)
{
    if (!ctx || world < 1 || world > 16 || rank < 0 || rank >= world || !handles_ens || !handles_lp)
        return SCORUS_ERR_INVALID_ARG;
    F64_ONLY(ctx);
    // the barrier epochs live in the peers' memory: a second attach would restart them below what the flags already hold
    if (ctx->d_peer_ens) return fail(ctx, SCORUS_ERR_INVALID_ARG, "this context is already attached to its peers");
    cudaSetDevice(ctx->device);
    {
        int rc = settle_swap(ctx);
        if (rc) return rc;
    }
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
    ctx->bar_epoch = 0;
    return SCORUS_OK;
}

// Measured NVLink read bandwidth of this GPU from one peer, the denominator of the one-sided step's roofline: (a) a
// copy-engine cudaMemcpyAsync from the peer's mapped ensemble, (b) an SM kernel streaming 16-byte loads from it, the
// access pattern of the step kernel.  `bytes` of the peer's ensemble buffer are read `iters` times; results in GB/s.
__global__ void __launch_bounds__(512) peer_read_kernel(const double2 *src, size_t nvec, double *sink)
{
    double acc = 0.0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < nvec; i += 4 * stride) {  // four independent 16-byte loads in flight per thread
        const double2 a = ld_stream(src + i), b = ld_stream(src + i + stride), c = ld_stream(src + i + 2 * stride),
                      d = ld_stream(src + i + 3 * stride);
        acc += a.x + a.y + b.x + b.y + c.x + c.y + d.x + d.y;
    }
    for (; i < nvec; i += stride) { const double2 a = ld_stream(src + i); acc += a.x + a.y; }
    if (acc == 1.2345e300) *sink = acc;  // keeps the loads alive
}

// the step kernel's pattern: whole rows (one warp-wide 16-byte-per-lane load per 512 bytes) at pseudo-random row indices
__global__ void __launch_bounds__(512) peer_read_rows_kernel(const double2 *src, uint32_t nrows, uint32_t vec_per_row, double *sink)
{
    double acc = 0.0;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t b = bij_bits(nrows);
    BijKeys K;
    for (int q = 0; q < 8; ++q) K.k[q] = 0x9E3779B9u * (uint32_t)(q + 1) + 12345u;
    for (uint32_t i = warp; i + 3 * nwarps < nrows; i += 4 * nwarps) {  // four rows in flight per warp
        const double2 *r0 = src + (size_t)bij(i, nrows, b, K) * vec_per_row, *r1 = src + (size_t)bij(i + nwarps, nrows, b, K) * vec_per_row,
                      *r2 = src + (size_t)bij(i + 2 * nwarps, nrows, b, K) * vec_per_row, *r3 = src + (size_t)bij(i + 3 * nwarps, nrows, b, K) * vec_per_row;
        for (uint32_t c = lane; c < vec_per_row; c += 32) {
            const double2 a = ld_stream(r0 + c), e = ld_stream(r1 + c), f = ld_stream(r2 + c), g = ld_stream(r3 + c);
            acc += a.x + a.y + e.x + e.y + f.x + f.y + g.x + g.y;
        }
    }
    if (acc == 1.2345e300) *sink = acc;
}

extern "C" int scorus_peer_read_bandwidth(scorus_ctx *ctx, int peer, size_t bytes, int iters, double *gbps_copy_engine,
                                          double *gbps_sm_loads, double *gbps_random_rows)
{
    if (!ctx || iters < 1) return SCORUS_ERR_INVALID_ARG;
    F64_ONLY(ctx);
    if (!ctx->d_peer_ens || peer < 0 || peer >= ctx->world) return fail(ctx, SCORUS_ERR_INVALID_ARG, "call scorus_ipc_attach first");
    cudaSetDevice(ctx->device);
    const size_t cap = ctx->n * (size_t)ctx->pitch * sizeof(double);
    if (bytes == 0 || bytes > cap) bytes = cap;
    bytes &= ~(size_t)15;
    if (!ctx->d_scratch) CU(cudaMalloc(&ctx->d_scratch, cap));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    float ms = 0.f;
    CU(cudaMemcpyAsync(ctx->d_scratch, ctx->peer_ens[peer], bytes, cudaMemcpyDeviceToDevice, ctx->stream));  // warm-up
    CU(cudaEventRecord(e0, ctx->stream));
    for (int k = 0; k < iters; ++k)
        CU(cudaMemcpyAsync(ctx->d_scratch, ctx->peer_ens[peer], bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    CU(cudaEventRecord(e1, ctx->stream));
    CU(cudaEventSynchronize(e1));
    CU(cudaEventElapsedTime(&ms, e0, e1));
    if (gbps_copy_engine) *gbps_copy_engine = (double)bytes * iters / (ms * 1e-3) / 1e9;
    const unsigned grid = (unsigned)ctx->num_sms * 4u;
    peer_read_kernel<<<grid, 512, 0, ctx->stream>>>((const double2 *)ctx->peer_ens[peer], bytes / 16, ctx->d_scratch);
    CU(cudaEventRecord(e0, ctx->stream));
    for (int k = 0; k < iters; ++k)
        peer_read_kernel<<<grid, 512, 0, ctx->stream>>>((const double2 *)ctx->peer_ens[peer], bytes / 16, ctx->d_scratch);
    CU(cudaEventRecord(e1, ctx->stream));
    CU(cudaEventSynchronize(e1));
    CU(cudaEventElapsedTime(&ms, e0, e1));
    if (gbps_sm_loads) *gbps_sm_loads = (double)bytes * iters / (ms * 1e-3) / 1e9;
    ctx->launches += (uint64_t)iters + 1;
    if (gbps_random_rows) {  // the rows of the peer's ensemble in a scrambled order, each read once per pass
        const uint32_t vpr = ctx->pitch / 2, nrows = (uint32_t)((bytes / 16) / vpr) & ~3u;
        peer_read_rows_kernel<<<grid, 512, 0, ctx->stream>>>((const double2 *)ctx->peer_ens[peer], nrows, vpr, ctx->d_scratch);
        CU(cudaEventRecord(e0, ctx->stream));
        for (int k = 0; k < iters; ++k)
            peer_read_rows_kernel<<<grid, 512, 0, ctx->stream>>>((const double2 *)ctx->peer_ens[peer], nrows, vpr, ctx->d_scratch);
        CU(cudaEventRecord(e1, ctx->stream));
        CU(cudaEventSynchronize(e1));
        CU(cudaEventElapsedTime(&ms, e0, e1));
        *gbps_random_rows = (double)nrows * vpr * 16.0 * iters / (ms * 1e-3) / 1e9;
        ctx->launches += (uint64_t)iters + 1;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return SCORUS_OK;
}

// Stream-ordered barrier across the attached ranks (flags in peer memory, no host synchronisation, no NCCL).
// Every rank must call it the same number of times.
extern "C" int scorus_peer_barrier(scorus_ctx *ctx)
{
    if (!ctx) return SCORUS_ERR_INVALID_ARG;
    F64_ONLY(ctx);
    if (!ctx->d_peer_ens || ctx->world < 1) return fail(ctx, SCORUS_ERR_INVALID_ARG, "call scorus_ipc_attach first");
    cudaSetDevice(ctx->device);
    PeerTable t;
    for (int q = 0; q < 16; ++q) t.lp[q] = ctx->peer_lp[q];
    peer_barrier_kernel<<<1, 32, 0, ctx->stream>>>(t, (uint32_t)ctx->n, ctx->world, ctx->rank, ++ctx->bar_epoch);
    ++ctx->launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(ctx, SCORUS_ERR_CUDA, "barrier launch: %s", cudaGetErrorString(e));
    return SCORUS_OK;
}

// sample_pt of a walker shard under the reference's whole-ensemble matching (ensemble_sample.rs:135-148), one-sided:
// every pair of the matching is OWNED by the rank that holds its even-position member; the owner reads both rows
// where they live (a remote partner's row straight from the peer's HBM over NVLink, no staging copy), proposes and
// decides for both walkers, and writes accepted rows and log-probs back to their home ranks.  A walker is in exactly
// one pair and a pair on exactly one rank, so nothing is computed twice, half of the cross-rank rows travel, and
// inside a step no rank touches a row another rank uses: the only synchronisation is ONE barrier before each step
// (fused into the pair-list kernel) and one after the last, so that whatever follows sees every peer's writes.
// Results are bit-identical to the single-GPU step (same streams, keyed by global walker ids).
extern "C" int scorus_sample_pt_peer(scorus_ctx *ctx, double a, const scorus_ufs *ufs, const double *beta, size_t nbeta,
                                     uint64_t nsteps)
{
    if (!ctx || !ufs || !beta) return SCORUS_ERR_INVALID_ARG;
    F64_ONLY(ctx);
    if (!ctx->d_peer_ens || ctx->world < 1) return fail(ctx, SCORUS_ERR_INVALID_ARG, "call scorus_ipc_attach first");
    if (ctx->n_global != ctx->n * (size_t)ctx->world) return fail(ctx, SCORUS_ERR_INVALID_ARG, "shards must be equal");
    if (ctx->n > (1u << 28)) return fail(ctx, SCORUS_ERR_INVALID_ARG, "at most 2^28 walkers per shard");
    if (ufs->kind == SCORUS_UFS_MASK && nsteps != 1) return fail(ctx, SCORUS_ERR_INVALID_ARG, "a host-evaluated mask covers exactly one step");
    for (uint64_t s = 0; s < nsteps; ++s) {
        StepParams p;
        Geometry g;
        bool all_flags = false;
        int rc = global_matching_setup(ctx, a, ufs, beta, nbeta, PAIRS_OWNER, ++ctx->bar_epoch, &p, &g, &all_flags);
        if (rc) return rc;
        for (int q = 0; q < 16; ++q) { p.peer_ens[q] = ctx->peer_ens[q]; p.peer_lp[q] = ctx->peer_lp[q]; }
        p.peer_n = (uint32_t)ctx->n;
        p.accepted_out = nullptr;
        // partner rows a tile ahead through a shared-memory stage (step_reg_kernel.cuh, kPeer == 2) where the kernel
        // keeps its proposals in registers: 16 rows + 8 mbarriers per warp on top of the flag words
        {
            const uint32_t nvec = ctx->pitch / 2;
            const bool direct = ctx->lp_kind != SCORUS_LP_CORR_GAUSSIAN && (SCORUS_G16_DIRECT || !(nvec > 8 && nvec <= 16));
            if (direct && env_int("SCORUS_PEER_STAGE", 1)) {
                const size_t base = g.smem / g.warps, extra = 16u * (size_t)g.row_stride + 128u;
                uint32_t w = g.warps;
                while (w > 1 && w * (base + extra) > (size_t)ctx->smem_optin) --w;
                if (w * (base + extra) <= (size_t)ctx->smem_optin) {
                    g.warps = w;
                    g.smem = w * (base + extra);
                    p.warps = w;
                    p.peer_stage = 1;
                    const uint32_t ntl = (uint32_t)((ctx->n + 31u) / 32u);
                    g.grid = std::min<uint32_t>((uint32_t)ctx->num_sms, (ntl + w - 1) / w);
                }
            }
        }
        cudaError_t e = launch_step(ctx->lp_kind, p, g, all_flags, ctx->stream);
        if (e != cudaSuccess) return fail(ctx, SCORUS_ERR_CUDA, "step launch: %s", cudaGetErrorString(e));
        ++ctx->launches;
        ++ctx->step;
        count_proposed(ctx, nbeta, p.npb, p.w_lo, p.w_hi, 1);
        if (ufs->kind == SCORUS_UFS_ONE_HOT_CYCLE) ctx->onehot = (ctx->onehot + ctx->n_global) % ctx->dim;
        if (ctx->tr_ncap != 0 || ctx->mom_rungs != 0) {  // per-step hooks read the local rows: peers must be done
            if ((rc = scorus_peer_barrier(ctx))) return rc;
            if ((rc = after_step(ctx))) return rc;
        }
    }
    return nsteps ? scorus_peer_barrier(ctx) : SCORUS_OK;
}

// rows of a cross-rank ladder swap, pulled one-sidedly: scratch[slot] <- (peer holding src)[src] for every local slot
// whose walker changes.  One thread per 16-byte chunk of a row (a warp per row left 27 of 32 lanes idle at 10-D and
// queued the rows of a warp behind one another's NVLink latency).
__global__ void __launch_bounds__(256) swap_pull_rows_kernel(const uint32_t *src_all, uint32_t w_lo, uint32_t n_local,
                                                             uint32_t pitch, double *const *peer_ens, double *scratch)
{
    const uint32_t nvec = pitch >> 1;
    const size_t total = (size_t)n_local * nvec;
    for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (size_t)gridDim.x * blockDim.x) {
        const uint32_t i = (uint32_t)(g / nvec), c = (uint32_t)(g - (size_t)i * nvec);
        const uint32_t s = __ldg(src_all + w_lo + i);
        if (s == w_lo + i) continue;
        const uint32_t q = s / n_local, r = s - q * n_local;
        reinterpret_cast<double2 *>(scratch)[g] = ld_stream(reinterpret_cast<const double2 *>(peer_ens[q] + (size_t)r * pitch) + c);
    }
}

__global__ void __launch_bounds__(256) swap_commit_rows_kernel(const uint32_t *src_all, const double *lp_all, uint32_t w_lo,
                                                               uint32_t n_local, uint32_t pitch, const double *scratch,
                                                               double *ens, double *lp)
{
    const uint32_t nvec = pitch >> 1;
    const size_t total = (size_t)n_local * nvec;
    for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (size_t)gridDim.x * blockDim.x) {
        const uint32_t i = (uint32_t)(g / nvec), c = (uint32_t)(g - (size_t)i * nvec);
        if (c == 0) lp[i] = __ldg(lp_all + w_lo + i);
        if (__ldg(src_all + w_lo + i) == w_lo + i) continue;
        reinterpret_cast<double2 *>(ens)[g] = reinterpret_cast<const double2 *>(scratch)[g];
    }
}

// Applies a swap plan (scorus_swap_plan_device) to this rank's slots with one-sided reads of the peers
// mapped by scorus_ipc_attach.  phase 0 pulls every incoming row into scratch (all ranks still hold
// their old rows); the caller then inserts a stream-ordered barrier; phase 1 writes rows and log-probs.
extern "C" int scorus_swap_apply_p2p(scorus_ctx *ctx, const uint32_t *src_all, const double *lp_all, int phase)
{
    if (!ctx || !src_all || !lp_all) return SCORUS_ERR_INVALID_ARG;
    F64_ONLY(ctx);
    if (!ctx->d_peer_ens) return fail(ctx, SCORUS_ERR_INVALID_ARG, "call scorus_ipc_attach first");
    cudaSetDevice(ctx->device);
    if (!ctx->d_scratch) CU(cudaMalloc(&ctx->d_scratch, ctx->n * (size_t)ctx->pitch * sizeof(double)));
    const size_t chunks = ctx->n * (size_t)(ctx->pitch / 2);
    const unsigned blocks = (unsigned)std::min<size_t>((chunks + 255) / 256, (size_t)ctx->num_sms * 16u);
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

// every rank's log-probs into one local array: coalesced one-sided reads of the peers' mappings (8 bytes per walker)
__global__ void __launch_bounds__(256) gather_lp_kernel(const PeerTable t, uint32_t n_local, uint32_t n_global, double *lp_all)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_global; i += gridDim.x * blockDim.x) {
        const uint32_t q = i / n_local;
        double v;  // another GPU wrote it: not through the non-coherent path
        asm volatile("ld.global.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(t.lp[q] + (i - q * n_local)));
        lp_all[i] = v;
    }
}

// swap_walkers (src/mcmc/utils.rs:72-112) of a ladder whose rungs are spread over the attached ranks, in ONE call and
// without NCCL: barrier (every rank's log-probs are final); every rank gathers all log-probs with coalesced one-sided
// reads (a scattered 8-byte NVLink read per chain stage inside the latency-bound chain kernel cost 0.8 ms at 8 GPUs)
// and decides the whole ladder; pulls the rows that move into its slots from wherever they live; barrier (nobody
// overwrites a row a peer still has to read); commit.  Identical to the single-GPU swap.
extern "C" int scorus_swap_walkers_peer(scorus_ctx *ctx, const double *beta, size_t nbeta)
{
    if (!ctx || !beta) return SCORUS_ERR_INVALID_ARG;
    F64_ONLY(ctx);
    if (!ctx->d_peer_ens || ctx->world < 1) return fail(ctx, SCORUS_ERR_INVALID_ARG, "call scorus_ipc_attach first");
    const size_t n = ctx->n_global;
    if (n != ctx->n * (size_t)ctx->world) return fail(ctx, SCORUS_ERR_INVALID_ARG, "shards must be equal");
    if (nbeta == 0) return fail(ctx, SCORUS_ERR_INVALID_ARG, "nbeta == 0");
    const size_t npb = n / nbeta;
    if (npb * nbeta != n) return fail(ctx, SCORUS_ERR_NWALKERS_MISMATCHES_NBETA, "NWalkersMismatchesNBeta");
    int rc = scorus_check_beta_order(beta, nbeta);
    if (rc) return fail(ctx, rc, "beta list must be in decreasing order");
    if (!ctx->uploaded) return fail(ctx, SCORUS_ERR_NOT_UPLOADED, "call scorus_upload first");
    cudaSetDevice(ctx->device);
    if (nbeta == 1) { ++ctx->swap_call; return SCORUS_OK; }
    if ((rc = upload_beta(ctx, beta, nbeta))) return rc;
    if ((rc = ensure(ctx, (void **)&ctx->d_lp_all, &ctx->lp_all_cap, n * sizeof(double)))) return rc;
    if ((rc = ensure(ctx, (void **)&ctx->d_src_all, &ctx->src_all_cap, n * sizeof(uint32_t)))) return rc;
    if (!ctx->d_scratch) CU(cudaMalloc(&ctx->d_scratch, ctx->n * (size_t)ctx->pitch * sizeof(double)));
    // SCORUS_SWAP_DIAG=1 (measurement only): CUDA events between the phases, printed by rank 0 after a synchronise
    const bool diag = env_int("SCORUS_SWAP_DIAG", 0) != 0;
    cudaEvent_t ev[7];
    int nev = 0;
    auto mark = [&]() { if (diag) { cudaEventCreate(&ev[nev]); cudaEventRecord(ev[nev], ctx->stream); ++nev; } };
    mark();
    if ((rc = scorus_peer_barrier(ctx))) return rc;
    mark();
    SwapParams sp;
    memset(&sp, 0, sizeof(sp));
    PeerTable t;
    for (int q = 0; q < 16; ++q) t.lp[q] = ctx->peer_lp[q];
    // one element per thread: every NVLink read in flight at once (a grid-stride loop of seven dependent-latency reads
    // per thread took 32 us for 8 MB at 8 GPUs)
    gather_lp_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(t, (uint32_t)ctx->n, (uint32_t)n, ctx->d_lp_all);
    ++ctx->launches;
    mark();
    sp.lp = ctx->d_lp_all; sp.lp_src = ctx->d_lp_all; sp.beta = ctx->d_beta; sp.src = ctx->d_src_all; sp.stats = ctx->d_stats;
    if ((rc = ensure_rung_stats(ctx, nbeta))) return rc;
    sp.stage_swapped = ctx->d_rung_stats + ctx->rung_cap;
    sp.seed = ctx->seed; sp.swap_call = ctx->swap_call;
    sp.npb = (uint32_t)npb; sp.nbeta = (uint32_t)nbeta; sp.bij_bits = bij_bits((uint32_t)npb);
    if (npb <= kExactShuffleMax) {
        const size_t m = (nbeta - 1) * npb;
        if ((rc = ensure(ctx, (void **)&ctx->d_jinv, &ctx->jinv_cap, m * sizeof(uint32_t)))) return rc;
        launch_exact_jvec(ctx, nbeta, npb, ctx->d_jinv);
        sp.jinv = ctx->d_jinv;
    }
    if ((rc = launch_swap_chain(ctx, sp))) return rc;
    mark();
    ++ctx->swap_call;
    count_swap_proposed(ctx, nbeta, npb);
    if ((rc = scorus_swap_apply_p2p(ctx, ctx->d_src_all, ctx->d_lp_all, 0))) return rc;
    mark();
    if ((rc = scorus_peer_barrier(ctx))) return rc;
    mark();
    rc = scorus_swap_apply_p2p(ctx, ctx->d_src_all, ctx->d_lp_all, 1);
    mark();
    if (diag) {
        cudaStreamSynchronize(ctx->stream);
        float ms[6];
        for (int k = 0; k + 1 < nev; ++k) cudaEventElapsedTime(&ms[k], ev[k], ev[k + 1]);
        if (ctx->rank == 0)
            fprintf(stderr, "swap_walkers_peer us: barrier %.1f gather_lp %.1f plan %.1f pull %.1f barrier %.1f commit %.1f\n",
                    1e3 * ms[0], 1e3 * ms[1], 1e3 * ms[2], 1e3 * ms[3], 1e3 * ms[4], 1e3 * ms[5]);
        for (int k = 0; k < nev; ++k) cudaEventDestroy(ev[k]);
    }
    return rc;
}
