// step_kernel.cuh -- K1: the fused stretch-move step (and K2: init_logprob).
//
// One launch performs, for every walker, everything src/mcmc/ensemble_sample.rs:135-189 does
// in six serial/parallel host stages: partner pick (:135-148), z draw (:149 -> utils.rs:33-45),
// update flags (:150-153 -> :38-54), proposal (:155-159 -> :57-75 -> utils.rs:47-56),
// log-density (:161) and Metropolis accept (:164-189).
//
// Shape of the work
//  * The pairing is a perfect matching, so a PAIR owns its two rows exclusively: both rows are
//    read once, both walkers are updated from the old values, and accepted rows are written
//    back in place.  HBM traffic is 16*D+16 bytes per walker-step, the algorithmic minimum.
//  * A tile is 32 matching positions = 16 pairs = one warp.  Lane l owns the walker at
//    position 32*t+l; its partner is lane l^1.  Rows are random gathers of pitch*8 contiguous
//    bytes, moved by the TMA engine (cp.async.bulk global->shared, completion on an mbarrier)
//    into a warp-private ring of shared-memory stages; accepted rows leave through
//    cp.async.bulk shared->global.  No thread ever issues a per-element global load of a row.
//  * Each lane then walks ITS row sequentially (thread-per-walker): per-walker sums have the
//    reference's left-to-right order, and the scalar work per walker (Philox, log, exp) is
//    done once per lane instead of once per warp.  Shared-memory rows are padded to an odd
//    number of 16-byte units so the 32 lanes' 128-bit accesses are bank-conflict free.
//  * Warps are independent: no __syncthreads in the loop, only __syncwarp between the read
//    and the in-place write of a chunk (lane l^1 reads lane l's row).
#pragma once
#include "tile_common.cuh"

namespace scorus {

constexpr int kChunk = 8;  // doubles per lane between two __syncwarp (4 x LDS.128 per row)

// Shared-memory carve-up (dynamic):
//   [0, 8*warps*stages)                  mbarriers, one per (warp, stage)
//   then per warp: stages * 32 * row_stride bytes of rows, then 32*flag_words*4 bytes of flag bits
template <class Plugin, bool kAllFlags>
__global__ void __launch_bounds__(512, 1) step_seq_kernel(const StepParams p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t S = p.stages;
    const uint32_t tile_bytes = 32u * p.row_stride;
    const uint32_t bars_bytes = (8u * p.warps * S + 127u) & ~127u;
    const uint32_t warp_bytes = S * tile_bytes + ((p.flag_words * 128u + 127u) & ~127u);
    unsigned char *wbase = smem + bars_bytes + (size_t)warp * warp_bytes;
    uint32_t *flagw = reinterpret_cast<uint32_t *>(wbase + S * tile_bytes);
    const uint32_t bar0 = smem_u32(smem) + 8u * warp * S;

    if (lane == 0)
        for (uint32_t s = 0; s < S; ++s) mbar_init(bar0 + 8u * s, 1);
    fence_mbar_init();
    __syncthreads();

    const uint32_t gwarp = blockIdx.x * p.warps + warp;
    const uint32_t nwarps = gridDim.x * p.warps;
    unsigned long long n_acc = 0;

    // ---- per-lane description of a tile that has been prefetched into a stage ----
    // ring position q (0..S-1) holds tile (gwarp + (issued index) * nwarps)
    auto prefetch = [&](uint32_t tile, uint32_t stage, uint32_t &w_out, double &lp_out, bool &act_out) {
        const uint32_t j = tile * 32u + lane;
        const bool act = tile < p.ntiles && j < p.n;
        uint32_t w = 0;
        if (act) w = walker_at(p, j);
        const uint32_t nact = __popc(__ballot_sync(0xffffffffu, act));
        const uint32_t bar = bar0 + 8u * stage;
        if (lane == 0 && nact) mbar_expect_tx(bar, nact * p.row_bytes);
        __syncwarp();
        double lpv = 0.0;
        if (act) {
            bulk_load(smem_u32(wbase + stage * tile_bytes + lane * p.row_stride),
                      p.ens + (size_t)w * p.pitch, p.row_bytes, bar);
            lpv = __ldg(p.lp + w);  // register prefetch, consumed one tile later
        }
        w_out = w;
        lp_out = lpv;
        act_out = act;
    };

    // ring bookkeeping kept in registers (S <= 4)
    uint32_t rw[4];
    double rlp[4];
    bool ract[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) { rw[s] = 0; rlp[s] = 0.0; ract[s] = false; }

    // prologue: fill S-1 stages
    uint32_t issued = 0;
    for (; issued + 1 < S; ++issued) {
        uint32_t w; double l; bool a;
        prefetch(gwarp + issued * nwarps, issued, w, l, a);
#pragma unroll
        for (int s = 0; s < 4; ++s) if ((uint32_t)s == issued) { rw[s] = w; rlp[s] = l; ract[s] = a; }
    }

    for (uint32_t k = 0;; ++k) {
        const uint32_t tile = gwarp + k * nwarps;
        if (tile >= p.ntiles) break;
        const uint32_t stage = k % S;
        const uint32_t parity = (k / S) & 1u;

        // refill the stage freed by tile k-1: its stores must have finished reading shared memory
        {
            bulk_wait_read_all();
            __syncwarp();
            const uint32_t ps = (k + S - 1) % S;
            uint32_t w; double l; bool a;
            prefetch(gwarp + (k + S - 1) * nwarps, ps, w, l, a);
#pragma unroll
            for (int s = 0; s < 4; ++s) if ((uint32_t)s == ps) { rw[s] = w; rlp[s] = l; ract[s] = a; }
        }

        uint32_t w = 0; double lp_old = 0.0; bool act = false;
#pragma unroll
        for (int s = 0; s < 4; ++s) if ((uint32_t)s == stage) { w = rw[s]; lp_old = rlp[s]; act = ract[s]; }

        // ---- scalar streams for this walker ----
        double z, u;
        if (p.z_in) {
            z = act ? __ldg(p.z_in + w) : 1.0;
            u = act ? __ldg(p.u_in + w) : 1.0;
        } else {
            U4 r = draw(p.seed, p.step, ST_ZU, w + p.w_off, 0);
            z = z_from_uniform(to_unit<double>(r.x, r.y), p.inv_sqrt_a, p.z_range);
            u = to_unit<double>(r.z, r.w);
        }

        // ---- update flags (:38-54) ----
        uint32_t nphi = p.dim;
        if (!kAllFlags) nphi = gen_flags(p, w, flagw, lane, act);

        // ---- wait for the rows, then propose + evaluate in one pass over the row ----
        mbar_wait(bar0 + 8u * stage, parity);

        unsigned char *own_b = wbase + stage * tile_bytes + lane * p.row_stride;
        unsigned char *par_b = wbase + stage * tile_bytes + (lane ^ 1u) * p.row_stride;
        double2 *own = reinterpret_cast<double2 *>(own_b);
        const double2 *par = reinterpret_cast<const double2 *>(par_b);
        const double omz = 1.0 - z;
        const uint32_t nvec = p.pitch >> 1;  // double2 per row

        Plugin plug;
        plug.init(p.pparams, p.nparams, p.dim);

        for (uint32_t c = 0; c < nvec; c += kChunk / 2) {
            double2 a[kChunk / 2], b[kChunk / 2];
#pragma unroll
            for (int q = 0; q < kChunk / 2; ++q) {
                if (c + q < nvec) { a[q] = own[c + q]; b[q] = par[c + q]; }
                else { a[q] = make_double2(0.0, 0.0); b[q] = a[q]; }
            }
            __syncwarp();  // lane^1 has read my row before I overwrite it
            uint32_t fbits = 0xFFFFFFFFu;
            if (!kAllFlags) fbits = flagw[((c * 2u) >> 5) * 32u + lane] >> ((c * 2u) & 31u);
#pragma unroll
            for (int q = 0; q < kChunk / 2; ++q) {
                const uint32_t d0 = (c + q) * 2u;
                if (d0 < p.dim) {
                    bool f0 = true, f1 = true;
                    if (!kAllFlags) { f0 = (fbits >> (2 * q)) & 1u; f1 = (fbits >> (2 * q + 1)) & 1u; }
                    // scale_vec, src/mcmc/utils.rs:55: (x1*z) + (x2*(1-z)), three roundings
                    double t1 = a[q].x * z, t2 = b[q].x * omz;
                    double y0 = t1 + t2;
                    if (!f0) y0 = a[q].x;  // propose_move, ensemble_sample.rs:69-71
                    plug.accum(d0, y0);
                    a[q].x = y0;
                    if (d0 + 1u < p.dim) {
                        double s1 = a[q].y * z, s2 = b[q].y * omz;
                        double y1 = s1 + s2;
                        if (!f1) y1 = a[q].y;
                        plug.accum(d0 + 1u, y1);
                        a[q].y = y1;
                    }
                    own[c + q] = a[q];
                }
            }
        }

        double lp_new;
        if constexpr (Plugin::kNeedsRow) {
            lp_new = plug.finish_row(reinterpret_cast<const double *>(own_b));
        } else {
            lp_new = plug.finish();
        }

        // ---- Metropolis accept, ensemble_sample.rs:181-188 ----
        const double beta = __ldg(p.beta + (act ? w / p.npb : 0u));
        const double delta_lp = lp_new - lp_old;
        const double q = exp((double)(nphi - 1u) * log(z) + delta_lp * beta);
        const bool accept = act && (u < q);

        if (accept) {
            fence_proxy_async();  // my generic-proxy writes of the row -> visible to the TMA store
            bulk_store(p.ens + (size_t)w * p.pitch, smem_u32(own_b), p.row_bytes);
            p.lp[w] = lp_new;
            if (p.dirty) p.dirty[w] = 1;
            ++n_acc;
        }
        bulk_commit();
        if (p.accepted_out && act) p.accepted_out[w] = accept ? 1 : 0;
        if (p.rung_acc) count_rung_accepts(p.rung_acc, act ? w / p.npb : 0u, act, accept, lane);
    }

    bulk_wait_all();
    // per-warp statistics: one atomic per warp per launch
    for (int off = 16; off; off >>= 1) n_acc += __shfl_xor_sync(0xffffffffu, n_acc, off);
    if (lane == 0 && n_acc) atomicAdd(p.stats, n_acc);
}

// K2: init_logprob (src/mcmc/ensemble_sample.rs:77-85).  Same staging, identity order, no move.
template <class Plugin>
__global__ void __launch_bounds__(512, 1) init_logprob_kernel(const StepParams p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t tile_bytes = 32u * p.row_stride;
    const uint32_t bars_bytes = (8u * p.warps + 127u) & ~127u;
    unsigned char *wbase = smem + bars_bytes + (size_t)warp * tile_bytes;
    const uint32_t bar = smem_u32(smem) + 8u * warp;
    if (lane == 0) mbar_init(bar, 1);
    fence_mbar_init();
    __syncthreads();

    const uint32_t gwarp = blockIdx.x * p.warps + warp;
    const uint32_t nwarps = gridDim.x * p.warps;
    uint32_t parity = 0;
    for (uint32_t tile = gwarp; tile < p.ntiles; tile += nwarps) {
        const uint32_t w = tile * 32u + lane;
        const bool act = w < p.n;
        const uint32_t nact = __popc(__ballot_sync(0xffffffffu, act));
        if (lane == 0) mbar_expect_tx(bar, nact * p.row_bytes);
        __syncwarp();
        unsigned char *own_b = wbase + lane * p.row_stride;
        if (act) bulk_load(smem_u32(own_b), p.ens + (size_t)w * p.pitch, p.row_bytes, bar);
        mbar_wait(bar, parity);
        parity ^= 1u;
        Plugin plug;
        plug.init(p.pparams, p.nparams, p.dim);
        const double *row = reinterpret_cast<const double *>(own_b);
        double v;
        if constexpr (Plugin::kNeedsRow) {
            v = act ? plug.finish_row(row) : 0.0;
        } else {
            if (act)
                for (uint32_t d = 0; d < p.dim; ++d) plug.accum(d, row[d]);
            v = plug.finish();
        }
        if (act) p.lp[w] = v;
        __syncwarp();  // all lanes done reading before the next bulk load overwrites the tile
    }
}

}  // namespace scorus
