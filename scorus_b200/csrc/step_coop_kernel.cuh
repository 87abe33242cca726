// step_coop_kernel.cuh -- K1: the fused stretch-move step, warp-cooperative version (default).
//
// One launch does everything src/mcmc/ensemble_sample.rs:135-189 does per step: partner pick
// (:135-148), z draw (:149 -> utils.rs:33-45), update flags (:150-153 -> :38-54), proposal
// (:155-159 -> :57-75 -> utils.rs:47-56), log-density (:161) and Metropolis accept (:164-189).
//
// Data movement is the same as in step_kernel.cuh: a tile is 32 matching positions = 16 pairs =
// one warp; each pair's two rows are gathered once by TMA bulk copies into a warp-private ring
// of shared-memory stages, both walkers of the pair are updated from the old values, accepted
// rows go back in place through TMA bulk stores (16*D+16 bytes per walker-step, the minimum).
//
// What differs is who computes.  The sequential kernel gives a lane a whole row; with only
// one or two warps per scheduler (shared memory bounds the rows in flight) it is bound by
// instruction latency.  Here the work is split into phases that each use all 32 lanes:
//   P1  lane = walker: Philox z / u, flag words, nphi, (nphi-1) ln z            (scalar work,
//       done once per walker but 32 walkers per instruction)
//   P2  a group of G lanes shares a pair; lane j owns double2 chunks j, j+G, ... of BOTH rows:
//       128-bit conflict-free LDS, y = x1*z + x2*(1-z) for both walkers, 128-bit STS in place.
//       A lane only ever touches its own chunks, so the 16 pair bodies are independent
//       straight-line code (high ILP) with no barrier between them.
//   P3  same lane/chunk ownership: log-density terms of the proposed rows (neighbour
//       coordinate re-read from shared memory for Rosenbrock), one partial per (lane, walker).
//   P4  transposed shuffle-tree reduction: G partials per lane in, lane = walker out.
//   P5  lane = walker: exp, accept, TMA store of the accepted row, log-prob store.
// G = 2..32 is the power of two covering a row's double2 count (D = 64: G = 32, one chunk per
// lane; D = 10: G = 8), so small-D ensembles process 32/G pairs at a time.
#pragma once
#include "tile_common.cuh"

namespace scorus {

// v[i], i < G, holds this lane's partial of value i; on return lane (g*G + i) holds the total of
// value i over the G lanes of its group.  Fixed tree: results do not depend on anything but G.
template <int G>
__device__ __forceinline__ double transpose_reduce(double (&v)[G], uint32_t lane)
{
#pragma unroll
    for (int h = G / 2; h >= 1; h >>= 1) {
        const bool up = (lane & (uint32_t)h) != 0;
#pragma unroll
        for (int i = 0; i < h; ++i) {
            const double send = up ? v[i] : v[i + h];
            const double keep = up ? v[i + h] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, h);
        }
    }
    return v[0];
}

template <class Plugin, int G>
__global__ void __launch_bounds__(512, 1) step_coop_kernel(const StepParams p)
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
    const uint32_t grp = lane / G, j = lane % G;  // group within the warp, lane within the group
    const uint32_t nvec = p.pitch >> 1;           // double2 chunks per row
    const bool all_flags = p.flag_kind == 1u;
    unsigned long long n_acc = 0;

    Plugin plug;
    plug.init(p.pparams, p.nparams, p.dim);

    auto prefetch = [&](uint32_t tile, uint32_t stage, uint32_t &w_out, double &lp_out, bool &act_out) {
        const uint32_t pos = tile * 32u + lane;
        const bool act = tile < p.ntiles && pos < p.n;
        uint32_t w = 0;
        if (act) w = walker_at(p, pos);
        const uint32_t nact = __popc(__ballot_sync(0xffffffffu, act));
        const uint32_t bar = bar0 + 8u * stage;
        if (lane == 0 && nact) mbar_expect_tx(bar, nact * p.row_bytes);
        __syncwarp();
        double lpv = 0.0;
        if (act) {
            bulk_load(smem_u32(wbase + stage * tile_bytes + lane * p.row_stride),
                      p.ens + (size_t)w * p.pitch, p.row_bytes, bar);
            lpv = __ldg(p.lp + w);  // register prefetch, consumed when the tile is computed
        }
        w_out = w;
        lp_out = lpv;
        act_out = act;
    };

    uint32_t rw[4];
    double rlp[4];
    bool ract[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) { rw[s] = 0; rlp[s] = 0.0; ract[s] = false; }

    for (uint32_t issued = 0; issued + 1 < S; ++issued) {
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

        {   // refill the stage freed by tile k-1 (its stores must have finished reading shared memory)
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

        // ---- P1: lane = walker ----
        double z, u;
        if (p.z_in) {
            z = act ? __ldg(p.z_in + w) : 1.0;
            u = act ? __ldg(p.u_in + w) : 1.0;
        } else {
            const U4 r = draw(p.seed, p.step, ST_ZU, w + p.w_off, 0);
            z = z_from_uniform(to_unit<double>(r.x, r.y), p.inv_sqrt_a, p.z_range);
            u = to_unit<double>(r.z, r.w);
        }
        uint32_t nphi = p.dim;
        if (!all_flags) nphi = gen_flags(p, w, flagw, lane, act);
        const double lz = (double)(nphi - 1u) * log(z);  // ensemble_sample.rs:184, first summand
        const double beta = __ldg(p.beta + (act ? w / p.npb : 0u));
        __syncwarp();  // flag words visible to the whole warp

        mbar_wait(bar0 + 8u * stage, parity);
        unsigned char *tbase = wbase + stage * tile_bytes;

        // ---- P2: proposals, both walkers of a pair, in place ----
        // (rolled loops on purpose: the fully unrolled body overflowed the instruction cache and
        //  the kernel stalled on instruction fetch -- profiles/r1_notes.md)
#pragma unroll 2
        for (int pr = 0; pr < G / 2; ++pr) {
            const uint32_t wa = grp * G + 2u * pr, wb = wa + 1u;  // tile-local walkers of this pair
            const double za = __shfl_sync(0xffffffffu, z, wa), zb = __shfl_sync(0xffffffffu, z, wb);
            const double omza = 1.0 - za, omzb = 1.0 - zb;
            double2 *rowa = reinterpret_cast<double2 *>(tbase + wa * p.row_stride);
            double2 *rowb = reinterpret_cast<double2 *>(tbase + wb * p.row_stride);
#pragma unroll 1
            for (uint32_t c = j; c < nvec; c += G) {
                const double2 xa = rowa[c], xb = rowb[c];
                const uint32_t d = 2u * c;
                uint32_t fa = 3u, fb = 3u;
                if (!all_flags) {
                    fa = flagw[(d >> 5) * 32u + wa] >> (d & 31u);
                    fb = flagw[(d >> 5) * 32u + wb] >> (d & 31u);
                }
                double2 ya, yb;
                // scale_vec, src/mcmc/utils.rs:55: (x1*z) + (x2*(1-z)), three roundings; then
                // propose_move, ensemble_sample.rs:69-71: coordinates whose flag is clear stay
                { double t1 = xa.x * za, t2 = xb.x * omza; ya.x = (fa & 1u) ? t1 + t2 : xa.x; }
                { double t1 = xa.y * za, t2 = xb.y * omza; ya.y = (fa & 2u) ? t1 + t2 : xa.y; }
                { double t1 = xb.x * zb, t2 = xa.x * omzb; yb.x = (fb & 1u) ? t1 + t2 : xb.x; }
                { double t1 = xb.y * zb, t2 = xa.y * omzb; yb.y = (fb & 2u) ? t1 + t2 : xb.y; }
                rowa[c] = ya;
                rowb[c] = yb;
            }
        }
        fence_proxy_async();  // my writes of the rows -> visible to the TMA store issued in P5
        __syncwarp();         // P3 reads neighbour coordinates written by other lanes

        // ---- P3 + P4: log-density terms of the proposed rows, then lane = walker again ----
        // Blocks of GB <= 8 walkers bound both the partials held in registers and the code size.
        constexpr int GB = G < 8 ? G : 8;
        double tot[Plugin::kAcc];
#pragma unroll
        for (int a = 0; a < Plugin::kAcc; ++a) tot[a] = 0.0;
        if constexpr (!Plugin::kNeedsRow) {
#pragma unroll 1
            for (int blk = 0; blk < G / GB; ++blk) {
                double acc[Plugin::kAcc][GB];
#pragma unroll
                for (int i = 0; i < GB; ++i) {
                    const unsigned char *rb = tbase + (grp * G + (uint32_t)(blk * GB + i)) * p.row_stride;
                    const double2 *row2 = reinterpret_cast<const double2 *>(rb);
                    const double *row1 = reinterpret_cast<const double *>(rb);
                    double part[Plugin::kAcc];
#pragma unroll
                    for (int a = 0; a < Plugin::kAcc; ++a) part[a] = 0.0;
#pragma unroll 1
                    for (uint32_t c = j; c < nvec; c += G) {
                        const uint32_t d = 2u * c;
                        const double2 y = row2[c];
                        const bool has1 = d + 1u < p.dim, hasnb = d + 2u < p.dim;
                        const double nb = hasnb ? row1[d + 2u] : 0.0;
                        plug.terms(d, y.x, y.y, nb, has1, hasnb, part);
                    }
#pragma unroll
                    for (int a = 0; a < Plugin::kAcc; ++a) acc[a][i] = part[a];
                }
#pragma unroll
                for (int a = 0; a < Plugin::kAcc; ++a) {
                    double r = transpose_reduce<GB>(acc[a], lane);
#pragma unroll
                    for (int h = GB; h < G; h <<= 1) r += __shfl_xor_sync(0xffffffffu, r, h);
                    if ((int)(j / GB) == blk) tot[a] = r;
                }
            }
        }

        // ---- P5: Metropolis accept, ensemble_sample.rs:181-188 ----
        unsigned char *own_b = tbase + lane * p.row_stride;
        double lp_new;
        if constexpr (Plugin::kNeedsRow) {
            lp_new = plug.finish_row(reinterpret_cast<const double *>(own_b));
        } else {
            lp_new = plug.finish_acc(tot);
        }
        const double delta_lp = lp_new - lp_old;
        const double q = exp(lz + delta_lp * beta);
        const bool accept = act && (u < q);
        if (accept) {
            fence_proxy_async();  // generic-proxy writes of the row -> visible to the TMA store
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
    for (int off = 16; off; off >>= 1) n_acc += __shfl_xor_sync(0xffffffffu, n_acc, off);
    if (lane == 0 && n_acc) atomicAdd(p.stats, n_acc);
}

}  // namespace scorus
