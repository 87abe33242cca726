// twalk_direct_kernel.cuh -- the t-walk half-pass kernel (twalk_kernel.cuh: same algorithm, same streams, same
// results) without the shared-memory tile: proposals stay in registers, the accept runs per block of rows and
// accepted rows go straight from the group's registers to global memory, as in step_reg_kernel.cuh's direct mode.
// Shared memory holds, per warp, the flag words and a short list of Walk z values per walker; the block size is
// bound by registers (16 warps per SM instead of 11 at 64-D).
//
// Fused mode (half = 2) needs no special case for the partner that has just moved: the second half-pass re-reads
// its rows from global memory, and the lane that wrote chunk j of an accepted row in the first half-pass is the
// lane that reads chunk j of that row in the second (same thread, same address: program order).  The prefetch of
// the second half-pass's first block is issued after the first half-pass's stores of that block.
//
// Used for the plugins whose log-density comes from per-coordinate terms and rows that are not 9..16 double2
// (the dense target and 16-lane groups keep twalk_kernel.cuh, for the reasons step_reg_kernel.cuh gives).
#pragma once
#include "twalk_kernel.cuh"

namespace scorus {

constexpr uint32_t kTwalkZList = 8;  // Walk z values parked per walker; a walker with more flagged coordinates draws in the loop

template <class Plugin, int G, int ITERS>
__global__ void __launch_bounds__(twalk_max_warps<Plugin, G>() * 32, 1) twalk_direct_kernel(const TwalkParams tp)
{
    constexpr int kWant = SCORUS_TW_ROWS / ITERS < 2 ? 2 : SCORUS_TW_ROWS / ITERS;
    constexpr int GB = G < kWant ? G : kWant, PB = GB / 2, NB = G / GB;
    const StepParams &p = tp.s;

    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t flag_bytes = (p.flag_words * 128u + 127u) & ~127u;
    const uint32_t warp_bytes = flag_bytes + 32u * kTwalkZList * (uint32_t)sizeof(double);
    uint32_t *flagw = reinterpret_cast<uint32_t *>(smem + (size_t)warp * warp_bytes);
    double *zlist = reinterpret_cast<double *>(smem + (size_t)warp * warp_bytes + flag_bytes);

    const uint32_t gwarp = blockIdx.x * p.warps + warp;
    const uint32_t nwarps = gridDim.x * p.warps;
    const uint32_t grp = lane / G, j = lane % G;
    const uint32_t nvec = p.pitch >> 1;
    const bool both = tp.half == 2u;
    const uint32_t h_begin = both ? 0u : tp.half, h_end = both ? 2u : tp.half + 1u;
    unsigned long long n_acc = 0;

    grid_dep_launch();

    Plugin plug;
    plug.init(p.pparams, p.nparams, p.dim);

    uint32_t key_rung = 0xFFFFFFFFu;
    BijKeys K;
    auto walker_of = [&](uint32_t pos) -> uint32_t {
        if (p.order) return __ldg(p.order + pos);
        const uint32_t r = pos / p.npb;
        if (r != key_rung) { K = bij_keys(p.seed, tp.order_step, ST_PERM, r + p.rung_off); key_rung = r; }
        return r * p.npb + bij(pos - r * p.npb, p.npb, p.bij_bits, K);
    };
    uint64_t pol_keep = 0, pol_drop = 0;
    if (both && tp.l2_keep) {
        if (tp.l2_keep == 1u) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));
        else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol_keep));
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_drop));
    }
    const bool hinted = both && tp.l2_keep != 0u;
    // loads block `blk` of a tile for half-pass h: x[2k] = the moving member of pair k, x[2k+1] = its partner
    auto load_block = [&](double2 (&x)[GB][ITERS], uint32_t wt, int blk, uint32_t h) {
        const uint64_t pol = h == 0u ? pol_keep : pol_drop;
#pragma unroll
        for (int i = 0; i < GB; ++i) {
            const uint32_t wl = grp * G + ((uint32_t)(blk * GB + i) ^ h);
            const uint32_t row = __shfl_sync(0xffffffffu, wt, wl);
            const double2 *src = reinterpret_cast<const double2 *>(p.ens + (size_t)row * p.pitch);
#pragma unroll
            for (int it = 0; it < ITERS; ++it) {
                const uint32_t c = (uint32_t)it * G + j;
                const double2 *q = src + (c < nvec ? c : nvec - 1u);
                x[i][it] = hinted ? ld_stream_hint(q, pol) : ld_stream(q);
            }
        }
    };

    grid_dep_wait();

    const uint32_t ntiles = (p.n + 31u) >> 5;
    uint32_t tile = gwarp;
    if (tile >= ntiles) return;
    uint32_t w;
    bool act;
    {
        const uint32_t pos = tile * 32u + lane;
        act = pos < p.n;
        w = act ? walker_of(pos) : 0u;
    }
    const uint32_t my_half = lane & 1u;
    const bool lane_in = both || my_half == tp.half;
    const uint64_t my_step = p.step + (both ? my_half : 0u);
    double lp_old = act && lane_in ? __ldg(p.lp + w) : 0.0;
    double2 xn[GB][ITERS];
    load_block(xn, w, 0, h_begin);

    for (;;) {
        const bool mv = act && lane_in;
        // ---- P1: lane = walker: kernel choice, b, accept uniform, flags, Walk z list ----
        uint32_t kern = 0;
        double b = 1.0, u = 1.0;
        if (tp.kernel_in) {
            if (mv) { kern = __ldg(tp.kernel_in + w); b = __ldg(tp.b_in + w); u = __ldg(p.u_in + w); }
        } else if (mv) {
            const U4 r = draw(p.seed, my_step, ST_ZU, w + p.w_off, 0);
            kern = u53(r.x, r.y) * tp.fw1 < tp.fw0 ? 0u : 1u;
            u = u53(r.z, r.w);
            if (kern == 1u) {
                const U4 q = draw(p.seed, my_step, ST_TW_B, w + p.w_off, 0);
                b = twalk_b(u53(q.x, q.y), u53(q.z, q.w), tp.at);
            }
        }
        __syncwarp();  // the previous tile's readers are done with the flag words and the z lists
        const uint32_t nphi = gen_flags(p, my_step, w, flagw, lane, mv);
        const double lb = kern == 1u ? (double)((int)nphi - 2) * log(b) : 0.0;
        // ln u once per tile with every lane busy: the per-block decision below compares logarithms and takes the exp only
        // inside the guard band (step_reg_kernel.cuh has the argument)
        [[maybe_unused]] double lu = 0.0;
        if constexpr (NB > 1) lu = u > 0.0 ? log(u) : real_nan<double>();
        const double beta = __ldg(p.beta + (act ? w / p.npb : 0u));
        // flagged coordinates before each flag word (8 bits per word: dim <= 256), and the z of the first
        // kTwalkZList flagged coordinates in increasing order (sim_walk :289-296; one Philox call per flagged chunk)
        uint64_t pref = 0;
        {
            uint32_t cnt = 0;
            for (uint32_t wd = 0; wd < p.flag_words; ++wd) {
                pref |= (uint64_t)cnt << (8u * wd);
                uint32_t bits = flagw[wd * 32u + lane];
                if (mv && kern == 0u) {
                    uint32_t k = cnt;
                    while (bits && k < kTwalkZList) {
                        const uint32_t sh = (uint32_t)(__ffs((int)bits) - 1) & ~1u;
                        const uint32_t d0 = wd * 32u + sh, two = (bits >> sh) & 3u;
                        double u0, u1;
                        if (tp.walk_u_in) {
                            const double *wu = tp.walk_u_in + (size_t)w * p.dim + d0;
                            u0 = __ldg(wu);
                            u1 = d0 + 1u < p.dim ? __ldg(wu + 1) : 0.0;
                        } else {
                            const U4 r = draw(p.seed, my_step, ST_TW_WALK, w + p.w_off, d0 >> 1);
                            u0 = u53(r.x, r.y);
                            u1 = u53(r.z, r.w);
                        }
                        if (two & 1u) { zlist[lane * kTwalkZList + k] = tp.aw_ratio * (tp.aw * (u0 * u0) + 2.0 * u0 - 1.0); ++k; }
                        if ((two & 2u) && k < kTwalkZList) { zlist[lane * kTwalkZList + k] = tp.aw_ratio * (tp.aw * (u1 * u1) + 2.0 * u1 - 1.0); ++k; }
                        bits &= ~(3u << sh);
                    }
                }
                cnt += __popc(flagw[wd * 32u + lane]);
            }
        }
        __syncwarp();

        uint32_t tile_n = tile + nwarps;
        if (p.tile_counter) {
            unsigned long long t = 0;
            if (lane == 0) t = atomicAdd(p.tile_counter, 1ull) - p.tile_base;
            t = __shfl_sync(0xffffffffu, t, 0);
            tile_n = t + nwarps > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)(t + nwarps);
        }
        uint32_t w_n = 0;
        bool act_n = false;
        if (tile_n < ntiles) {
            const uint32_t pos = tile_n * 32u + lane;
            act_n = pos < p.n;
            w_n = act_n ? walker_of(pos) : 0u;
        }
        const double lp_n = act_n && lane_in ? __ldg(p.lp + w_n) : 0.0;

        bool accept = false;
#pragma unroll 1
        for (uint32_t half = h_begin; half < h_end; ++half) {
            const uint64_t hstep = p.step + (both ? half : 0u);
            double tot[Plugin::kAcc];
            bool my_eq = false;
#pragma unroll
            for (int a = 0; a < Plugin::kAcc; ++a) tot[a] = 0.0;

#pragma unroll 1
            for (int blk = 0; blk < NB; ++blk) {
                double2 xc[GB][ITERS];
#pragma unroll
                for (int i = 0; i < GB; ++i)
#pragma unroll
                    for (int it = 0; it < ITERS; ++it) xc[i][it] = xn[i][it];
                // next block: this half-pass's, the next tile's -- or, after a first half-pass, this tile's block 0
                // again, which must wait for this block's stores when the tile is a single block
                const bool again = blk + 1 == NB && half + 1u < h_end;
                if (blk + 1 < NB) load_block(xn, w, blk + 1, half);
                else if (again) { if (NB > 1) load_block(xn, w, 0, half + 1u); }
                else if (tile_n < ntiles) load_block(xn, w_n, 0, h_begin);

                double acc[Plugin::kAcc][GB];
                double2 yk[PB][ITERS];
#pragma unroll
                for (int pr = 0; pr < PB; ++pr) {
                    const uint32_t wm = grp * G + (uint32_t)(blk * GB + 2 * pr) + half;
                    const uint32_t kern_m = __shfl_sync(0xffffffffu, kern, wm);
                    const double b_m = __shfl_sync(0xffffffffu, b, wm);
                    const uint32_t id_m = __shfl_sync(0xffffffffu, w, wm);
                    const uint32_t nphi_m = __shfl_sync(0xffffffffu, nphi, wm);
                    const uint64_t pref_m = __shfl_sync(0xffffffffu, pref, wm);
                    bool eq = false;
#pragma unroll
                    for (int it = 0; it < ITERS; ++it) {
                        const uint32_t c = (uint32_t)it * G + j;
                        const uint32_t d = 2u * c;
                        const double2 xm = xc[2 * pr][it], xp = xc[2 * pr + 1][it];
                        const uint32_t wd = (d >> 5) < p.flag_words ? (d >> 5) : 0u;
                        const uint32_t word = flagw[wd * 32u + wm];
                        const uint32_t f = c < nvec ? word >> (d & 31u) : 0u;
                        double z0 = 0.0, z1 = 0.0;
                        if (kern_m == 0u && (f & 3u)) {  // sim_walk :289-296
                            const uint32_t rank = (uint32_t)((pref_m >> (8u * wd)) & 0xFFu) + __popc(word & ((1u << (d & 31u)) - 1u));
                            if (nphi_m <= kTwalkZList) {  // parked by the walker's lane
                                if (f & 1u) z0 = zlist[wm * kTwalkZList + rank];
                                if (f & 2u) z1 = zlist[wm * kTwalkZList + rank + (f & 1u)];
                            } else {  // a walker with many flagged coordinates (pphi close to 1): draw here
                                double u0, u1;
                                if (tp.walk_u_in) {
                                    const double *wu = tp.walk_u_in + (size_t)id_m * p.dim + d;
                                    u0 = __ldg(wu);
                                    u1 = d + 1u < p.dim ? __ldg(wu + 1) : 0.0;
                                } else {
                                    const U4 r = draw(p.seed, hstep, ST_TW_WALK, id_m + p.w_off, c);
                                    u0 = u53(r.x, r.y);
                                    u1 = u53(r.z, r.w);
                                }
                                if (f & 1u) z0 = tp.aw_ratio * (tp.aw * (u0 * u0) + 2.0 * u0 - 1.0);
                                if (f & 2u) z1 = tp.aw_ratio * (tp.aw * (u1 * u1) + 2.0 * u1 - 1.0);
                            }
                        }
                        // Walk: x + (x - xp) z on every coordinate (:298); Traverse: xp + b (xp - x) on flagged ones (:346)
                        const double wx = xm.x + (xm.x - xp.x) * z0, wy = xm.y + (xm.y - xp.y) * z1;
                        const double tx = (f & 1u) ? xp.x + b_m * (xp.x - xm.x) : xm.x;
                        const double ty = (f & 2u) ? xp.y + b_m * (xp.y - xm.y) : xm.y;
                        yk[pr][it].x = kern_m == 0u ? wx : tx;
                        yk[pr][it].y = kern_m == 0u ? wy : ty;
                        // all_different(yp, x) :226-240, x = the partner
                        if (c < nvec) eq = eq || yk[pr][it].x == xp.x || (d + 1u < p.dim && yk[pr][it].y == xp.y);
                    }
                    double pa[Plugin::kAcc];
#pragma unroll
                    for (int a = 0; a < Plugin::kAcc; ++a) pa[a] = 0.0;
#pragma unroll
                    for (int it = 0; it < ITERS; ++it) {
                        const uint32_t c = (uint32_t)it * G + j;
                        const uint32_t d = 2u * c;
                        double nb = 0.0;
                        if constexpr (Plugin::kNeedsNext) {
                            nb = __shfl_down_sync(0xffffffffu, yk[pr][it].x, 1, G);
                            if (ITERS > 1 && it + 1 < ITERS) {
                                const double w0 = __shfl_sync(0xffffffffu, yk[pr][it + 1 < ITERS ? it + 1 : it].x, 0, G);
                                if (j == G - 1) nb = w0;
                            }
                        }
                        if (c < nvec) plug.terms(d, yk[pr][it].x, yk[pr][it].y, nb, d + 1u < p.dim, d + 2u < p.dim, pa);
                    }
#pragma unroll
                    for (int a = 0; a < Plugin::kAcc; ++a) { acc[a][2 * pr] = pa[a]; acc[a][2 * pr + 1] = 0.0; }
                    const uint32_t any_eq = __ballot_sync(0xffffffffu, eq) & (G == 32 ? 0xffffffffu : ((1u << (G & 31)) - 1u) << (grp * G));
                    if (lane == wm && any_eq) my_eq = true;
                }
#pragma unroll
                for (int a = 0; a < Plugin::kAcc; ++a) {
                    double r = transpose_reduce<GB>(acc[a], lane);
#pragma unroll
                    for (int h = GB; h < G; h <<= 1) r += __shfl_xor_sync(0xffffffffu, r, h);
                    // pair k's total sits on the lane of its first member; in half-pass 1 the second member moves
                    if (half) r = __shfl_xor_sync(0xffffffffu, r, 1);
                    if ((int)(j / GB) == blk) tot[a] = r;
                }
                // ---- calc_a :480-501 and the accept test :756-759 for this block's movers; accepted rows leave from
                // the registers of the group's lanes ----
                bool acc_now = false;
                if ((int)(j / GB) == blk && act && my_half == half) {
                    const double lp_new = plug.finish_acc(tot);
                    // calc_a: A = 0 for an empty move, else exp(t); accepted iff u < A (:756-759)
                    const double t = (lp_new - lp_old) * beta + lb;
                    if (nphi == 0u || my_eq) {
                        acc_now = false;  // u < 0.0
                    } else if constexpr (NB > 1) {
                        const double d = swap_guard<double>() * (1.0 + fabs(t));
                        const bool yes = lu < t - d, no = lu > t + d;
                        acc_now = yes;
                        if (__builtin_expect(!(yes | no), 0)) acc_now = u < exp(t);
                    } else {
                        acc_now = u < exp(t);
                    }
                    if (acc_now) {
                        p.lp[w] = lp_new;
                        if (p.dirty) p.dirty[w] = 1;
                        ++n_acc;
                        accept = true;
                    }
                }
                const uint32_t accmask = __ballot_sync(0xffffffffu, acc_now);
#pragma unroll
                for (int pr = 0; pr < PB; ++pr) {
                    const uint32_t wm = grp * G + (uint32_t)(blk * GB + 2 * pr) + half;
                    const uint32_t row = __shfl_sync(0xffffffffu, w, wm);
                    if ((accmask >> wm) & 1u) {
                        double2 *dst = reinterpret_cast<double2 *>(p.ens + (size_t)row * p.pitch);
#pragma unroll
                        for (int it = 0; it < ITERS; ++it) {
                            const uint32_t c = (uint32_t)it * G + j;
                            if (c < nvec) dst[c] = yk[pr][it];
                        }
                    }
                }
                if (again && NB == 1) load_block(xn, w, 0, half + 1u);  // after the stores it must see
            }
        }
        if (p.accepted_out && act && lane_in) p.accepted_out[w] = accept ? 1 : 0;
        if (p.rung_acc) count_rung_accepts(p.rung_acc, act ? w / p.npb : 0u, act, accept, lane);

        if (tile_n >= ntiles) break;
        tile = tile_n;
        w = w_n;
        act = act_n;
        lp_old = lp_n;
    }

    for (int off = 16; off; off >>= 1) n_acc += __shfl_xor_sync(0xffffffffu, n_acc, off);
    if (lane == 0 && n_acc) atomicAdd(p.stats, n_acc);
}

}  // namespace scorus
