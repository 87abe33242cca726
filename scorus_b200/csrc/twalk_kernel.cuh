// twalk_kernel.cuh -- one half-pass of the ensemble / parallel-tempering t-walk,
// src/mcmc/twalk.rs: sample1 :650-762 = propose_move :505-532 (TWalkKernal::random :40-63, sim_walk
// :274-304, sim_b :306-322, sim_traverse :324-353, gen_update_flags :259-272) for one member of every
// pair, the log-density of the proposals, calc_a :463-503 and the accept loop :736-760, in one launch.
// twalk::sample (:586-648) is two such half-passes over the same matching, the second with the roles
// of each pair's members exchanged.  Because the pairs are disjoint and both half-passes share the
// matching, the second half-pass of a pair depends on nothing but the first half-pass of the SAME
// pair: with half = 2 the kernel runs both on a tile before moving on (one launch per t-walk step;
// the rows are re-read for the second half-pass, from L2: they were streamed a few microseconds
// earlier), so HBM sees each row once in and at most once out per step, as in the stretch move.
//
// Same shape as the stretch-move kernel (step_reg_kernel.cuh): a tile is 32 consecutive matching
// positions = 16 pairs (positions 2k, 2k+1), a group of G lanes streams both rows of a pair into
// registers, the proposal of the MOVING member (position parity == half) and its log-density terms
// are computed from registers, the proposed row goes to shared memory and leaves it by a TMA bulk
// store only if accepted.  The partner's row is read, never written.  Launched per half-pass (half = 0
// or 1: the fixed-stream entry point, or SCORUS_TWALK_FUSED=0) every row is read from HBM twice per
// t-walk step; fused (half = 2, the default) once.
//
// Rounding follows the reference operation by operation (-fmad=false): sim_walk computes
// x + (x - xp) * z with z = (aw / (1 + aw)) * (aw * u*u + 2 u - 1) on flagged coordinates and z = 0 on
// the others (:287-301); sim_traverse xp + b * (xp - x) on flagged coordinates (:344-347).
#pragma once
#include "step_reg_kernel.cuh"

namespace scorus {

enum : uint32_t {
    ST_TW_B = 9,     // idx = walker, sub = 0: (r0,r1) -> x, (r2,r3) -> v of sim_b
    ST_TW_WALK = 10  // idx = walker, sub = d/2: sim_walk uniforms of coordinates d, d+1
};

struct TwalkParams {
    StepParams s;               // ensemble, streams, geometry; s.step = this half-pass's counter
    double aw, aw_ratio, at;    // TWalkParams :75-83; aw_ratio = aw / (1 + aw)
    double fw0, fw1;            // cumulative kernel weights (:93-101)
    const uint8_t *kernel_in;   // caller-supplied draws, indexed by walker (or nullptr -> device streams)
    const double *b_in;         // [n]
    const double *walk_u_in;    // [n][dim]
    uint64_t order_step;        // counter whose matching both half-passes share
    uint32_t half;              // the member of each pair that moves (position parity): 0, 1, or 2 = 0 then 1
                                // fused, the second keyed by step counter s.step + 1
    uint32_t l2_keep;           // fused mode, L2 policy of the first read of a row (it is read again ~20 us later):
                                // 1 evict_last, 2 evict_normal, 0 no hint (SCORUS_TWALK_L2)
    uint32_t direct;            // launch twalk_direct_kernel (no shared-memory tile) where the plugin / row width allow
};

// 128-bit streaming load with an explicit L2 eviction policy.  The plain L1::no_allocate load is treated as
// evict_first by L2 (ncu: lts__t_sectors_srcunit_tex_op_read_evict_first_*), right for rows read once and
// wrong for the fused kernel's first read, whose second read then misses.
__device__ __forceinline__ double2 ld_stream_hint(const double2 *p, uint64_t policy)
{
    double2 v;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;"
                 : "=d"(v.x), "=d"(v.y)
                 : "l"(p), "l"(policy));
    return v;
}

#ifndef SCORUS_TW_ROWS
#define SCORUS_TW_ROWS 4
#endif

// sim_b (:306-322) from its two uniforms
__device__ __forceinline__ double twalk_b(double x, double v, double at)
{
    if (x == 0.0) return x;
    return v < (at - 1.0) / (2.0 * at) ? pow(x, 1.0 / (at + 1.0)) : pow(x, 1.0 / (1.0 - at));
}

template <class Plugin, int G, int ITERS>
__global__ void __launch_bounds__(twalk_max_warps<Plugin, G>() * 32, 1) twalk_half_kernel(const TwalkParams tp)
{
    // rows per software-pipeline block: SCORUS_TW_ROWS (default 4 = two pairs) keeps the unrolled pair loop, and
    // with it the per-tile instruction footprint, small enough for the 32 KB L1.5 instruction cache
    constexpr int kWant = SCORUS_TW_ROWS / ITERS < 2 ? 2 : SCORUS_TW_ROWS / ITERS;
    constexpr int GB = G < kWant ? G : kWant, PB = GB / 2, NB = G / GB;
    const StepParams &p = tp.s;

    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t tile_bytes = 32u * p.row_stride;
    const uint32_t warp_bytes = tile_bytes + ((p.flag_words * 128u + 127u) & ~127u);
    unsigned char *ytile = smem + (size_t)warp * warp_bytes;
    uint32_t *flagw = reinterpret_cast<uint32_t *>(ytile + tile_bytes);

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
        const uint64_t pol = h == 0u ? pol_keep : pol_drop;  // fused mode: the rows come back for half-pass 1
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

    grid_dep_wait();  // from here on: what earlier launches wrote (order, ensemble, log-probs)

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
    // tiles start at even positions: lane parity = position parity.  A lane takes part in the half-pass
    // that moves its parity; in fused mode that is every lane, odd lanes one step counter later.
    const uint32_t my_half = lane & 1u;
    const bool lane_in = both || my_half == tp.half;
    const uint64_t my_step = p.step + (both ? my_half : 0u);
    double lp_old = act && lane_in ? __ldg(p.lp + w) : 0.0;
    double2 xn[GB][ITERS];
    load_block(xn, w, 0, h_begin);

    for (;;) {
        const bool mv = act && lane_in;
        // ---- P1: lane = walker: kernel choice, b, accept uniform, flags ----
        uint32_t kern = 0;
        double b = 1.0, u = 1.0;
        if (tp.kernel_in) {
            if (mv) { kern = __ldg(tp.kernel_in + w); b = __ldg(tp.b_in + w); u = __ldg(p.u_in + w); }
        } else if (mv) {
            const U4 r = draw(p.seed, my_step, ST_ZU, w + p.w_off, 0);
            kern = u53(r.x, r.y) * tp.fw1 < tp.fw0 ? 0u : 1u;  // TWalkKernal::random :57-62
            u = u53(r.z, r.w);
            if (kern == 1u) {
                const U4 q = draw(p.seed, my_step, ST_TW_B, w + p.w_off, 0);
                b = twalk_b(u53(q.x, q.y), u53(q.z, q.w), tp.at);
            }
        }
        bulk_wait_read_all();  // the previous tile's accepted rows have left the y tile / flag words
        __syncwarp();
        const uint32_t nphi = gen_flags(p, my_step, w, flagw, lane, mv);
        // calc_a :487-489, second summand: (nphi - 2) ln b for Traverse, nothing for Walk
        const double lb = kern == 1u ? (double)((int)nphi - 2) * log(b) : 0.0;
        const double beta = __ldg(p.beta + (act ? w / p.npb : 0u));
        __syncwarp();

        // sim_walk's z (:289-296) of a walker's flagged coordinates, computed by the walker's own lane (one Philox call
        // per flagged chunk: a handful per walker) and parked in the walker's row of the y tile, where the proposal loop
        // picks them up before it overwrites the row with y.  Doing this in the loop instead costs a Philox call per
        // pair and trip for the whole warp (19 % of the kernel's instructions).
        auto scatter_walk_z = [&](bool mine_now) {
            if (mine_now && kern == 0u) {
                double *row = reinterpret_cast<double *>(ytile + lane * p.row_stride);
                for (uint32_t wd = 0; wd < p.flag_words; ++wd) {
                    uint32_t bits = flagw[wd * 32u + lane];
                    while (bits) {
                        const uint32_t sh = (uint32_t)(__ffs((int)bits) - 1) & ~1u;  // the chunk's even coordinate, within the word
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
                        if (two & 1u) row[d0] = tp.aw_ratio * (tp.aw * (u0 * u0) + 2.0 * u0 - 1.0);
                        if (two & 2u) row[d0 + 1u] = tp.aw_ratio * (tp.aw * (u1 * u1) + 2.0 * u1 - 1.0);
                        bits &= ~(3u << sh);
                    }
                }
            }
        };
        // dense targets rewrite the partner's row during a half-pass (tile evaluation), so their z go in per half-pass
        if constexpr (!Plugin::kNeedsRow) {
            scatter_walk_z(mv);
            __syncwarp();
        }

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

        bool accept = false;  // this lane's decision (set in the half-pass that moves it)
#pragma unroll 1
        for (uint32_t half = h_begin; half < h_end; ++half) {
            // fused second half-pass: the partner is the first half-pass's mover, at its NEW position
            const bool second = both && half == 1u;
            if constexpr (Plugin::kNeedsRow) {
                scatter_walk_z(act && my_half == half);
                __syncwarp();
            }
            double tot[Plugin::kAcc];
            bool my_eq = false;  // all_different (:226-240) failed for the pair this lane moves in
#pragma unroll
            for (int a = 0; a < Plugin::kAcc; ++a) tot[a] = 0.0;

#pragma unroll 1
            for (int blk = 0; blk < NB; ++blk) {
                double2 xc[GB][ITERS];
#pragma unroll
                for (int i = 0; i < GB; ++i)
#pragma unroll
                    for (int it = 0; it < ITERS; ++it) xc[i][it] = xn[i][it];
                if (blk + 1 < NB) load_block(xn, w, blk + 1, half);
                else if (half + 1u < h_end) load_block(xn, w, 0, half + 1u);  // the same tile again (old rows, from L2)
                else if (tile_n < ntiles) load_block(xn, w_n, 0, h_begin);

                double acc[Plugin::kAcc][GB];
#pragma unroll
                for (int pr = 0; pr < PB; ++pr) {
                    const uint32_t wm = grp * G + (uint32_t)(blk * GB + 2 * pr) + half, wp = wm ^ 1u;  // tile rows: mover, partner
                    const uint32_t kern_m = __shfl_sync(0xffffffffu, kern, wm);
                    const double b_m = __shfl_sync(0xffffffffu, b, wm);
                    const bool moved_p = __shfl_sync(0xffffffffu, (uint32_t)accept, wp) != 0u && second;
                    double2 *rowm = reinterpret_cast<double2 *>(ytile + wm * p.row_stride);
                    double2 *rowp = reinterpret_cast<double2 *>(ytile + wp * p.row_stride);
                    double2 y[ITERS];
                    bool eq = false;
#pragma unroll
                    for (int it = 0; it < ITERS; ++it) {
                        const uint32_t c = (uint32_t)it * G + j;
                        const uint32_t d = 2u * c;
                        const double2 xm = xc[2 * pr][it];
                        double2 xp = xc[2 * pr + 1][it];
                        // its accepted proposal is still in the tile (and on its way to global memory)
                        if (moved_p && c < nvec) xp = rowp[c];
                        const uint32_t wd = (d >> 5) < p.flag_words ? (d >> 5) : 0u;
                        const uint32_t f = c < nvec ? flagw[wd * 32u + wm] >> (d & 31u) : 0u;
                        double z0 = 0.0, z1 = 0.0;
                        if (kern_m == 0u && (f & 3u)) {  // sim_walk :289-296: z of the flagged coordinates, parked in the row
                            const double2 zz = rowm[c];
                            if (f & 1u) z0 = zz.x;
                            if (f & 2u) z1 = zz.y;
                        }
                        // Walk: x + (x - xp) z on every coordinate (:298); Traverse: xp + b (xp - x) on flagged ones (:346)
                        const double wx = xm.x + (xm.x - xp.x) * z0, wy = xm.y + (xm.y - xp.y) * z1;
                        const double tx = (f & 1u) ? xp.x + b_m * (xp.x - xm.x) : xm.x;
                        const double ty = (f & 2u) ? xp.y + b_m * (xp.y - xm.y) : xm.y;
                        y[it].x = kern_m == 0u ? wx : tx;
                        y[it].y = kern_m == 0u ? wy : ty;
                        if (c < nvec) {
                            rowm[c] = y[it];
                            if constexpr (Plugin::kNeedsRow) {
                                if (!second) rowp[c] = xp;  // the tile evaluation reads all 32 rows
                            }
                            // all_different(yp, x) :226-240, x = the partner
                            eq = eq || y[it].x == xp.x || (d + 1u < p.dim && y[it].y == xp.y);
                        }
                    }
                    double pa[Plugin::kAcc];
#pragma unroll
                    for (int a = 0; a < Plugin::kAcc; ++a) pa[a] = 0.0;
                    if constexpr (!Plugin::kNeedsRow) {
#pragma unroll
                        for (int it = 0; it < ITERS; ++it) {
                            const uint32_t c = (uint32_t)it * G + j;
                            const uint32_t d = 2u * c;
                            double nb = 0.0;
                            if constexpr (Plugin::kNeedsNext) {
                                nb = __shfl_down_sync(0xffffffffu, y[it].x, 1, G);
                                if (ITERS > 1 && it + 1 < ITERS) {
                                    const double w0 = __shfl_sync(0xffffffffu, y[it + 1 < ITERS ? it + 1 : it].x, 0, G);
                                    if (j == G - 1) nb = w0;
                                }
                            }
                            if (c < nvec) plug.terms(d, y[it].x, y[it].y, nb, d + 1u < p.dim, d + 2u < p.dim, pa);
                        }
                    }
#pragma unroll
                    for (int a = 0; a < Plugin::kAcc; ++a) { acc[a][2 * pr] = pa[a]; acc[a][2 * pr + 1] = 0.0; }
                    // the G lanes of a group cover the mover's row, and the mover's own lane is one of them
                    const uint32_t any_eq = __ballot_sync(0xffffffffu, eq) & (G == 32 ? 0xffffffffu : ((1u << (G & 31)) - 1u) << (grp * G));
                    if (lane == wm && any_eq) my_eq = true;
                }
                if constexpr (!Plugin::kNeedsRow) {
#pragma unroll
                    for (int a = 0; a < Plugin::kAcc; ++a) {
                        double r = transpose_reduce<GB>(acc[a], lane);
#pragma unroll
                        for (int h = GB; h < G; h <<= 1) r += __shfl_xor_sync(0xffffffffu, r, h);
                        if ((int)(j / GB) == blk) tot[a] = r;
                    }
                }
            }
            // the totals of pair k sit on the lane of its first member; in half-pass 1 the second member moves
            if (half) {
#pragma unroll
                for (int a = 0; a < Plugin::kAcc; ++a) tot[a] = __shfl_xor_sync(0xffffffffu, tot[a], 1);
            }
            fence_proxy_async();
            __syncwarp();

            // ---- P5: lane = walker: calc_a :480-501 and the accept test :756-759, for this half-pass's movers ----
            unsigned char *own_b = ytile + lane * p.row_stride;
            double lp_new;
            if constexpr (Plugin::kNeedsRow) {
                lp_new = plug.tile_logprob(ytile, p.row_stride, lane);
            } else {
                lp_new = plug.finish_acc(tot);
            }
            const bool mine = act && my_half == half;
            const double a_acc = (nphi == 0u || my_eq) ? 0.0 : exp((lp_new - lp_old) * beta + lb);
            const bool acc_now = mine && (u < a_acc);
            if (acc_now) {
                bulk_store(p.ens + (size_t)w * p.pitch, smem_u32(own_b), p.row_bytes);
                p.lp[w] = lp_new;
                if (p.dirty) p.dirty[w] = 1;
                ++n_acc;
                accept = true;
            }
            bulk_commit();
            if (p.accepted_out && mine) p.accepted_out[w] = acc_now ? 1 : 0;
            if (p.rung_acc) count_rung_accepts(p.rung_acc, act ? w / p.npb : 0u, act, acc_now, lane);
        }

        if (tile_n >= ntiles) break;
        tile = tile_n;
        w = w_n;
        act = act_n;
        lp_old = lp_n;
    }

    bulk_wait_all();
    for (int off = 16; off; off >>= 1) n_acc += __shfl_xor_sync(0xffffffffu, n_acc, off);
    if (lane == 0 && n_acc) atomicAdd(p.stats, n_acc);
}

}  // namespace scorus
