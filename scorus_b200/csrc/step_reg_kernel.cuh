// step_reg_kernel.cuh -- K1: the fused stretch-move step, register-streaming version (default
// for dim <= 256).
//
// One launch does everything src/mcmc/ensemble_sample.rs:135-189 does per step: partner pick
// (:135-148), z draw (:149 -> utils.rs:33-45), update flags (:150-153 -> :38-54), proposal
// (:155-159 -> :57-75 -> utils.rs:47-56), log-density (:161) and Metropolis accept (:164-189).
//
// Why this shape (profiles/r1_notes.md has the measurements): staging the OLD rows in shared memory
// (step_kernel.cuh, step_coop_kernel.cuh) costs two tiles of shared memory per warp, which caps an SM at 6
// warps for D = 64 -- the kernel ran at 43 % of the HBM roofline, issue-bound.  Here:
//   * a group of G lanes owns a pair; lane j loads double2 chunk j (+ k*G) of BOTH rows straight into
//     registers with 128-bit coalesced loads (a row is one contiguous 16*G-byte burst), one block of 8 rows
//     ahead of the arithmetic (software pipeline, 4 KB in flight per warp);
//   * the proposals y = x1*z + x2*(1-z) (both walkers of the pair, from the old values) and their log-density
//     terms are computed from registers; Rosenbrock's neighbour coordinate comes from the next lane by shuffle;
//   * per-lane partial sums are combined by a transposed shuffle tree, which hands each row's total to the lane
//     whose walker it is, so the scalar tail (exp, accept) runs with lane = walker;
//   * direct mode (kDirect, the default): the tail runs per block of 8 rows and accepted rows go from the
//     group's registers straight to global memory -- shared memory holds the flag words only and the block size
//     is bound by registers (16 warps per SM: 0.64-0.66 ms per 2^22 x 64-D step);
//   * tile mode (the dense target, whose tensor-core evaluation reads the tile; rows of 9..16 double2): the
//     PROPOSED rows wait in a shared-memory tile, the tail runs once per tile and accepted rows leave through
//     TMA bulk stores (cp.async.bulk shared->global).  This was the only mode at first (11 warps per SM at
//     D = 64, 0.726 ms).
// HBM traffic stays below the algorithmic 16*D+16 bytes per walker-step (rejected rows are not written).
// Positions are bit-identical to the reference's rounding sequence; log-probs differ from a left-to-right
// sum by a few ulp of the sum of |terms| (fixed tree, deterministic).
#pragma once
#include "launch.cuh"            // SCORUS_REG_ROWS, SCORUS_REG_WARPS
#include "step_coop_kernel.cuh"  // transpose_reduce
#include "tile_common.cuh"

namespace scorus {

template <int G, int ITERS>
struct RegGeom {
    static constexpr int kRows = ((G == 32 && ITERS <= 2) || (G == 16 && SCORUS_G16_WIDE)) ? SCORUS_REG_ROWS_WIDE : SCORUS_REG_ROWS;
    static constexpr int kRowsPerBlock = (G < kRows / ITERS) ? G : (kRows / ITERS < 2 ? 2 : kRows / ITERS);  // GB
    static constexpr int kPairs = kRowsPerBlock / 2;
    static constexpr int kBlocks = G / kRowsPerBlock;
};

// kPeer: one-sided multi-GPU mode -- this rank owns the pairs of its list and reads / writes BOTH rows wherever they
// live (p.peer_ens / p.peer_lp: CUDA-IPC mappings of the peers' buffers, loads and stores travel over NVLink).  A pair
// is owned by exactly one rank and a walker is in exactly one pair, so within a step nobody else touches either row:
// one barrier between steps is all the synchronisation there is (api_shard.cu).  A template parameter because even
// uniform branches on it cost 12 % in the single-GPU kernel.
//   kPeer == 1: both rows of a pair stream through registers as in the single-GPU kernel (tile mode: the dense target,
//               rows of 9..16 double2).
//   kPeer == 2 (direct mode): the PARTNER rows -- the ones that may live on another GPU -- of a whole tile are fetched
//               a tile ahead by TMA bulk copies (cp.async.bulk, one instruction per row, completion on an mbarrier per
//               block slot) into a per-warp shared-memory stage, refilled slot by slot as the blocks consume it.  An
//               NVLink read takes several times an HBM read; with the register pipeline (one block of 4 rows ahead) a
//               warp sat out that latency once per block and the step ran at 65 % of the measured NVLink read
//               bandwidth, local and remote traffic in turn rather than together (2 GPUs: 0.65 ms against 0.31 ms
//               shard-local, profiles/r2_notes.md).  The owner's own rows keep the register pipeline.
// Warps per block (= per SM): 20 under a 96-register cap for rows wider than 16 double2 (launch.cuh), 16 under 128
// registers otherwise; the dense plugin keeps 8 warps and up to 255 registers for its DMMA accumulators.
template <class Plugin, int G, int ITERS>
constexpr int reg_kernel_max_warps()
{
    return Plugin::kMaxWarps == 14 ? (((G == 32 && ITERS <= 2) || (G == 16 && SCORUS_G16_WIDE)) ? SCORUS_REG_WARPS_WIDE : SCORUS_REG_WARPS) : Plugin::kMaxWarps;
}

// the t-walk kernels (twalk_kernel.cuh, twalk_direct_kernel.cuh) keep 16 warps: they need up to 128 registers
template <class Plugin, int G>
constexpr int twalk_max_warps()
{
    return Plugin::kMaxWarps == 14 ? 16 : Plugin::kMaxWarps;
}

template <class Plugin, int G>
__host__ __device__ constexpr bool reg_kernel_direct()
{
    return !Plugin::kNeedsRow && (G != 16 || SCORUS_G16_DIRECT);
}

// kRemap: the step after a ladder swap moves the rows (StepParams::remap); single GPU, direct mode only
template <class Plugin, int G, int ITERS, int kPeer = 0, bool kRemap = false>
__global__ void __launch_bounds__(reg_kernel_max_warps<Plugin, G, ITERS>() * 32, 1) step_reg_kernel(const StepParams p)
{
    using Geo = RegGeom<G, ITERS>;
    constexpr int GB = Geo::kRowsPerBlock, PB = Geo::kPairs, NB = Geo::kBlocks;
    constexpr bool kStageB = kPeer == 2;                 // partner rows through the shared-memory stage
    constexpr int RL = kStageB ? 2 : 1;                  // register pipeline: every row, or the even (own) rows only
    constexpr int NR = GB / RL;                          // rows per block held in registers
    constexpr uint32_t kSlotRows = (32 / G) * (GB / 2);  // partner rows per block slot of the stage
    // direct: proposals never touch shared memory (only the flag words do).  The dense plugin keeps the y tile, which
    // its tensor-core evaluation reads; so do rows of 9..16 double2 (G = 16): their tile is small enough for 16
    // warps anyway, and deciding per block of 8 rows would run the scalar tail twice per tile at half occupancy
    // (measured at 32-D: 0.129 ms with the tile, 0.148 ms direct)
    constexpr bool kDirect = reg_kernel_direct<Plugin, G>();

    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warp = threadIdx.x >> 5;
    static_assert(!kStageB || kDirect, "the partner stage is built for the direct mode");
    static_assert(!kRemap || (kDirect && kPeer == 0), "the row-moving step is built for the single-GPU direct mode");
    const uint32_t tile_bytes = kDirect ? 0u : 32u * p.row_stride;
    const uint32_t stage_bytes = kStageB ? 16u * p.row_stride + 128u : 0u;  // 16 partner rows + NB mbarriers
    const uint32_t warp_bytes = tile_bytes + stage_bytes + ((p.flag_words * 128u + 127u) & ~127u);
    unsigned char *ytile = smem + (size_t)warp * warp_bytes;
    [[maybe_unused]] unsigned char *bstage = ytile + tile_bytes;
    [[maybe_unused]] const uint32_t bars = smem_u32(bstage + 16u * p.row_stride);  // NB x 8 bytes
    uint32_t *flagw = reinterpret_cast<uint32_t *>(ytile + tile_bytes + stage_bytes);

    const uint32_t gwarp = blockIdx.x * p.warps + warp;
    const uint32_t nwarps = gridDim.x * p.warps;
    const uint32_t grp = lane / G, j = lane % G;
    const uint32_t nvec = p.pitch >> 1;
    const bool all_flags = p.flag_kind == 1u;
    unsigned long long n_acc = 0;

    grid_dep_launch();  // the next step's blocks may be scheduled as soon as an SM frees up

    // One-block ensembles: exact uniform shuffle of every rung right here (rank of 64-bit Philox keys,
    // the same keys and tie-break as exact_order_kernel), saving a launch per step.
    const uint32_t *sm_order = nullptr;
    if (p.fused_order) {
        unsigned char *extra = smem + (size_t)p.warps * warp_bytes;
        uint64_t *keys = reinterpret_cast<uint64_t *>(extra);
        uint32_t *ord = reinterpret_cast<uint32_t *>(extra + (size_t)p.n * sizeof(uint64_t));
        for (uint32_t i = threadIdx.x; i < p.n; i += blockDim.x) {
            const U4 o = draw(p.seed, p.step, ST_PERM_KEY, p.w_off + i, 0);
            keys[i] = ((uint64_t)o.y << 32) | o.x;
        }
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < p.n; i += blockDim.x) {
            const uint32_t r = i / p.npb, base = r * p.npb;
            const uint64_t ki = keys[i];
            uint32_t rank = 0;
            for (uint32_t jj = 0; jj < p.npb; ++jj) {
                const uint64_t kj = keys[base + jj];
                rank += (kj < ki) || (kj == ki && base + jj < i);
            }
            ord[base + rank] = i;
        }
        __syncthreads();
        sm_order = ord;
    }

    Plugin plug;
    plug.init(p.pparams, p.nparams, p.dim);

    // walker at matching position `pos` (keys of the pairing bijection cached per rung)
    uint32_t key_rung = 0xFFFFFFFFu;
    BijKeys K;
    auto walker_of = [&](uint32_t pos) -> uint32_t {
        if (sm_order) return sm_order[pos];
        if (p.order) return __ldg(p.order + pos);
        if (p.debug_identity) return pos;
        const uint32_t blk = pos / p.mblk;  // matching block (the whole rung unless the matching is blocked)
        if (blk != key_rung) {
            const uint32_t r = pos / p.npb;
            K = bij_keys(p.seed, p.step, ST_PERM, r + p.rung_off, 2u * (blk - r * (p.npb / p.mblk) + p.mblk_off));
            key_rung = blk;
        }
        return blk * p.mblk + bij(pos - blk * p.mblk, p.mblk, p.mblk_bits, K);
    };
    // peer mode: home of a walker as (rank << 28 | row), computed once per tile by the walker's lane
    auto home_of = [&](uint32_t wid) -> uint32_t {
        const uint32_t q = wid / p.peer_n;
        return (q << 28) | (wid - q * p.peer_n);
    };

    // loads block `blk` of a tile whose lane = walker ids are `wt` (peer mode: lane = walker homes): rows
    // wl = grp*G + blk*GB + i
    // (kRemap: lane = slot the walker's row is in now)
    auto load_block = [&](double2 (&x)[NR][ITERS], uint32_t wt, int blk) {
#pragma unroll
        for (int i = 0; i < NR; ++i) {
            const uint32_t wl = grp * G + (uint32_t)(blk * GB + i * RL);
            const uint32_t row = __shfl_sync(0xffffffffu, wt, wl);
            const double *base;
            if constexpr (kPeer != 0) base = p.peer_ens[row >> 28] + (size_t)(row & 0x0FFFFFFFu) * p.pitch;  // maybe over NVLink
            else base = p.src_ens + (size_t)row * p.pitch;
            const double2 *src = reinterpret_cast<const double2 *>(base);
#pragma unroll
            for (int it = 0; it < ITERS; ++it) {
                // lanes beyond the row re-read its last chunk (never used): an unconditional load keeps
                // the result out of any select, so nothing waits on it before the arithmetic does
                const uint32_t c = (uint32_t)it * G + j;
                x[i][it] = ld_stream(src + (c < nvec ? c : nvec - 1u));  // streamed: read exactly once
            }
        }
    };

    // partner stage: slot `blk` holds the partner (odd-lane) rows of block blk of a tile; every odd lane fetches its own
    // walker's row with one bulk copy, lane 0 arms the slot's mbarrier with the byte count
    [[maybe_unused]] auto stage_issue = [&](uint32_t hm, int blk) {
        if constexpr (kStageB) {
            const uint32_t bar = bars + 8u * (uint32_t)blk;
            if (lane == 0) mbar_expect_tx(bar, kSlotRows * p.row_bytes);
            if ((lane & 1u) && (int)((lane % G) / GB) == blk)
                bulk_load(smem_u32(bstage + (lane >> 1) * p.row_stride),
                          p.peer_ens[hm >> 28] + (size_t)(hm & 0x0FFFFFFFu) * p.pitch, p.row_bytes, bar);
        }
    };
    if constexpr (kStageB) {
        if (lane == 0) {
#pragma unroll
            for (int b = 0; b < NB; ++b) mbar_init(bars + 8u * b, 1);
            fence_mbar_init();
        }
        __syncwarp();
    }
    [[maybe_unused]] uint32_t stage_phase = 0;

    // Everything above touches launch parameters and constants only.  From here on the kernel reads what
    // earlier launches wrote (order / pair lists, the ensemble, the log-probs): wait for them.
    grid_dep_wait();

    // matching positions to process: N, or (global-matching mode) the length of the local pair list
    const uint32_t npos = p.n_dev ? __ldg(p.n_dev) : p.n;
    const uint32_t ntiles = (npos + 31u) >> 5;
    uint32_t tile = gwarp;
    if (tile >= ntiles) return;
    uint32_t w;
    bool act;
    {
        const uint32_t pos = tile * 32u + lane;
        act = pos < npos;
        w = act ? walker_of(pos) : p.w_lo;
    }
    [[maybe_unused]] uint32_t home = 0, home_n = 0;  // peer mode: where my walker (of this / the next tile) lives
    if constexpr (kPeer != 0) home = home_of(w);
    // old log-prob: from the (gathered) array, or in one-sided mode from the walker's home rank
    [[maybe_unused]] uint32_t rsrc = 0, rsrc_n = 0;  // kRemap: where my walker's row (of this / the next tile) is now
    auto load_lp = [&](uint32_t wid, uint32_t hm, bool a, uint32_t &row_now) -> double {
        if constexpr (kRemap) {
            SwapOut o;
            o.lp = 0.0; o.src = wid; o.pad = 0u;
            if (a) {
                if (p.remap) {  // after a ladder swap: one 16-byte read-only load
                    const int4 v = __ldg(reinterpret_cast<const int4 *>(p.remap) + wid);
                    static_assert(sizeof(SwapOut) == sizeof(int4), "SwapOut is one 16-byte record");
                    memcpy(&o, &v, sizeof(o));
                } else {        // host-buffer call: the row is where it belongs, in the caller's buffer
                    o.lp = __ldg(p.src_lp + wid);
                }
            }
            row_now = o.src;
            return o.lp;
        } else if constexpr (kPeer != 0) {
            return a ? ld_peer(p.peer_lp[hm >> 28] + (hm & 0x0FFFFFFFu)) : 0.0;
        } else {
            return a ? __ldg(p.src_lp + wid) : 0.0;
        }
    };
    double lp_old = load_lp(w, home, act, rsrc);
    if constexpr (kStageB) {
#pragma unroll
        for (int b = 0; b < NB; ++b) stage_issue(home, b);  // the first tile's partner rows, all slots
    }
    double2 xn[NR][ITERS];
    load_block(xn, kPeer != 0 ? home : (kRemap ? rsrc : w), 0);

    for (;;) {
        // ---- P1: lane = walker: z, u, flags, (nphi-1) ln z ----
        double z, u;
        if (p.z_in) {
            z = act ? __ldg(p.z_in + w) : 1.0;
            u = act ? __ldg(p.u_in + w) : 1.0;
        } else {
            const U4 r = draw(p.seed, p.step, ST_ZU, w + p.w_off, 0);
            z = z_from_uniform(to_unit<double>(r.x, r.y), p.inv_sqrt_a, p.z_range);
            u = to_unit<double>(r.z, r.w);
        }
        // Direct mode decides per block of GB rows, with GB of the warp's 32 lanes active: an exp there runs NB times per
        // tile at GB/32 efficiency (9 % of the instructions and 14 % of the stall samples of the 64-D step,
        // profiles/r2b_c5_step_lines.txt).  ln u is taken HERE instead, once per tile with every lane busy, and the
        // decision u < exp(t) becomes ln u < t wherever the two are not within the guard band of each other -- far
        // outside the rounding of log and exp, so it is exactly the literal comparison's answer (the ladder swap's
        // argument, swap_kernel.cuh); inside the band, for u = 0 and for a t that is not finite the lanes take the exp.
        [[maybe_unused]] double lu = 0.0;
        if constexpr (kDirect && NB > 1) lu = u > 0.0 ? log(u) : real_nan<double>();
        // the previous tile's accepted rows must have left the y tile before reuse
        if constexpr (!kDirect) bulk_wait_read_all();
        __syncwarp();
        uint32_t nphi = p.dim;
        if (!all_flags) nphi = gen_flags(p, w, flagw, lane, act);
        const double lz = (double)(nphi - 1u) * log(z);  // ensemble_sample.rs:184, first summand
        const double beta = __ldg(p.beta + (act ? w / p.npb : 0u));
        __syncwarp();  // flag words visible to the whole warp

        // ids of the next tile (its first block is prefetched during this tile's last block): static
        // stride, or -- single-GPU mode -- the next unclaimed tile, which trims the last, ragged round
        uint32_t tile_n = tile + nwarps;
        if (p.tile_counter) {
            unsigned long long t = 0;
            if (lane == 0) t = atomicAdd(p.tile_counter, 1ull) - p.tile_base;
            t = __shfl_sync(0xffffffffu, t, 0);
            tile_n = t + nwarps > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)(t + nwarps);
        }
        uint32_t w_n = p.w_lo;
        bool act_n = false;
        double lp_n = 0.0;
        if (tile_n < ntiles) {
            const uint32_t pos = tile_n * 32u + lane;
            act_n = pos < npos;
            w_n = act_n ? walker_of(pos) : p.w_lo;
        } else {
            w_n = p.w_lo;
        }
        if constexpr (kPeer != 0) home_n = home_of(w_n);
        lp_n = load_lp(w_n, home_n, act_n, rsrc_n);
        if constexpr (kPeer == 0) {
            if (p.l2_prefetch && act_n) {  // my next-tile row -> L2, a whole tile ahead of its loads
                const char *r = reinterpret_cast<const char *>(p.src_ens + (size_t)(kRemap ? rsrc_n : w_n) * p.pitch);
                for (uint32_t off = 0; off < p.row_bytes; off += 128u) prefetch_l2(r + off);
                prefetch_l2(r + p.row_bytes - 1u);
            }
        }

        double tot[Plugin::kAcc];
#pragma unroll
        for (int a = 0; a < Plugin::kAcc; ++a) tot[a] = 0.0;
        // global-matching mode: walkers outside [w_lo, w_hi) belong to another rank, which updates them; peer mode: this
        // rank owns the pair and updates both walkers at their homes
        const bool mine = kPeer != 0 ? act : (act && w >= p.w_lo && w < p.w_hi);
        // where my walker's results go: local row, or (peer mode) its home rank's buffers
        double *const my_lp = kPeer != 0 ? p.peer_lp[home >> 28] + (home & 0x0FFFFFFFu) : p.lp + (w - p.w_lo);
        bool accept = false;

#pragma unroll 1
        for (int blk = 0; blk < NB; ++blk) {
            double2 xc[NR][ITERS];
#pragma unroll
            for (int i = 0; i < NR; ++i)
#pragma unroll
                for (int it = 0; it < ITERS; ++it) xc[i][it] = xn[i][it];
            if (blk + 1 < NB) load_block(xn, kPeer != 0 ? home : (kRemap ? rsrc : w), blk + 1);
            else if (tile_n < ntiles) load_block(xn, kPeer != 0 ? home_n : (kRemap ? rsrc_n : w_n), 0);
            if constexpr (kStageB) mbar_wait(bars + 8u * (uint32_t)blk, stage_phase);  // this block's partner rows have landed

            double acc[Plugin::kAcc][GB];
            [[maybe_unused]] double2 yk[kDirect ? GB : 1][ITERS];  // direct path: the block's proposals stay in registers
#pragma unroll
            for (int pr = 0; pr < PB; ++pr) {
                const uint32_t wa = grp * G + (uint32_t)(blk * GB + 2 * pr), wb = wa + 1u;
                const double za = __shfl_sync(0xffffffffu, z, wa), zb = __shfl_sync(0xffffffffu, z, wb);
                const double omza = 1.0 - za, omzb = 1.0 - zb;
                double2 *rowa = reinterpret_cast<double2 *>(ytile + wa * p.row_stride);
                double2 *rowb = reinterpret_cast<double2 *>(ytile + wb * p.row_stride);
                double2 ya[ITERS], yb[ITERS];
#pragma unroll
                for (int it = 0; it < ITERS; ++it) {
                    const uint32_t c = (uint32_t)it * G + j;
                    const uint32_t d = 2u * c;
                    double2 xa, xb;
                    if constexpr (kStageB) {
                        xa = xc[pr][it];
                        xb = *reinterpret_cast<const double2 *>(bstage + (wb >> 1) * p.row_stride + (c < nvec ? c : nvec - 1u) * 16u);
                    } else {
                        xa = xc[2 * pr][it];
                        xb = xc[2 * pr + 1][it];
                    }
                    uint32_t fa = 3u, fb = 3u;
                    if (!all_flags) {
                        const uint32_t wd = (d >> 5) < p.flag_words ? (d >> 5) : 0u;
                        fa = flagw[wd * 32u + wa] >> (d & 31u);
                        fb = flagw[wd * 32u + wb] >> (d & 31u);
                    }
                    // scale_vec, src/mcmc/utils.rs:55: (x1*z) + (x2*(1-z)), three roundings; then
                    // propose_move, ensemble_sample.rs:69-71: coordinates whose flag is clear stay
                    { double t1 = xa.x * za, t2 = xb.x * omza; ya[it].x = (fa & 1u) ? t1 + t2 : xa.x; }
                    { double t1 = xa.y * za, t2 = xb.y * omza; ya[it].y = (fa & 2u) ? t1 + t2 : xa.y; }
                    { double t1 = xb.x * zb, t2 = xa.x * omzb; yb[it].x = (fb & 1u) ? t1 + t2 : xb.x; }
                    { double t1 = xb.y * zb, t2 = xa.y * omzb; yb[it].y = (fb & 2u) ? t1 + t2 : xb.y; }
                    if constexpr (kDirect) {
                        yk[2 * pr][it] = ya[it];
                        yk[2 * pr + 1][it] = yb[it];
                    } else if (c < nvec) {
                        rowa[c] = ya[it];
                        rowb[c] = yb[it];
                    }
                }
                // log-density terms from registers; coordinate d+2 lives in the next lane
                double pa[Plugin::kAcc], pb[Plugin::kAcc];
#pragma unroll
                for (int a = 0; a < Plugin::kAcc; ++a) { pa[a] = 0.0; pb[a] = 0.0; }
                if constexpr (!Plugin::kNeedsRow) {
#pragma unroll
                    for (int it = 0; it < ITERS; ++it) {
                        const uint32_t c = (uint32_t)it * G + j;
                        const uint32_t d = 2u * c;
                        double nba = 0.0, nbb = 0.0;
                        if constexpr (Plugin::kNeedsNext) {
                            nba = __shfl_down_sync(0xffffffffu, ya[it].x, 1, G);
                            nbb = __shfl_down_sync(0xffffffffu, yb[it].x, 1, G);
                            if (ITERS > 1 && it + 1 < ITERS) {  // last lane of the group: next trip, lane 0
                                const double wa0 = __shfl_sync(0xffffffffu, ya[it + 1 < ITERS ? it + 1 : it].x, 0, G);
                                const double wb0 = __shfl_sync(0xffffffffu, yb[it + 1 < ITERS ? it + 1 : it].x, 0, G);
                                if (j == G - 1) { nba = wa0; nbb = wb0; }
                            }
                        }
                        if (c < nvec && d < p.dim) {  // (f32 rows are padded to 4 coordinates: a chunk can be all padding)
                            const bool has1 = d + 1u < p.dim, hasnb = d + 2u < p.dim;
                            plug.terms(d, ya[it].x, ya[it].y, nba, has1, hasnb, pa);
                            plug.terms(d, yb[it].x, yb[it].y, nbb, has1, hasnb, pb);
                        }
                    }
                }
#pragma unroll
                for (int a = 0; a < Plugin::kAcc; ++a) { acc[a][2 * pr] = pa[a]; acc[a][2 * pr + 1] = pb[a]; }
            }
            if constexpr (kStageB) {
                // every lane has the slot's rows in registers: refill it with the next tile's partner rows
                __syncwarp();
                if (tile_n < ntiles) stage_issue(home_n, blk);
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
            if constexpr (kDirect) {
                // ---- P5, per block: the lanes whose walkers are this block's rows decide (Metropolis accept,
                // ensemble_sample.rs:181-188); accepted rows go from the registers of the group's lanes straight to
                // global memory, one coalesced 16 G-byte burst per row and trip -- no shared-memory tile, so the
                // block size is bound by registers only (16 warps per SM instead of 11 at 64-D)
                bool acc_now = false;
                if ((int)(j / GB) == blk) {
                    const double lp_new = plug.finish_acc(tot);
                    const double t = lz + (lp_new - lp_old) * beta;  // ensemble_sample.rs:184: ln q
                    bool below;                                      // u < q = exp(t), :185
                    if constexpr (NB > 1) {
                        const double d = swap_guard<double>() * (1.0 + fabs(t));
                        const bool yes = lu < t - d, no = lu > t + d;
                        below = yes;
                        if (__builtin_expect(!(yes | no), 0)) below = u < exp(t);
                    } else {
                        below = u < exp(t);
                    }
                    acc_now = mine && below;
                    if constexpr (kRemap) {
                        if (mine && !acc_now) *my_lp = lp_old;  // the swap's log-prob reaches the array here
                    }
                    if (acc_now) {
                        *my_lp = lp_new;
                        if constexpr (kRemap) {
                            if (p.host_lp) p.host_lp[w] = lp_new;  // host-buffer call: straight back into the caller's array
                        }
                        if (p.dirty) {  // host-buffer call: this walker's row has to travel back (one-sided mode: mark it on its home GPU)
                            if constexpr (kPeer != 0) reinterpret_cast<uint8_t *>(p.peer_lp[home >> 28] + p.peer_n + 32)[home & 0x0FFFFFFFu] = 1;
                            else p.dirty[w - p.w_lo] = 1;
                        }
                        ++n_acc;
                        accept = true;
                    }
                }
                const uint32_t accmask = __ballot_sync(0xffffffffu, acc_now);
                // kRemap: every row of mine goes to its slot in the second buffer, accepted or not
                [[maybe_unused]] const uint32_t minemask = kRemap ? __ballot_sync(0xffffffffu, mine && (int)(j / GB) == blk) : 0u;
#pragma unroll
                for (int i = 0; i < GB; ++i) {
                    const uint32_t rl = grp * G + (uint32_t)(blk * GB + i);
                    const uint32_t row = __shfl_sync(0xffffffffu, kPeer != 0 ? home : w, rl);
                    if constexpr (kRemap) {
                        // Rows narrower than a warp (several rows per lane group and block): a rejected row is fetched again
                        // -- from L2, it was read a moment ago -- rather than kept in registers across the log-density
                        // (10-D: 272 bytes of spills and +12 us per step otherwise).  Full-width rows (the shape that also
                        // streams from host memory, where a second read would cross PCIe again) keep theirs.
                        constexpr bool kReload = G < 32;
                        [[maybe_unused]] const uint32_t row_src = kReload ? __shfl_sync(0xffffffffu, rsrc, rl) : 0u;
                        if ((accmask >> rl) & 1u) {
                            if (p.host_ens) {  // host-buffer call: the accepted row goes back over PCIe now, while others still arrive
                                double2 *hdst = reinterpret_cast<double2 *>(p.host_ens + (size_t)row * p.pitch);
#pragma unroll
                                for (int it = 0; it < ITERS; ++it) {
                                    const uint32_t c = (uint32_t)it * G + j;
                                    if (c < nvec) hdst[c] = yk[i][it];
                                }
                            }
                        } else if constexpr (kReload) {
                            if ((minemask >> rl) & 1u) {
                                const double2 *srcp = reinterpret_cast<const double2 *>(p.src_ens + (size_t)row_src * p.pitch);
#pragma unroll
                                for (int it = 0; it < ITERS; ++it) {
                                    const uint32_t c = (uint32_t)it * G + j;
                                    yk[i][it] = ld_stream(srcp + (c < nvec ? c : nvec - 1u));
                                }
                            }
                        } else {
#pragma unroll
                            for (int it = 0; it < ITERS; ++it) yk[i][it] = xc[i][it];
                        }
                    }
                    if (((kRemap ? minemask : accmask) >> rl) & 1u) {
                        double *rowp;
                        if constexpr (kPeer != 0) rowp = p.peer_ens[row >> 28] + (size_t)(row & 0x0FFFFFFFu) * p.pitch;
                        else rowp = p.ens + (size_t)(row - p.w_lo) * p.pitch;
                        double2 *dst = reinterpret_cast<double2 *>(rowp);
#pragma unroll
                        for (int it = 0; it < ITERS; ++it) {
                            const uint32_t c = (uint32_t)it * G + j;
                            if (c < nvec) __stcs(dst + c, yk[i][it]);  // written once, next read a whole step away
                        }
                    }
                }
            }
        }
        if constexpr (!kDirect) {
            fence_proxy_async();  // my writes of the proposed rows -> visible to the TMA stores below
            __syncwarp();

            // ---- P5: lane = walker: Metropolis accept, ensemble_sample.rs:181-188 ----
            unsigned char *own_b = ytile + lane * p.row_stride;
            double lp_new;
            if constexpr (Plugin::kNeedsRow) {
                // K3: dense targets evaluate the whole tile of proposed rows on the fp64 tensor cores
                lp_new = plug.tile_logprob(ytile, p.row_stride, lane);
            } else {
                lp_new = plug.finish_acc(tot);
            }
            const double q = exp(lz + (lp_new - lp_old) * beta);
            accept = mine && (u < q);
            if (accept) {
                double *rowp;
                if constexpr (kPeer != 0) rowp = p.peer_ens[home >> 28] + (size_t)(home & 0x0FFFFFFFu) * p.pitch;
                else rowp = p.ens + (size_t)(w - p.w_lo) * p.pitch;
                bulk_store(rowp, smem_u32(own_b), p.row_bytes);
                *my_lp = lp_new;
                if (p.dirty) {  // host-buffer call: this walker's row has to travel back (one-sided mode: mark it on its home GPU)
                            if constexpr (kPeer != 0) reinterpret_cast<uint8_t *>(p.peer_lp[home >> 28] + p.peer_n + 32)[home & 0x0FFFFFFFu] = 1;
                            else p.dirty[w - p.w_lo] = 1;
                        }
                ++n_acc;
            }
            bulk_commit();
        }
        if constexpr (kPeer == 0) { if (p.accepted_out && mine) p.accepted_out[w - p.w_lo] = accept ? 1 : 0; }
        if (p.rung_acc) count_rung_accepts(p.rung_acc, act ? w / p.npb : 0u, act, accept, lane);

        if (tile_n >= ntiles) break;
        tile = tile_n;
        w = w_n;
        act = act_n;
        lp_old = lp_n;
        if constexpr (kPeer != 0) home = home_n;
        if constexpr (kRemap) rsrc = rsrc_n;  // (blocks 1.. of the tile and the re-fetch of rejected rows read through it)
        if constexpr (kStageB) stage_phase ^= 1u;
    }

    if constexpr (!kDirect) bulk_wait_all();
    for (int off = 16; off; off >>= 1) n_acc += __shfl_xor_sync(0xffffffffu, n_acc, off);
    if (lane == 0 && n_acc) atomicAdd(p.stats, n_acc);
}

}  // namespace scorus
