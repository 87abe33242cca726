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
#include <cstdint>
#include <cuda_runtime.h>

#include "logprob.cuh"
#include "rng.cuh"

namespace scorus {

enum : int { FK_ALL = 0, FK_BITS = 1, FK_ONEHOT = 2 };

struct StepParams {
    double *ens;            // [n][pitch]
    double *lp;             // [n]
    const double *beta;     // [nbeta] (device)
    const double *pparams;  // plugin params (device)
    const uint32_t *order;  // [n] matching order (device) or nullptr -> keyed bijection
    const double *z_in;     // fixed streams (device) or nullptr -> Philox
    const double *u_in;
    const uint8_t *flags_in;  // [n][flag_len] bytes or nullptr
    uint8_t *accepted_out;    // [n] or nullptr
    unsigned long long *stats;  // [0] accepted, [1] proposed
    uint64_t seed, step, onehot;
    uint64_t flag_thr;      // Bernoulli threshold on flag_bits random bits
    double inv_sqrt_a, z_range;
    uint32_t n, npb, dim, pitch, nbeta, nparams;
    uint32_t flag_len, flag_bits, flag_kind;  // flag_kind: SCORUS_UFS_*
    uint32_t row_bytes, row_stride;           // bytes copied per row / shared-memory row stride
    uint32_t stages, warps, ntiles, flag_words;
    uint32_t bij_bits;
};

// ---- PTX wrappers: mbarrier + bulk async copies (TMA, non-tensor form) ----
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {}
}
__device__ __forceinline__ void bulk_load(uint32_t dst_smem, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
        "l"(src), "r"(bytes), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void bulk_store(void *dst, uint32_t src_smem, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_smem),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all()
{
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

// walker at matching position j (positions 2k, 2k+1 are partners; both in the same rung)
__device__ __forceinline__ uint32_t walker_at(const StepParams &p, uint32_t j)
{
    if (p.order) return __ldg(p.order + j);
    uint32_t r = j / p.npb;
    uint32_t k = j - r * p.npb;
    BijKeys K = bij_keys(p.seed, p.step, ST_PERM, r);
    return r * p.npb + bij(k, p.npb, p.bij_bits, K);
}

constexpr int kChunk = 8;  // doubles per lane between two __syncwarp (4 x LDS.128 per row)

// Shared-memory carve-up (dynamic):
//   [0, 8*warps*stages)                  mbarriers, one per (warp, stage)
//   then per warp: stages * 32 * row_stride bytes of rows, then 32*flag_words*4 bytes of flag bits
template <class Plugin, bool kAllFlags>
__global__ void __launch_bounds__(512, 1) step_kernel(const StepParams p)
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
            U4 r = draw(p.seed, p.step, ST_ZU, w, 0);
            z = z_from_uniform(u53(r.x, r.y), p.inv_sqrt_a, p.z_range);
            u = u53(r.z, r.w);
        }

        // ---- update flags (:38-54) ----
        uint32_t nphi = p.dim;
        uint32_t hot = 0;
        if (!kAllFlags) {
            if (p.flag_kind == 2u) {  // ONE_HOT_CYCLE: the closure is called once per walker, in order
                hot = (uint32_t)((p.onehot + 1ull + (uint64_t)w) % (uint64_t)p.dim);
                nphi = 1;
            } else if (p.flag_kind == 3u) {  // MASK: caller-evaluated Func / fixed streams
                nphi = 0;
                for (uint32_t wd = 0; wd < p.flag_words; ++wd) {
                    uint32_t bits = 0;
                    for (uint32_t b = 0; b < 32u; ++b) {
                        uint32_t d = wd * 32u + b;
                        if (d >= p.dim) break;
                        bool f = d >= p.flag_len ? true
                                                 : (act ? __ldg(p.flags_in + (size_t)w * p.flag_len + d) != 0 : true);
                        bits |= (uint32_t)f << b;
                        nphi += (f && d < p.flag_len) ? 1u : 0u;
                    }
                    flagw[wd * 32u + lane] = bits;
                }
            } else {  // PROB: Bernoulli(p) per coordinate, redrawn until at least one is set
                const uint32_t bpf = p.flag_bits, fpc = 128u / bpf;
                const uint32_t nblk = (p.dim + fpc - 1u) / fpc;
                const uint64_t vmask = bpf == 32u ? 0xFFFFFFFFull : ((1ull << bpf) - 1ull);
                for (uint32_t attempt = 0;; ++attempt) {
                    nphi = 0;
                    uint32_t bits = 0;
                    U4 r = {0u, 0u, 0u, 0u};
                    for (uint32_t d = 0, q = 0, blk = 0; d < p.dim; ++d) {
                        if (q == 0) r = draw(p.seed, p.step, ST_FLAGS, w, attempt * nblk + blk);
                        const uint32_t bit = q * bpf, wi = bit >> 5;
                        const uint32_t word = wi == 0 ? r.x : wi == 1 ? r.y : wi == 2 ? r.z : r.w;
                        const uint64_t v = ((uint64_t)(word >> (bit & 31u))) & vmask;
                        if (v < p.flag_thr) { bits |= 1u << (d & 31u); ++nphi; }
                        if ((d & 31u) == 31u || d + 1u == p.dim) { flagw[(d >> 5) * 32u + lane] = bits; bits = 0; }
                        if (++q == fpc) { q = 0; ++blk; }
                    }
                    if (nphi > 0 || !act) break;
                }
            }
        }

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
            if (!kAllFlags && p.flag_kind != 2u) fbits = flagw[((c * 2u) >> 5) * 32u + lane] >> ((c * 2u) & 31u);
#pragma unroll
            for (int q = 0; q < kChunk / 2; ++q) {
                const uint32_t d0 = (c + q) * 2u;
                if (d0 < p.dim) {
                    bool f0 = true, f1 = true;
                    if (!kAllFlags) {
                        if (p.flag_kind == 2u) { f0 = d0 == hot; f1 = d0 + 1u == hot; }
                        else { f0 = (fbits >> (2 * q)) & 1u; f1 = (fbits >> (2 * q + 1)) & 1u; }
                    }
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
            ++n_acc;
        }
        bulk_commit();
        if (p.accepted_out && act) p.accepted_out[w] = accept ? 1 : 0;
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
