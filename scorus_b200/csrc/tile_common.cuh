// tile_common.cuh -- pieces shared by the step kernels: launch parameters, the PTX wrappers for
// mbarrier + bulk async copies (TMA, non-tensor form), the matching order and the update-flag
// generator.
#pragma once
#include "swap_params.cuh"
#include <cstdint>
#include <cuda_runtime.h>

#include "logprob.cuh"
#include "rng.cuh"

namespace scorus {

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
    uint8_t *dirty;           // [n] or nullptr: set to 1 when a walker is accepted (never cleared by the kernel)
    unsigned long long *stats;  // [0] accepted, [1] proposed
    uint64_t seed, step, onehot;
    uint64_t flag_thr;      // Bernoulli threshold on flag_bits random bits
    // flag_bits == 32 (a p that is not a multiple of 2^-16): [dim] 32-bit thresholds of the position of the FIRST set
    // flag given that one is set, cdf[j] = P(first <= j | any) * 2^32 (gen_flags_first below); device memory
    const uint32_t *flag_cdf;
    double inv_sqrt_a, z_range;
    uint32_t n, npb, dim, pitch, nbeta, nparams;
    uint32_t flag_len, flag_bits, flag_kind;  // flag_kind: SCORUS_UFS_*
    uint32_t row_bytes, row_stride;           // bytes copied per row / shared-memory row stride
    uint32_t stages, warps, ntiles, flag_words;
    uint32_t bij_bits;
    uint32_t debug_identity;  // measurement only (SCORUS_DEBUG_IDENTITY=1): position j holds walker j
    // ---- sharding (one process per GPU) ----
    // Random streams are keyed by GLOBAL walker / rung ids so that a sharded run draws exactly what
    // the single-GPU run draws: global walker = w_off + w, global rung = rung_off + r.
    uint32_t w_off, rung_off;
    uint64_t n_global;        // walkers of the whole ensemble (one-hot cycle advances by this per step)
    // Global-matching mode (register kernel only): ids in `order` are global, rows and log-probs are
    // read from the gathered copies src_ens / src_lp, results are written for walkers in
    // [w_lo, w_hi) only, at local row w - w_lo.  Single-GPU: src_* = ens / lp, [0, n).
    const double *src_ens;
    const double *src_lp;
    const uint32_t *n_dev;    // if set, the number of matching positions is read from device memory
    // One-sided ("peer") mode: the pair list in `order` holds the pairs this rank OWNS (global ids); the kernel reads
    // both rows wherever they live -- peer_ens[q] / peer_lp[q] are the CUDA-IPC mappings of rank q's buffers (own
    // rank: the local buffers) -- and writes both walkers' accepted rows and log-probs back to their home ranks over
    // NVLink.  Walker w lives on rank w / peer_n at row w % peer_n.  peer_n == 0: not in this mode.
    double *peer_ens[16];
    double *peer_lp[16];
    uint32_t peer_n;
    uint32_t peer_stage;  // the launch has room for the partner-row stage (16 rows + mbarriers per warp): kPeer == 2
    uint32_t w_lo, w_hi;
    // Blocked matchings: every rung is cut into blocks of `mblk` walkers that are matched
    // separately (mblk == npb: the reference's whole-rung matching); block c of a rung takes its bijection keys at
    // sub = 2 (c + mblk_off), so that a shard holding block c of the whole ensemble draws what the single-GPU run draws.
    uint32_t mblk, mblk_bits, mblk_off;
    // Dynamic tile scheduling (register kernel, single-GPU mode): after its first tile a warp takes
    // the next unclaimed one, tile = nwarps + (atomicAdd(tile_counter, 1) - tile_base).  Every launch
    // advances the counter by exactly ntiles, so the host knows the next base without a reset.
    unsigned long long *tile_counter;
    unsigned long long tile_base;
    // One-block ensembles (register kernel, grid == 1): the exact shuffle of every rung is computed by the
    // step kernel itself in shared memory instead of a separate exact_order_kernel launch.
    uint32_t fused_order;
    // Chain statistics: per-rung acceptance counts, indexed like beta; nullptr for
    // single-rung calls, whose count is stats[0].
    unsigned long long *rung_acc;
    // Register kernel: when a warp learns the walkers of its NEXT tile (one tile ahead), every lane asks L2 for that
    // walker's row, so that the block loads of the next tile find their lines in L2 instead of waiting on DRAM
    // (SCORUS_L2_PREFETCH; by default only for rows of 128..256 bytes, where it pays).
    uint32_t l2_prefetch;
    // Register kernel, direct mode, single GPU (kRemap): the step right after a ladder swap.  The swap left its result
    // interleaved, remap[w] = {log-prob of the walker that belongs in slot w, slot its row is still in} (swap_params.cuh
    // SwapOut), and moved nothing: this step reads walker w's row from src_ens at remap[w].src and writes EVERY row --
    // proposal if accepted, the old row if not -- and every log-prob to slot w of ens / lp, the context's second buffer,
    // which the host then makes the ensemble.  The swap's row pass (read + write of the whole ensemble) disappears.
    const void *remap;
    // The same variant as the first step of a host-buffer call (api_host.cu): remap == nullptr, src_ens is the caller's
    // pinned, device-mapped ensemble (rows pitch == dim doubles apart) -- the rows are never uploaded: the step streams
    // them over PCIe itself, writes every row to ens (the device copy the following steps use) and every ACCEPTED row
    // and log-prob back to host_ens / host_lp as it is decided, so the two directions of the link run at once.
    double *host_ens;
    double *host_lp;
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
// Programmatic dependent launch (sm_90+): a kernel launched with programmatic stream serialization
// may start while its predecessor drains; grid_dep_wait blocks until the predecessor has completed and
// its writes are visible, grid_dep_launch lets the successor's blocks be scheduled early.
__device__ __forceinline__ void grid_dep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void grid_dep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

// walker at matching position j (positions 2k, 2k+1 are partners; both in the same rung)
__device__ __forceinline__ uint32_t walker_at(const StepParams &p, uint32_t j)
{
    if (p.order) return __ldg(p.order + j);
    if (p.debug_identity) return j;
    const uint32_t blk = j / p.mblk, r = j / p.npb;
    const uint32_t c = blk - r * (p.npb / p.mblk);
    BijKeys K = bij_keys(p.seed, p.step, ST_PERM, r + p.rung_off, 2u * (c + p.mblk_off));
    return blk * p.mblk + bij(j - blk * p.mblk, p.mblk, p.mblk_bits, K);
}

// ---- update flags (src/mcmc/ensemble_sample.rs:38-54), one lane per walker ----
// Flag words live in shared memory as flagw[word * 32 + lane]: bit (d & 31) of word (d >> 5) set
// means coordinate d moves.  Returns nphi = number of set flags that count (:164-167).

// Prob(p): BPF random bits per coordinate, flag = value < thr; redrawn (attempt + 1) until one is set
// (ensemble_sample.rs:43-50).  The loop is bounded: after kMaxFlagAttempts all-false draws (probability
// (1-p)^(64 D): never at the p*D of any example) ONE coordinate is set, uniformly chosen from a further draw
// (sub = kMaxFlagAttempts * nblk) -- which is also the exact p -> 0 limit of the reference's conditional law,
// "exactly one flag, uniformly placed".  No input can make a kernel spin; oracle/philox_streams.c restates it.
constexpr uint32_t kMaxFlagAttempts = 64;
template <int BPF>
__device__ __forceinline__ uint32_t gen_flags_prob(const StepParams &p, uint64_t step, uint32_t w, uint32_t *flagw,
                                                   uint32_t lane, bool act)
{
    constexpr uint32_t FPC = 128u / BPF;   // flags per Philox call
    constexpr uint32_t VPW = 32u / BPF;    // values per 32-bit random word
    constexpr uint32_t VMASK = BPF == 32 ? 0xFFFFFFFFu : ((1u << (BPF & 31)) - 1u);
    const uint32_t nblk = (p.dim + FPC - 1u) / FPC;
    const uint32_t thr = (uint32_t)(p.flag_thr > 0xFFFFFFFFull ? 0xFFFFFFFFull : p.flag_thr);
    const bool always = p.flag_thr > (uint64_t)VMASK;  // p == 1: every value is below the threshold
    uint32_t nphi = 0;
    for (uint32_t attempt = 0;; ++attempt) {
        nphi = 0;
        uint32_t cur = 0;
        for (uint32_t blk = 0; blk < nblk; ++blk) {
            const U4 r = draw(p.seed, step, ST_FLAGS, w + p.w_off, attempt * nblk + blk);
            const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
            for (uint32_t wi = 0; wi < 4; ++wi) {
                uint32_t bits = 0;  // VPW flag bits from this random word
                if (BPF == 1) {
                    bits = always ? 0xFFFFFFFFu : (thr ? ~rr[wi] : 0u);  // value < 1  <=>  bit clear
                } else {
#pragma unroll
                    for (uint32_t q = 0; q < VPW; ++q) {
                        const uint32_t v = (rr[wi] >> ((q * BPF) & 31u)) & VMASK;
                        bits |= (uint32_t)(always || v < thr) << q;
                    }
                }
                const uint32_t d0 = blk * FPC + wi * VPW;  // first coordinate these bits belong to
                if (d0 < p.dim) {
                    if (d0 + VPW > p.dim) bits &= (1u << (p.dim - d0)) - 1u;
                    nphi += __popc(bits);
                    cur |= bits << (d0 & 31u);
                    if (((d0 + VPW) & 31u) == 0u || d0 + VPW >= p.dim) {
                        flagw[(d0 >> 5) * 32u + lane] = cur;
                        cur = 0;
                    }
                }
            }
        }
        if (nphi > 0 || !act) break;
        if (attempt + 1u == kMaxFlagAttempts) {  // every word above was just written as zero
            const U4 r = draw(p.seed, step, ST_FLAGS, w + p.w_off, kMaxFlagAttempts * nblk);
            const uint32_t d = __umulhi(r.x, p.dim);
            flagw[(d >> 5) * 32u + lane] = 1u << (d & 31u);
            nphi = 1;
            break;
        }
    }
    return nphi;
}

// Prob(p) at 32 bits per coordinate WITHOUT the redraw loop.  The reference draws D Bernoulli(p) flags and starts over
// while none is set (ensemble_sample.rs:43-50): the result is D iid flags conditioned on "at least one".  The same law,
// drawn in one pass: the position J of the first set flag is truncated-geometric, P(J <= j | any) =
// (1 - (1-p)^(j+1)) / (1 - (1-p)^D) -- read off a table of 32-bit thresholds (computed once on the host) with the random
// word of coordinate 0, which the construction never needs as a Bernoulli draw (coordinate 0 is either the first set
// flag or clear); coordinates before J are clear, J is set, coordinates after J are iid Bernoulli(p) from their own words.
// No loop, so nothing to bound, and no extra random numbers.  At c4 (10-D, p = 0.2: 11 % of the walkers redraw, i.e.
// nearly every warp ran the three Philox calls of an attempt two or three times) this was 40 % of the step's
// instructions (profiles/r2_c4_step_reg_ncu_summary.txt).  oracle/philox_streams.c restates it.
__device__ __forceinline__ uint32_t gen_flags_first(const StepParams &p, uint64_t step, uint32_t w, uint32_t *flagw,
                                                    uint32_t lane)
{
    const uint32_t nblk = (p.dim + 3u) / 4u;
    const uint32_t thr = (uint32_t)(p.flag_thr > 0xFFFFFFFFull ? 0xFFFFFFFFull : p.flag_thr);
    U4 r = draw(p.seed, step, ST_FLAGS, w + p.w_off, 0);
    uint32_t lo = 0, hi = p.dim - 1u;  // J = number of j < dim-1 with word0 >= cdf[j]
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (r.x >= __ldg(p.flag_cdf + mid)) lo = mid + 1u;
        else hi = mid;
    }
    const uint32_t J = lo;
    uint32_t nphi = 0, cur = 0;
    for (uint32_t blk = 0; blk < nblk; ++blk) {
        if (blk) r = draw(p.seed, step, ST_FLAGS, w + p.w_off, blk);
        const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (uint32_t wi = 0; wi < 4; ++wi) {
            const uint32_t d = blk * 4u + wi;
            if (d < p.dim) {
                const uint32_t f = d < J ? 0u : (d == J ? 1u : (uint32_t)(rr[wi] < thr));
                nphi += f;
                cur |= f << (d & 31u);
                if (((d + 1u) & 31u) == 0u || d + 1u >= p.dim) {
                    flagw[(d >> 5) * 32u + lane] = cur;
                    cur = 0;
                }
            }
        }
    }
    return nphi;
}

// `step`, `onehot`: the counters the Prob stream / the cycling closure are keyed by (p.step, p.onehot, except in the
// kernels that run several steps or half-passes per launch)
__device__ __forceinline__ uint32_t gen_flags(const StepParams &p, uint64_t step, uint64_t onehot, uint32_t w, uint32_t *flagw,
                                              uint32_t lane, bool act)
{
    if (p.flag_kind == 2u) {  // ONE_HOT_CYCLE: the cycling closure, called once per walker in order
        const uint32_t hot = (uint32_t)((onehot + 1ull + (uint64_t)w + (uint64_t)p.w_off) % (uint64_t)p.dim);
        for (uint32_t wd = 0; wd < p.flag_words; ++wd)
            flagw[wd * 32u + lane] = (hot >> 5) == wd ? (1u << (hot & 31u)) : 0u;
        return 1u;
    }
    if (p.flag_kind == 3u) {  // MASK: caller-evaluated Func / fixed streams, [n][flag_len] bytes
        uint32_t nphi = 0;
        for (uint32_t wd = 0; wd < p.flag_words; ++wd) {
            uint32_t bits = 0;
            for (uint32_t b = 0; b < 32u; ++b) {
                const uint32_t d = wd * 32u + b;
                if (d >= p.dim) break;
                // coordinates beyond flag_len always move and are not counted (:68, :164-167)
                const bool f = d >= p.flag_len ? true
                                               : (act ? __ldg(p.flags_in + (size_t)w * p.flag_len + d) != 0 : true);
                bits |= (uint32_t)f << b;
                nphi += (f && d < p.flag_len) ? 1u : 0u;
            }
            flagw[wd * 32u + lane] = bits;
        }
        return nphi;
    }
    switch (p.flag_bits) {  // PROB
    case 1: return gen_flags_prob<1>(p, step, w, flagw, lane, act);
    case 2: return gen_flags_prob<2>(p, step, w, flagw, lane, act);
    case 4: return gen_flags_prob<4>(p, step, w, flagw, lane, act);
    case 8: return gen_flags_prob<8>(p, step, w, flagw, lane, act);
    case 16: return gen_flags_prob<16>(p, step, w, flagw, lane, act);
    default: return p.flag_cdf ? gen_flags_first(p, step, w, flagw, lane) : gen_flags_prob<32>(p, step, w, flagw, lane, act);
    }
}
__device__ __forceinline__ uint32_t gen_flags(const StepParams &p, uint64_t step, uint32_t w, uint32_t *flagw, uint32_t lane,
                                              bool act)
{
    return gen_flags(p, step, p.onehot, w, flagw, lane, act);
}
__device__ __forceinline__ uint32_t gen_flags(const StepParams &p, uint32_t w, uint32_t *flagw, uint32_t lane, bool act)
{
    return gen_flags(p, p.step, p.onehot, w, flagw, lane, act);
}

// per-rung acceptance counts: one atomic per tile when the tile lies inside one rung (the usual case),
// one per accepted walker for a tile that straddles a rung boundary.  All 32 lanes call this.
__device__ __forceinline__ void count_rung_accepts(unsigned long long *rung_acc, uint32_t rung, bool act, bool accept,
                                                   uint32_t lane)
{
    const unsigned m = __ballot_sync(0xffffffffu, accept);
    const uint32_t r0 = __shfl_sync(0xffffffffu, rung, 0);
    if (__all_sync(0xffffffffu, !act || rung == r0)) {
        if (lane == 0 && m) atomicAdd(rung_acc + r0, (unsigned long long)__popc(m));
    } else if (accept) {
        atomicAdd(rung_acc + rung, 1ull);
    }
}

}  // namespace scorus
