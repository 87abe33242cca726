// swap_kernel.cuh -- K4 (exact-shuffle order generator) and K5 (temperature-ladder swap).
//
// swap_walkers (src/mcmc/utils.rs:72-112) sweeps the rungs hottest -> coldest; stage i pairs
// slot j1 of rung i with slot j2 of rung i-1 through a permutation jvec_i (j1 = jvec_i[j2]) and
// swaps the two walkers with probability min(1, exp((beta_{i-1}-beta_i)(lp_hot - lp_cold))).
// The B-1 stages look sequential, but the slots a stage touches are linked only along CHAINS:
// slot (i-1, j2) is next touched by stage i-1 as its hot slot.  Following
// a_{i-1} = jinv_i[a_i] from a_{B-1} = c gives npb disjoint chains of B slots, each swept by one
// thread with the travelling walker's (log-prob, origin) carried in registers -- no inter-block
// synchronisation, npb-way parallel (the slot walk: also over segments of the ladder, see below).  Decisions depend only on
// log-probs, so the sweep produces a source map src[slot] and the final log-probs first; rows
// move once afterwards (gather into scratch, copy back), only the slots that changed.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include <type_traits>

#include "rng.cuh"
#include "swap_params.cuh"

namespace scorus {

// Exact uniform shuffle of every rung by ranking 64-bit Philox keys (npb <= kExactShuffleMax).
// order[r*npb + rank] = r*npb + i.  One block per rung.
__global__ void __launch_bounds__(1024) exact_order_kernel(uint64_t seed, uint64_t step, uint32_t npb,
                                                           uint32_t *order, uint32_t w_off)
{
    __shared__ uint64_t keys[kExactShuffleMax];
    const uint32_t r = blockIdx.x, i = threadIdx.x;
    uint64_t ki = 0;
    if (i < npb) {
        U4 o = draw(seed, step, ST_PERM_KEY, w_off + r * npb + i, 0);
        ki = ((uint64_t)o.y << 32) | o.x;
        keys[i] = ki;
    }
    __syncthreads();
    if (i < npb) {
        uint32_t rank = 0;
        for (uint32_t j = 0; j < npb; ++j) {
            uint64_t kj = keys[j];
            rank += (kj < ki) || (kj == ki && j < i);
        }
        order[r * npb + rank] = r * npb + i;
    }
}

// Exact jvec for the swap (npb <= kExactShuffleMax): block b is hot rung i = b+1, stage
// s = nbeta-1-i.  jvec[s][rank] = k and its inverse jinv[s][k] = rank.
__global__ void __launch_bounds__(1024) exact_jvec_kernel(uint64_t seed, uint64_t swap_call, uint32_t npb,
                                                          uint32_t nbeta, uint32_t *jinv)
{
    __shared__ uint64_t keys[kExactShuffleMax];
    const uint32_t i = blockIdx.x + 1, k = threadIdx.x, s = nbeta - 1 - i;
    uint64_t kk = 0;
    if (k < npb) {
        U4 o = draw(seed, swap_call, ST_SWAP_KEY, i * npb + k, 0);
        kk = ((uint64_t)o.y << 32) | o.x;
        keys[k] = kk;
    }
    __syncthreads();
    if (k < npb) {
        uint32_t rank = 0;
        for (uint32_t j = 0; j < npb; ++j) {
            uint64_t kj = keys[j];
            rank += (kj < kk) || (kj == kk && j < k);
        }
        jinv[(size_t)s * npb + k] = rank;
    }
}


// The sweep is latency, not work: 16384 chains are 512 warps -- a few per SM -- so every instruction of a chain's
// thread is paid at its full latency.  History (profiles/r2_notes.md section 2): one dependent loop per chain with keys,
// bijection, a DRAM-latency log-prob load, a Philox draw and an exp per stage: 76 us at 64 rungs x 16384 chains, as long
// as a step; uniforms hoisted into a parallel kernel and the chain split in two passes per thread (slots, then
// decisions): 9.7 + 43.8 us, still 188 instructions per stage issued by a lone warp per scheduler.  What has to be
// sequential PER CHAIN is only the carry of the travelling walker; which slots a chain visits is a function of the
// permutations alone, and every slot of a rung lies on exactly one chain -- so the slot walk parallelises over
// SEGMENTS of the ladder as well as over chains:
//   * swap_paths_kernel, one thread per (segment g of kSwapSeg stages, hot slot x of the segment's first stage): walks
//     the sub-chain that starts at x (bijection or jinv, integer work), then -- independent per stage -- draws the
//     swap uniform of each visited cold slot, takes its logarithm, gathers the cold slot's log-prob, and writes the
//     three into tables indexed [stage][x].  ceil((B-1)/kSwapSeg) x npb threads: 4096 warps at c4, throughput-bound.
//   * swap_walk_kernel, one thread per chain: per segment ONE dependent fetch (the records of sub-chain x, all stages at
//     once, and with the last of them the next segment's x; the next segment is requested before this one is decided),
//     then per stage the decision r < min(1, exp(x)), x = (beta_cold - beta_hot)(lp_carry - lp_cold), utils.rs:63-69,105,
//     taken WITHOUT the exp wherever ln r and x are not within 1e-9 of each other: ln r < x - d  =>  r < exp(x) (1 - 1e-9...),
//     far outside the rounding of either side, so the shortcut returns exactly what the exp comparison returns; the
//     one-in-10^9 close calls, r = 0 and non-finite x take the literal path.  The dependent chain per stage is two adds,
//     two multiplies, the compares and a select.
constexpr uint32_t kSwapSeg = 8;  // stages per segment

// swap uniform of (stage s, cold slot j2); hot rung i = nbeta-1-s
__device__ __forceinline__ double swap_uniform(const SwapParams &p, uint32_t s, uint32_t j2)
{
    if (p.r_in) return __ldg(p.r_in + (size_t)s * p.npb + j2);
    const U4 o = draw(p.seed, p.swap_call, ST_SWAP_U, j2, p.nbeta - 1u - s);
    return to_unit<double>(o.x, o.y);
}

__global__ void __launch_bounds__(256) swap_paths_kernel(const SwapParams p)
{
    __shared__ uint32_t seg_keys[kSwapSeg][8];
    const uint32_t B = p.nbeta, npb = p.npb, S = B - 1u;
    const uint32_t s0 = blockIdx.y * kSwapSeg;
    const uint32_t ns = min(kSwapSeg, S - s0);
    if (!p.jinv) {
        if (threadIdx.x < ns) {
            const BijKeys K = bij_keys(p.seed, p.swap_call, ST_SWAP_PERM, B - 1u - (s0 + threadIdx.x));
#pragma unroll
            for (int q = 0; q < 8; ++q) seg_keys[threadIdx.x][q] = K.k[q];
        }
        __syncthreads();
    }
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= npb) return;
    // the sub-chain's slots: hot slot a of stage s meets cold slot j2 = next_s(a), the hot slot of stage s + 1
    uint32_t j2v[kSwapSeg];
    uint32_t a = x;
#pragma unroll
    for (uint32_t k = 0; k < kSwapSeg; ++k) {
        if (k < ns) {
            uint32_t j2;
            if (p.jinv) {
                j2 = __ldg(p.jinv + (size_t)(s0 + k) * npb + a);
            } else {
                BijKeys K;
#pragma unroll
                for (int q = 0; q < 8; ++q) K.k[q] = seg_keys[k][q];
                j2 = bij(a, npb, p.bij_bits, K);
            }
            j2v[k] = j2;
            a = j2;
        }
    }
    // per stage, independent of one another: cold log-prob, ln r
#pragma unroll
    for (uint32_t k = 0; k < kSwapSeg; ++k) {
        if (k < ns) {
            const uint32_t s = s0 + k, i = B - 1u - s, j2 = j2v[k];
            const double r = swap_uniform(p, s, j2);
            SwapRec rec;
            rec.lp = __ldg(p.lp_src + (size_t)(i - 1u) * npb + j2);
            // ln r in f32: the decision's guard band is orders of magnitude wider than the rounding (swap_walk_kernel).
            // Anything but 0 < r < 1 (only a caller-supplied r can be) is marked NaN and decided by the literal expression
            rec.lr = (r > 0.0 && r < 1.0) ? logf((float)r) : real_nan<float>();
            rec.cold = (i - 1u) * npb + j2;
            p.tbl_rec[(size_t)s * npb + x] = rec;
        }
    }
}

// kSmem: the per-stage beta differences and exchange counters of the block live in shared memory (any ladder of up to
// kSwapSmemStages stages; longer ones read beta and count through global memory)
constexpr uint32_t kSwapSmemStages = 3072;

template <bool kSmem>
__global__ void __launch_bounds__(128) swap_walk_kernel(const SwapParams p)
{
    extern __shared__ __align__(16) unsigned char walk_smem[];
    const uint32_t B = p.nbeta, npb = p.npb, S = B - 1u, T = blockDim.x, tid = threadIdx.x;
    double *dbeta = reinterpret_cast<double *>(walk_smem);            // [S] beta_cold - beta_hot of stage s
    unsigned int *stage_cnt = reinterpret_cast<unsigned int *>(dbeta + S);  // [S] accepted exchanges of this block
    const bool count = p.stage_swapped != nullptr;
    if constexpr (kSmem) {
        for (uint32_t s = tid; s < S; s += T) {
            const uint32_t i = B - 1u - s;
            dbeta[s] = __ldg(p.beta + i - 1u) - __ldg(p.beta + i);  // utils.rs:63-66: (beta2 - beta1), hot rung i
            stage_cnt[s] = 0u;
        }
        __syncthreads();
    }
    // every lane runs the sweep (ballots stay whole-warp); lanes past the last chain shadow it and write nothing
    const uint32_t c_raw = blockIdx.x * T + tid;
    const bool valid = c_raw < npb;
    const uint32_t c = valid ? c_raw : npb - 1u;
    uint32_t hot = (B - 1u) * npb + c;  // slot of the chain at the current stage (global index)
    uint32_t carry_src = hot;
    double carry_lp = p.lp_src[hot];
    unsigned long long nsw = 0;

    // the walker that ends in `slot`: where its row is now, and its log-prob
    auto store_slot = [&](uint32_t slot, uint32_t src, double lp) {
        if (p.out) {
            SwapOut o;
            o.lp = lp; o.src = src; o.pad = 0u;
            p.out[slot] = o;
        } else {
            p.src[slot] = src;
            p.lp[slot] = lp;
        }
    };
    SwapRec recn[kSwapSeg];
    double dbn[kSwapSeg];
    // the records of sub-chain x over the stages of the segment that starts at s0 (stages past the ladder's end re-read
    // its last record, never used: no select behind the loads)
    auto fetch = [&](uint32_t s0, uint32_t x) {
#pragma unroll
        for (uint32_t k = 0; k < kSwapSeg; ++k) {
            const uint32_t s = s0 + k < S ? s0 + k : S - 1u;
            recn[k] = p.tbl_rec[(size_t)s * npb + x];
            if constexpr (kSmem) dbn[k] = dbeta[s];
            else dbn[k] = __ldg(p.beta + (B - 2u - s)) - __ldg(p.beta + (B - 1u - s));
        }
    };
    auto segment = [&](uint32_t s0, auto full_tag) {
        constexpr bool kFull = decltype(full_tag)::value;
        SwapRec rec[kSwapSeg];
        uint32_t cold[kSwapSeg];
        double db[kSwapSeg];
#pragma unroll
        for (uint32_t k = 0; k < kSwapSeg; ++k) { rec[k] = recn[k]; cold[k] = recn[k].cold; db[k] = dbn[k]; }
        if constexpr (kFull) {
            // the next segment's sub-chain starts where this one ends: requested before this segment is decided
            if (s0 + kSwapSeg < S) fetch(s0 + kSwapSeg, cold[kSwapSeg - 1u] - (B - 1u - s0 - kSwapSeg) * npb);
        }
        uint32_t swbits = 0;
#pragma unroll
        for (uint32_t k = 0; k < kSwapSeg; ++k) {
            const uint32_t s = s0 + k;
            if (kFull || s < S) {  // hottest stage first (utils.rs:89-108)
                const double lp_cold = rec[k].lp, lr = (double)rec[k].lr;
                const double x = db[k] * (-lp_cold + carry_lp);  // exchange_prob's exponent, utils.rs:63-66
                const double d = swap_guard<double>() * (1.0 + fabs(x));
                const bool c1 = (lr == lr) & (x >= d);  // exp(x) > 1: probability 1, and 0 < r < 1
                const bool c2 = lr < x - d;             // r < exp(x) < 1 by a margin far above any rounding
                const bool c3 = lr > x + d;             // r > exp(x) by such a margin
                bool sw = c1 | c2;
                if (__builtin_expect(!(c1 | c2 | c3), 0)) {
                    // too close to call, r outside (0, 1), or x not finite: the literal expression
                    const uint32_t i = B - 1u - s;
                    const double r = swap_uniform(p, s, cold[k] - (i - 1u) * npb);
                    const double ex = exp(x);
                    const double ep = ex > 1.0 ? 1.0 : ex;
                    sw = r < ep;  // utils.rs:105 (NaN -> no swap)
                }
                // swap: the cold walker comes up and is final here, the carried walker goes down; else the carried
                // walker stays here and the cold one travels on
                if (valid) store_slot(hot, sw ? cold[k] : carry_src, sw ? lp_cold : carry_lp);
                carry_src = sw ? carry_src : cold[k];
                carry_lp = sw ? carry_lp : lp_cold;
                hot = cold[k];
                swbits |= (sw ? 1u : 0u) << k;
            }
        }
        if (!valid) swbits = 0u;
        nsw += (unsigned)__popc(swbits);
        if (p.swapped_out && valid) {
#pragma unroll
            for (uint32_t k = 0; k < kSwapSeg; ++k)
                if (kFull || s0 + k < S) {
                    const uint32_t s = s0 + k, i = B - 1u - s;
                    p.swapped_out[(size_t)s * npb + (cold[k] - (i - 1u) * npb)] = (swbits >> k) & 1u;
                }
        }
        if (count) {
            // all lanes of the warp ran the same stages: one atomic per warp per stage, off the carry chain
#pragma unroll
            for (uint32_t k = 0; k < kSwapSeg; ++k)
                if (kFull || s0 + k < S) {
                    const unsigned m = __ballot_sync(0xffffffffu, (swbits >> k) & 1u);
                    if ((tid & 31u) == 0u && m) {
                        if constexpr (kSmem) atomicAdd(stage_cnt + s0 + k, (unsigned)__popc(m));
                        else atomicAdd(p.stage_swapped + (B - 2u - s0 - k), (unsigned long long)__popc(m));
                    }
                }
        }
    };

    fetch(0u, c);
    uint32_t s0 = 0;
    for (; s0 + kSwapSeg <= S; s0 += kSwapSeg) segment(s0, std::true_type{});
    if (s0 < S) segment(s0, std::false_type{});
    if (valid) store_slot(hot, carry_src, carry_lp);
    for (int off = 16; off; off >>= 1) nsw += __shfl_xor_sync(0xffffffffu, nsw, off);
    if ((tid & 31u) == 0 && nsw) atomicAdd(p.stats + 2, nsw);
    if constexpr (kSmem) {
        if (count) {
            __syncthreads();
            for (uint32_t s = tid; s < S; s += T) {
                const unsigned v = stage_cnt[s];
                if (v) atomicAdd(p.stage_swapped + (B - 2u - s), (unsigned long long)v);
            }
        }
    }
}

// rows after a swap in ONE pass: dst[slot] <- src_ens[src[slot]] for every slot (moved or not); the caller then makes
// dst the live buffer.  One warp-sized group of lanes per row chunk, 16-byte accesses.
__global__ void __launch_bounds__(256) gather_rows_kernel(const double2 *__restrict__ src_ens, double2 *__restrict__ dst,
                                                          const uint32_t *__restrict__ src, uint32_t n, uint32_t vec_per_row)
{
    const size_t total = (size_t)n * vec_per_row;
    for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (size_t)gridDim.x * blockDim.x) {
        const uint32_t slot = (uint32_t)(g / vec_per_row);
        const uint32_t v = (uint32_t)(g - (size_t)slot * vec_per_row);
        dst[g] = src_ens[(size_t)__ldg(src + slot) * vec_per_row + v];
    }
}

// the same from the interleaved result of the sweep (SwapOut), unpacking the log-probs on the way: what a swap leaves
// pending until the next step (which does this itself, step_reg_kernel.cuh kRemap) or until anything else looks
__global__ void __launch_bounds__(256) apply_swap_kernel(const double2 *__restrict__ src_ens, double2 *__restrict__ dst,
                                                         const SwapOut *__restrict__ out, double *__restrict__ lp, uint32_t n,
                                                         uint32_t vec_per_row)
{
    const size_t total = (size_t)n * vec_per_row;
    for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (size_t)gridDim.x * blockDim.x) {
        const uint32_t slot = (uint32_t)(g / vec_per_row);
        const uint32_t v = (uint32_t)(g - (size_t)slot * vec_per_row);
        const SwapOut o = out[slot];
        dst[g] = src_ens[(size_t)o.src * vec_per_row + v];
        if (v == 0u) lp[slot] = o.lp;
    }
}

// rows: scratch[slot] <- ens[src[slot]] for moved slots (phase 0), ens[slot] <- scratch[slot] (phase 1)
__global__ void __launch_bounds__(256) permute_rows_kernel(double2 *ens, double2 *scratch, const uint32_t *src,
                                                           uint32_t n, uint32_t vec_per_row, int phase)
{
    const size_t total = (size_t)n * vec_per_row;
    for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < total;
         g += (size_t)gridDim.x * blockDim.x) {
        const uint32_t slot = (uint32_t)(g / vec_per_row);
        const uint32_t v = (uint32_t)(g - (size_t)slot * vec_per_row);
        const uint32_t s = __ldg(src + slot);
        if (s != slot) {
            if (phase == 0) scratch[g] = ens[(size_t)s * vec_per_row + v];
            else ens[g] = scratch[g];
        }
    }
}

}  // namespace scorus
