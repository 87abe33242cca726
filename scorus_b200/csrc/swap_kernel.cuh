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
// thread is paid at its full latency, and a lone warp per scheduler issues IN ORDER: whatever is not on the critical path
// still stands in front of it.  History (profiles/r2_notes.md section 2): one dependent loop per chain with keys,
// bijection, a DRAM-latency log-prob load, a Philox draw and an exp per stage: 76 us at 64 rungs x 16384 chains, as long
// as a step; uniforms hoisted into a parallel kernel and the chain split in two passes per thread (slots, then
// decisions): 9.7 + 43.8 us; slot walk parallel over segments with one 16-byte record per (stage, x), eight scattered
// sectors per chain and segment: 14 + 21 us, bound by the scattered misses (a separate decide / emit pair: 17 + 11).
// What has to be sequential PER CHAIN is only the carry of the travelling walker's log-prob; which slots a chain visits
// is a function of the permutations alone, and every slot of a rung lies on exactly one chain:
//   * swap_paths_kernel, threads over (segment g of kSwapSeg stages, hot slot x of the segment's first stage, half h):
//     walks the sub-chain that starts at x (bijection or jinv, integer work), then -- independent per stage, one half of
//     the segment per thread -- draws the swap uniform of each visited cold slot, gathers the cold slot's log-prob and
//     turns the decision into a threshold on the carried log-prob (SwapSeg, swap_params.cuh).  8192 warps at c4.
//   * swap_walk_kernel, one thread per chain: per segment ONE dependent fetch (the chain's SwapSeg block: contiguous; its
//     last cold slot names the next segment's block, which is requested before this segment is decided), then per stage
//     two compares of the carried log-prob and the selects; inside the band (one decision in 10^4..10^5), for r outside
//     (0, 1) and for non-finite values the literal expression r < min(1, exp(x)), utils.rs:63-69,105.

// swap uniform of (stage s, cold slot j2); hot rung i = nbeta-1-s
__device__ __forceinline__ double swap_uniform(const SwapParams &p, uint32_t s, uint32_t j2)
{
    if (p.r_in) return __ldg(p.r_in + (size_t)s * p.npb + j2);
    const U4 o = draw(p.seed, p.swap_call, ST_SWAP_U, j2, p.nbeta - 1u - s);
    return to_unit<double>(o.x, o.y);
}

__global__ void __launch_bounds__(256) swap_paths_kernel(const SwapParams p)
{
    constexpr uint32_t kHalf = kSwapSeg / 2;
    constexpr uint32_t kXs = 128;                            // slots x per block, two threads (halves) each
    constexpr uint32_t NQ = sizeof(SwapSeg) / 16;            // 16-byte pieces of a block of the table
    constexpr uint32_t NT = kHalf * sizeof(double) / 16;     // ... of a half of mid[] / lp[]
    constexpr uint32_t kLpCol = kSwapSeg * sizeof(double) / 16, kWCol = 2 * kLpCol, kColdCol = kWCol + kSwapSeg * 4 / 16;
    static_assert(kHalf == 4 && kColdCol + 2 == NQ, "layout of SwapSeg");
    // the block's part of the table, staged so that it leaves in whole cache lines (a thread's pieces are 16..32 bytes
    // at a stride of sizeof(SwapSeg): written straight to global memory they cost the load/store unit one slot per lane
    // and piece, 6 us more at c4); rows padded by one piece: conflict-free 16-byte accesses
    __shared__ int4 stage[kXs * (NQ + 1)];
    __shared__ uint32_t seg_keys[kSwapSeg][8];
    __shared__ double seg_inv_db[kSwapSeg];
    const uint32_t B = p.nbeta, npb = p.npb, S = B - 1u;
    const uint32_t g = blockIdx.y, s0 = g * kSwapSeg;
    const uint32_t ns = min(kSwapSeg, S - s0);
    // which half of the segment's stages this thread draws for (its walk stops where its half ends)
    const uint32_t xl = threadIdx.x % kXs, h = threadIdx.x / kXs;
    const uint32_t k_lo = h * kHalf, k_hi = min(k_lo + kHalf, ns);
    if (threadIdx.x < ns) {
        const uint32_t i = B - 1u - (s0 + threadIdx.x);
        if (!p.jinv) {
            const BijKeys K = bij_keys(p.seed, p.swap_call, ST_SWAP_PERM, i);
#pragma unroll
            for (int q = 0; q < 8; ++q) seg_keys[threadIdx.x][q] = K.k[q];
        }
        const double db = __ldg(p.beta + i - 1u) - __ldg(p.beta + i);  // utils.rs:63-66: (beta2 - beta1), hot rung i
        seg_inv_db[threadIdx.x] = db > 0.0 ? 1.0 / db : real_nan<double>();
    }
    __syncthreads();
    const uint32_t x0 = blockIdx.x * kXs;
    const uint32_t x = min(x0 + xl, npb - 1u);  // (threads past the last slot shadow it; their rows are not written out)
    // the sub-chain's slots: hot slot a of stage s meets cold slot j2 = next_s(a), the hot slot of stage s + 1
    uint32_t j2v[kSwapSeg];
    uint32_t a = x;
#pragma unroll
    for (uint32_t k = 0; k < kSwapSeg; ++k) {
        if (k < k_hi) {
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
    // per stage, independent of one another: cold log-prob, threshold, band
    double mid[kHalf], lpv[kHalf];
    float wv[kHalf];
    uint32_t cv[kHalf];
#pragma unroll
    for (uint32_t kk = 0; kk < kHalf; ++kk) {
        const uint32_t k = k_lo + kk;
        mid[kk] = real_nan<double>(); lpv[kk] = 0.0; wv[kk] = 0.0f; cv[kk] = 0u;
        if (k < k_hi) {
            const uint32_t s = s0 + k, i = B - 1u - s, j2 = h ? j2v[kHalf + kk] : j2v[kk];  // (constant indices: registers)
            const double r = swap_uniform(p, s, j2);
            const double lp_cold = __ldg(p.lp_src + (size_t)(i - 1u) * npb + j2);
            // ln r in f32 and 1/db instead of a division: their errors (1e-7 relative) are far inside the band.
            // Anything but 0 < r < 1 (only a caller-supplied r can be) leaves mid = NaN: decided by the literal expression
            const double L = (r > 0.0 && r < 1.0) ? (double)logf((float)r) : real_nan<double>();
            const double inv_db = seg_inv_db[k];
            const double q = L * inv_db;
            // band: the old guard of the exponent (ln r vs x within 1e-5 (1 + |ln r|)) carried over to log-prob units,
            // plus the rounding of mid itself; rounded up into f32
            const double w = (swap_guard<double>() * (1.0 + fabs(L))) * inv_db + swap_round_guard<double>() * (fabs(lp_cold) + fabs(q));
            mid[kk] = lp_cold + q;
            lpv[kk] = lp_cold;
            wv[kk] = (float)w * 1.001f;
            cv[kk] = (i - 1u) * npb + j2;
        }
    }
    // this thread's half of row xl (for the stages past the ladder's end: mid = NaN, never read)
    if (p.debug_direct_store) {  // measurement / debugging only: straight to global memory
        if (x0 + xl < npb) {
            SwapSeg *blk = p.tbl_seg + (size_t)g * npb + x;
            int4 v[NT];
            memcpy(v, mid, sizeof(mid));
#pragma unroll
            for (uint32_t q = 0; q < NT; ++q) reinterpret_cast<int4 *>(blk->mid + k_lo)[q] = v[q];
            memcpy(v, lpv, sizeof(lpv));
#pragma unroll
            for (uint32_t q = 0; q < NT; ++q) reinterpret_cast<int4 *>(blk->lp + k_lo)[q] = v[q];
            int4 u;
            memcpy(&u, wv, 16);
            *reinterpret_cast<int4 *>(blk->w + k_lo) = u;
            memcpy(&u, cv, 16);
            *reinterpret_cast<int4 *>(blk->cold + k_lo) = u;
        }
        return;
    }
    {
        int4 *row = stage + xl * (NQ + 1u);
        int4 v[NT];
        memcpy(v, mid, sizeof(mid));
#pragma unroll
        for (uint32_t q = 0; q < NT; ++q) row[h * NT + q] = v[q];
        memcpy(v, lpv, sizeof(lpv));
#pragma unroll
        for (uint32_t q = 0; q < NT; ++q) row[kLpCol + h * NT + q] = v[q];
        int4 u;
        memcpy(&u, wv, 16);
        row[kWCol + h] = u;
        memcpy(&u, cv, 16);
        row[kColdCol + h] = u;
    }
    __syncthreads();
    const uint32_t nrows = min(kXs, npb - x0);
    int4 *dst = reinterpret_cast<int4 *>(p.tbl_seg + (size_t)g * npb + x0);
    for (uint32_t e = threadIdx.x; e < nrows * NQ; e += blockDim.x) dst[e] = stage[(e / NQ) * (NQ + 1u) + e % NQ];
}

// kSmem: the exchange counters of the block live in shared memory (any ladder of up to kSwapSmemStages stages; longer
// ones count through global memory)
constexpr uint32_t kSwapSmemStages = 8192;

template <bool kSmem>
__global__ void __launch_bounds__(128) swap_walk_kernel(const SwapParams p)
{
    extern __shared__ __align__(16) unsigned char walk_smem[];
    const uint32_t B = p.nbeta, npb = p.npb, S = B - 1u, T = blockDim.x, tid = threadIdx.x;
    unsigned int *stage_cnt = reinterpret_cast<unsigned int *>(walk_smem);  // [S] accepted exchanges of this block
    const bool count = p.stage_swapped != nullptr;
    if constexpr (kSmem) {
        if (count) {
            for (uint32_t s = tid; s < S; s += T) stage_cnt[s] = 0u;
            __syncthreads();
        }
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
            st16(p.out + slot, o);
        } else {
            p.src[slot] = src;
            p.lp[slot] = lp;
        }
    };
    SwapSeg nxt = ld16(p.tbl_seg + c);
    auto segment = [&](uint32_t s0, auto full_tag) {
        constexpr bool kFull = decltype(full_tag)::value;
        const SwapSeg sg = nxt;
        if constexpr (kFull) {
            // the next segment's sub-chain starts where this one ends: requested before this segment is decided
            if (s0 + kSwapSeg < S)
                nxt = ld16(p.tbl_seg + (size_t)(s0 / kSwapSeg + 1u) * npb + (sg.cold[kSwapSeg - 1u] - (B - 1u - s0 - kSwapSeg) * npb));
        }
        uint32_t swbits = 0;
#pragma unroll
        for (uint32_t k = 0; k < kSwapSeg; ++k) {
            const uint32_t s = s0 + k;
            if (kFull || s < S) {  // hottest stage first (utils.rs:89-108)
                const double lp_cold = sg.lp[k], wd = (double)sg.w[k];
                const double hi = sg.mid[k] + wd, lo = sg.mid[k] - wd;
                const bool up = carry_lp > hi;  // r < exp(x) by a margin far above any rounding (or x > 0)
                const bool dn = carry_lp < lo;  // r > exp(x) by such a margin
                bool sw = up;
                if (__builtin_expect(!(up | dn), 0)) {
                    // inside the band, r outside (0, 1), or something not finite: the literal expression
                    const uint32_t i = B - 1u - s;
                    const double r = swap_uniform(p, s, sg.cold[k] - (i - 1u) * npb);
                    const double beta1 = __ldg(p.beta + i), beta2 = __ldg(p.beta + i - 1u);
                    const double x = (beta2 - beta1) * (-lp_cold + carry_lp);  // exchange_prob's exponent, utils.rs:63-66
                    const double ex = exp(x);
                    const double ep = ex > 1.0 ? 1.0 : ex;
                    sw = r < ep;  // utils.rs:105 (NaN -> no swap)
                }
                // swap: the cold walker comes up and is final here, the carried walker goes down; else the carried
                // walker stays here and the cold one travels on
                if (valid) store_slot(hot, sw ? sg.cold[k] : carry_src, sw ? lp_cold : carry_lp);
                carry_src = sw ? carry_src : sg.cold[k];
                carry_lp = sw ? carry_lp : lp_cold;
                hot = sg.cold[k];
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
                    p.swapped_out[(size_t)s * npb + (sg.cold[k] - (i - 1u) * npb)] = (swbits >> k) & 1u;
                }
        }
        if (count) {
            // all lanes of the warp ran the same stages: lane k gathers the count of stage k, one atomic instruction per
            // warp and segment, off the carry chain
            unsigned mine = 0;
#pragma unroll
            for (uint32_t k = 0; k < kSwapSeg; ++k) {
                const unsigned m = __ballot_sync(0xffffffffu, (swbits >> k) & 1u);
                if ((tid & 31u) == k) mine = (unsigned)__popc(m);
            }
            const uint32_t k = tid & 31u;
            if (k < kSwapSeg && s0 + k < S && mine) {
                if constexpr (kSmem) atomicAdd(stage_cnt + s0 + k, mine);
                else atomicAdd(p.stage_swapped + (B - 2u - s0 - k), (unsigned long long)mine);
            }
        }
    };

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
        const SwapOut o = ld16(out + slot);
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
