// swap_kernel.cuh -- K4 (exact-shuffle order generator) and K5 (temperature-ladder swap).
//
// swap_walkers (src/mcmc/utils.rs:72-112) sweeps the rungs hottest -> coldest; stage i pairs
// slot j1 of rung i with slot j2 of rung i-1 through a permutation jvec_i (j1 = jvec_i[j2]) and
// swaps the two walkers with probability min(1, exp((beta_{i-1}-beta_i)(lp_hot - lp_cold))).
// The B-1 stages look sequential, but the slots a stage touches are linked only along CHAINS:
// slot (i-1, j2) is next touched by stage i-1 as its hot slot.  Following
// a_{i-1} = jinv_i[a_i] from a_{B-1} = c gives npb disjoint chains of B slots, each swept by one
// thread with the travelling walker's (log-prob, origin) carried in registers -- no inter-block
// synchronisation, B-1 dependent steps per thread, npb-way parallel.  Decisions depend only on
// log-probs, so the sweep produces a source map src[slot] and the final log-probs first; rows
// move once afterwards (gather into scratch, copy back), only the slots that changed.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

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


__global__ void __launch_bounds__(128) swap_chain_kernel(const SwapParams p)
{
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long nsw = 0;
    if (c < p.npb) {
        const uint32_t B = p.nbeta, npb = p.npb;
        uint32_t a = c;
        uint32_t carry_src = (B - 1) * npb + a;
        double carry_lp = p.lp[carry_src];
        for (uint32_t i = B - 1; i >= 1; --i) {
            const uint32_t s = B - 1 - i;
            uint32_t j2;
            if (p.jinv) {
                j2 = __ldg(p.jinv + (size_t)s * npb + a);
            } else {
                BijKeys K = bij_keys(p.seed, p.swap_call, ST_SWAP_PERM, i);
                j2 = bij(a, npb, p.bij_bits, K);
            }
            const uint32_t hot = i * npb + a, cold = (i - 1) * npb + j2;
            const double lp_cold = p.lp[cold];
            const double beta1 = __ldg(p.beta + i), beta2 = __ldg(p.beta + i - 1);
            // exchange_prob, src/mcmc/utils.rs:63-69
            double x = exp((beta2 - beta1) * (-lp_cold + carry_lp));
            const double ep = x > 1.0 ? 1.0 : x;
            double r;
            if (p.r_in) {
                r = __ldg(p.r_in + (size_t)s * npb + j2);
            } else {
                U4 o = draw(p.seed, p.swap_call, ST_SWAP_U, j2, i);
                r = u53(o.x, o.y);
            }
            const bool sw = r < ep;  // utils.rs:105 (NaN -> no swap)
            if (sw) {
                // the cold walker comes up and is final here; the carried walker goes down
                p.src[hot] = cold;
                p.lp[hot] = lp_cold;
                ++nsw;
            } else {
                p.src[hot] = carry_src;
                p.lp[hot] = carry_lp;
                carry_src = cold;
                carry_lp = lp_cold;
            }
            if (p.swapped_out) p.swapped_out[(size_t)s * npb + j2] = sw ? 1 : 0;
            if (p.stage_swapped) {
                // every active lane of the warp runs the same B-1 stages: one atomic per warp per stage
                const unsigned live = __activemask();
                const unsigned m = __ballot_sync(live, sw);
                if ((threadIdx.x & 31u) == (uint32_t)(__ffs(live) - 1) && m)
                    atomicAdd(p.stage_swapped + (i - 1), (unsigned long long)__popc(m));
            }
            a = j2;
        }
        p.src[a] = carry_src;
        p.lp[a] = carry_lp;
    }
    for (int off = 16; off; off >>= 1) nsw += __shfl_xor_sync(0xffffffffu, nsw, off);
    if ((threadIdx.x & 31u) == 0 && nsw) atomicAdd(p.stats + 2, nsw);
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
