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


// The sweep is latency, not work: 16384 chains are 512 warps -- a few per SM -- so every instruction of a chain's
// thread is paid at its full latency, and the first version (keys, bijection, a DRAM-latency log-prob load, a Philox
// draw and an exp per stage, all in one dependent loop) took 76 us at 64 rungs x 16384 chains, as long as a step
// (profiles/r2_c4_swap_chain_ncu_summary.txt; a two-pass split that still drew and stored per stage: 92 us, stalled on
// the stores waiting for their loads).  What has to be sequential is only (a) which slots the chain visits, a function
// of the permutations, and (b) the carry of the travelling walker.  Everything else is hoisted out:
//   * swap_uniforms_kernel, one thread per (stage, cold slot): the swap uniform r and ln r, fully parallel;
//   * pass 1 of the chain thread: the slot chain, integer arithmetic only (bijection keys of every stage in shared
//     memory), and per stage three cp.async copies -- cold log-prob, ln r -- into the thread's column of a
//     shared-memory table: asynchronous, nothing waits on them;
//   * pass 2: the decision r < min(1, exp(x)), x = (beta_cold - beta_hot)(lp_carry - lp_cold), utils.rs:63-69,105, taken
//     WITHOUT the exp wherever ln r and x are not within 1e-9 of each other: ln r < x - d  =>  r < exp(x) (1 - 1e-9...),
//     far outside the rounding of either side, so the shortcut returns exactly what the exp comparison returns; the
//     one-in-10^9 close calls, r = 0 and non-finite x take the literal path.  The dependent chain per stage is an add,
//     a multiply and two compares.
constexpr uint32_t kSwapChunk = 64;  // stages per pass (shared-memory table: kSwapChunk x 128 threads x 20 bytes)

__global__ void __launch_bounds__(256) swap_uniforms_kernel(const SwapParams p)
{
    const size_t m = (size_t)(p.nbeta - 1) * p.npb;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < m; e += (size_t)gridDim.x * blockDim.x) {
        const uint32_t s = (uint32_t)(e / p.npb), j2 = (uint32_t)(e - (size_t)s * p.npb), i = p.nbeta - 1 - s;
        double r;
        if (p.r_in) {
            r = __ldg(p.r_in + e);
        } else {
            const U4 o = draw(p.seed, p.swap_call, ST_SWAP_U, j2, i);
            r = to_unit<double>(o.x, o.y);
        }
        p.chain_r[e] = r;
        // ln r for the exp-free decision; anything but 0 < r < 1 (only a caller-supplied r can be) is marked NaN and
        // decided by the literal expression
        p.chain_lr[e] = (r > 0.0 && r < 1.0) ? log(r) : real_nan<double>();
    }
}

// asynchronous copy of one scalar (BYTES = its size), global -> shared
template <int BYTES>
__device__ __forceinline__ void cp_async_scalar(void *smem_dst, const void *gsrc)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc),
                 "n"(BYTES)
                 : "memory");
}

__global__ void __launch_bounds__(128) swap_chain_kernel(const SwapParams p)
{
    extern __shared__ __align__(16) unsigned char swap_smem[];
    const uint32_t B = p.nbeta, npb = p.npb, T = blockDim.x, tid = threadIdx.x;
    // table columns of this thread: [stage in chunk][thread]
    double *t_lp = reinterpret_cast<double *>(swap_smem);
    double *t_lr = t_lp + (size_t)kSwapChunk * T;
    uint32_t *t_j2 = reinterpret_cast<uint32_t *>(t_lr + (size_t)kSwapChunk * T);
    uint32_t *stage_keys = t_j2 + (size_t)kSwapChunk * T;  // [(nbeta-1)][8], bijection mode only
    if (!p.jinv) {
        for (uint32_t i = 1 + tid; i < B; i += T) {
            const BijKeys K = bij_keys(p.seed, p.swap_call, ST_SWAP_PERM, i);
#pragma unroll
            for (int q = 0; q < 8; ++q) stage_keys[(i - 1) * 8 + q] = K.k[q];
        }
        __syncthreads();
    }
    const uint32_t c = blockIdx.x * T + tid;
    unsigned long long nsw = 0;
    if (c < npb) {
        uint32_t a1 = c, a2 = c;  // hot slot of the chain in pass 1 / pass 2
        uint32_t carry_src = (B - 1) * npb + c;
        double carry_lp = p.lp_src[carry_src];
        for (uint32_t s0 = 0; s0 + 1 < B; s0 += kSwapChunk) {
            const uint32_t ns = min(kSwapChunk, B - 1 - s0);
            // ---- pass 1: the chain's slots; log-probs and ln r on their way into the table ----
            for (uint32_t k = 0; k < ns; ++k) {
                const uint32_t s = s0 + k, i = B - 1 - s;
                uint32_t j2;
                if (p.jinv) {
                    j2 = __ldg(p.jinv + (size_t)s * npb + a1);
                } else {
                    BijKeys K;
#pragma unroll
                    for (int q = 0; q < 8; ++q) K.k[q] = stage_keys[(i - 1) * 8 + q];
                    j2 = bij(a1, npb, p.bij_bits, K);
                }
                t_j2[k * T + tid] = j2;
                cp_async_scalar<sizeof(double)>(t_lp + k * T + tid, p.lp_src + (size_t)(i - 1) * npb + j2);
                cp_async_scalar<sizeof(double)>(t_lr + k * T + tid, p.chain_lr + (size_t)s * npb + j2);
                a1 = j2;
            }
            asm volatile("cp.async.wait_all;" ::: "memory");
            // ---- pass 2: decisions, hottest stage first (utils.rs:89-108) ----
            for (uint32_t k = 0; k < ns; ++k) {
                const uint32_t s = s0 + k, i = B - 1 - s;
                const uint32_t j2 = t_j2[k * T + tid];
                const double lp_cold = t_lp[k * T + tid], lr = t_lr[k * T + tid];
                const uint32_t hot = i * npb + a2, cold = (i - 1) * npb + j2;
                const double beta1 = __ldg(p.beta + i), beta2 = __ldg(p.beta + i - 1);
                const double x = (beta2 - beta1) * (-lp_cold + carry_lp);  // exchange_prob's exponent, utils.rs:63-66
                const double d = swap_guard<double>() * (1.0 + fabs(x));
                bool sw;
                if (lr == lr && x >= d) {
                    sw = true;                       // exp(x) > 1: probability 1, and 0 < r < 1
                } else if (lr < x - d) {
                    sw = true;                       // r < exp(x) < 1 by a margin far above any rounding
                } else if (lr > x + d) {
                    sw = false;
                } else {                             // too close to call, r outside (0, 1), or x not finite: the literal expression
                    const double r = p.chain_r[(size_t)s * npb + j2];
                    const double ex = exp(x);
                    const double ep = ex > 1.0 ? 1.0 : ex;
                    sw = r < ep;                     // utils.rs:105 (NaN -> no swap)
                }
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
                    if ((tid & 31u) == (uint32_t)(__ffs(live) - 1) && m)
                        atomicAdd(p.stage_swapped + (i - 1), (unsigned long long)__popc(m));
                }
                a2 = j2;
            }
        }
        p.src[a2] = carry_src;
        p.lp[a2] = carry_lp;
    }
    for (int off = 16; off; off >>= 1) nsw += __shfl_xor_sync(0xffffffffu, nsw, off);
    if ((tid & 31u) == 0 && nsw) atomicAdd(p.stats + 2, nsw);
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
