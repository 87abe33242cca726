// swap_params.cuh -- argument block of the ladder-swap kernels (swap_kernel.cuh), separate from them so that
// translation units which only fill it in need not instantiate the kernels.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace scorus {

constexpr uint32_t kSwapSeg = 8;  // stages per segment of the ladder (swap_kernel.cuh)

// What the sub-chain that starts at hot slot x of a segment's first stage meets over the segment's stages, one
// contiguous block per (segment, x): the sweep fetches a whole segment of a chain with consecutive 16-byte loads of one
// or two cache lines.  Per stage: the carried walker swaps with the cold slot's iff r < min(1, exp(db (lp_carry -
// lp_cold))), db = beta_cold - beta_hot (utils.rs:63-69,105)  <=>  lp_carry > lp_cold + ln r / db = mid, so the sweep
// compares the carried log-prob with mid -/+ w and evaluates the literal expression only inside that band; w covers
// the rounding of everything on both sides by orders of magnitude (swap_paths_kernel).
struct __align__(16) SwapSeg {
    double mid[kSwapSeg];     // NaN: decide literally (r outside (0, 1), db <= 0, non-finite log-prob)
    double lp[kSwapSeg];      // log-prob of the cold slot
    float w[kSwapSeg];        // half-width of the literal band around mid
    uint32_t cold[kSwapSeg];  // global index of the cold slot
};

// interleaved result of a swap, [n] indexed by slot: one 16-byte scattered write per slot instead of two, and the step
// that follows reads a walker's log-prob and its row's whereabouts with one load (step_reg_kernel.cuh, kRemap)
struct __align__(16) SwapOut {
    double lp;     // log-prob of the walker that ends in this slot
    uint32_t src;  // slot its row is in before the rows move
    uint32_t pad;
};

// the records move as 16-byte accesses (the compiler splits a struct copy into 8-byte ones, and a scattered access costs
// the load/store unit one slot per lane whatever its width)
#ifdef __CUDACC__
template <class R>
__device__ __forceinline__ R ld16(const R *p)
{
    static_assert(sizeof(R) % 16 == 0, "16-byte records");
    constexpr int NQ = sizeof(R) / 16;
    int4 v[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) v[q] = reinterpret_cast<const int4 *>(p)[q];
    R o;
    memcpy(&o, v, sizeof(R));
    return o;
}
template <class R>
__device__ __forceinline__ void st16(R *p, const R &o)
{
    static_assert(sizeof(R) % 16 == 0, "16-byte records");
    constexpr int NQ = sizeof(R) / 16;
    int4 v[NQ];
    memcpy(v, &o, sizeof(R));
#pragma unroll
    for (int q = 0; q < NQ; ++q) reinterpret_cast<int4 *>(p)[q] = v[q];
}
#endif

struct SwapParams {
    SwapOut *out;          // [n] out, interleaved -- or nullptr: results go to the two arrays lp / src below
    double *lp;            // [n] out: log-prob of the walker that ends in each slot (may be lp_src: in place)
    const double *lp_src;  // [n] in: log-probs before the swap
    const double *beta;    // [nbeta]
    const uint32_t *jinv;  // [(nbeta-1)][npb] cold slot of each hot slot, or nullptr -> bijection
    const double *r_in;    // [(nbeta-1)][npb] indexed by cold slot, or nullptr -> Philox
    uint32_t *src;         // [n] out: original slot of the walker that ends in each slot
    uint8_t *swapped_out;  // [(nbeta-1)][npb] indexed by cold slot, or nullptr
    unsigned long long *stats;  // [2] swapped, [3] swap_proposed
    unsigned long long *stage_swapped;  // [nbeta-1] accepted exchanges between rung k and k+1, or nullptr
    // scratch of the swap, [segments][npb], indexed [segment][hot slot x of the segment's first stage], written by
    // swap_paths_kernel for swap_walk_kernel
    SwapSeg *tbl_seg;
    uint32_t debug_direct_store;  // SCORUS_SWAP_DIRECT_STORE=1: swap_paths_kernel skips its shared-memory stage
    uint64_t seed, swap_call;
    uint32_t npb, nbeta, bij_bits;
};

}  // namespace scorus
