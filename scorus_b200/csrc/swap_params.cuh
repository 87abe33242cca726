// swap_params.cuh -- argument block of the ladder-swap kernels (swap_kernel.cuh), separate from them so that
// translation units which only fill it in need not instantiate the kernels.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace scorus {

// what the sub-chain starting at hot slot x meets at one stage: one 16-byte scattered read per stage of the sweep
struct __align__(16) SwapRec {
    double lp;      // log-prob of the cold slot
    float lr;       // ln r of its swap uniform, rounded to f32 (the guard band of the decision covers the rounding); NaN: decide literally
    uint32_t cold;  // global index of the cold slot
};

// interleaved result of a swap, [n] indexed by slot: one 16-byte scattered write per slot instead of two, and the step
// that follows reads a walker's log-prob and its row's whereabouts with one load (step_reg_kernel.cuh, kRemap)
struct __align__(16) SwapOut {
    double lp;     // log-prob of the walker that ends in this slot
    uint32_t src;  // slot its row is in before the rows move
    uint32_t pad;
};

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
    // scratch of the swap, [(nbeta-1)][npb], indexed [stage][hot slot x of the stage's segment start], written by
    // swap_paths_kernel for swap_walk_kernel
    SwapRec *tbl_rec;
    uint64_t seed, swap_call;
    uint32_t npb, nbeta, bij_bits;
};

}  // namespace scorus
