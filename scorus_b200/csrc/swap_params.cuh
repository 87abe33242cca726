// swap_params.cuh -- argument block of the ladder-swap kernels (swap_kernel.cuh), separate from them so that
// translation units which only fill it in need not instantiate the kernels.
#pragma once
#include <cstdint>

namespace scorus {

struct SwapParams {
    double *lp;            // [n] out: log-prob of the walker that ends in each slot (single GPU: in place, = lp_src)
    const double *lp_src;  // [n] in: log-probs before the swap
    const double *beta;    // [nbeta]
    const uint32_t *jinv;  // [(nbeta-1)][npb] cold slot of each hot slot, or nullptr -> bijection
    const double *r_in;    // [(nbeta-1)][npb] indexed by cold slot, or nullptr -> Philox
    uint32_t *src;         // [n] out: original slot of the walker that ends in each slot
    uint8_t *swapped_out;  // [(nbeta-1)][npb] indexed by cold slot, or nullptr
    unsigned long long *stats;  // [2] swapped, [3] swap_proposed
    unsigned long long *stage_swapped;  // [nbeta-1] accepted exchanges between rung k and k+1, or nullptr
    // scratch of the swap, [(nbeta-1)][npb] each, indexed [stage][cold slot]: the swap uniform and its logarithm,
    // written by swap_uniforms_kernel for the chain kernel
    double *chain_r, *chain_lr;
    uint64_t seed, swap_call;
    uint32_t npb, nbeta, bij_bits;
};

}  // namespace scorus
