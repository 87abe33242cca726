// hmc_plugin.cu -- instantiates the HMC trajectory kernel for ONE log-density plugin and exports its launcher.
// Compiled once per plugin with -DSCORUS_PLUGIN=<struct name> (see Makefile).
#include "hmc_kernel.cuh"
#include "launch.cuh"
#include "logprob.cuh"

#ifndef SCORUS_PLUGIN
#error "compile with -DSCORUS_PLUGIN=LpGaussian (or another plugin of logprob.cuh)"
#endif
#define SCORUS_CAT2(a, b) a##b
#define SCORUS_CAT(a, b) SCORUS_CAT2(a, b)

namespace scorus {

template <class Plugin, int G, int ITERS>
static cudaError_t launch_one(const HmcParams &hp, uint32_t grid, cudaStream_t st)
{
    hmc_kernel<Plugin, G, ITERS><<<grid, 256, 0, st>>>(hp);
    return cudaGetLastError();
}

cudaError_t SCORUS_CAT(launch_hmc_, SCORUS_PLUGIN)(const HmcParams &hp, uint32_t num_sms, cudaStream_t st)
{
    using Plugin = SCORUS_PLUGIN;
    if (!Plugin::kHasGrad) return cudaErrorNotSupported;
    const uint32_t nvec = hp.s.pitch / 2;
    uint32_t grp = 1;
    while (grp < 32 && grp < nvec) grp <<= 1;
    const uint32_t iters = nvec > 64 ? 4 : (nvec > 32 ? 2 : 1);
    const uint32_t groups_per_block = 8u * (32u / grp);
    uint32_t grid = (hp.s.n + groups_per_block - 1) / groups_per_block;
    const uint32_t cap = num_sms * 8u;  // 256 threads per block: up to 8 blocks per SM
    if (grid > cap) grid = cap;
    if (iters == 4) return launch_one<Plugin, 32, 4>(hp, grid, st);
    if (iters == 2) return launch_one<Plugin, 32, 2>(hp, grid, st);
    switch (grp) {
    case 1: return launch_one<Plugin, 1, 1>(hp, grid, st);
    case 2: return launch_one<Plugin, 2, 1>(hp, grid, st);
    case 4: return launch_one<Plugin, 4, 1>(hp, grid, st);
    case 8: return launch_one<Plugin, 8, 1>(hp, grid, st);
    case 16: return launch_one<Plugin, 16, 1>(hp, grid, st);
    default: return launch_one<Plugin, 32, 1>(hp, grid, st);
    }
}

}  // namespace scorus
