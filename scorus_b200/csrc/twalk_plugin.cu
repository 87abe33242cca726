// twalk_plugin.cu -- instantiates the t-walk half-pass kernel for ONE log-density plugin and exports its
// launcher.  Compiled once per plugin with -DSCORUS_PLUGIN=<struct name> (see Makefile).
#include "launch.cuh"
#include "twalk_direct_kernel.cuh"
#include "twalk_kernel.cuh"

#ifndef SCORUS_PLUGIN
#error "compile with -DSCORUS_PLUGIN=LpGaussian (or another plugin of logprob.cuh)"
#endif
#define SCORUS_CAT2(a, b) a##b
#define SCORUS_CAT(a, b) SCORUS_CAT2(a, b)

namespace scorus {

template <class K>
static cudaError_t launch_twalk_kernel(K kernel, const TwalkParams &tp, const Geometry &g, cudaStream_t st)
{
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem);
    if (e != cudaSuccess) return e;
    if (g.pdl) {  // as the stretch-move kernel: blocks are scheduled early and wait in griddepcontrol.wait
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(g.grid);
        cfg.blockDim = dim3(g.warps * 32);
        cfg.dynamicSmemBytes = g.smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        return cudaLaunchKernelEx(&cfg, kernel, tp);
    }
    kernel<<<g.grid, g.warps * 32, g.smem, st>>>(tp);
    return cudaGetLastError();
}

// a template so that `if constexpr` really discards the register kernel for the dense plugin
template <class Plugin>
static cudaError_t launch_twalk_impl(const TwalkParams &tp, const Geometry &g, cudaStream_t st)
{
    const uint32_t nvec = tp.s.pitch / 2;
    if constexpr (!Plugin::kNeedsRow) {
        if (tp.direct && !(nvec > 8 && nvec <= 16)) {  // mirrors reg_kernel_direct (step_reg_kernel.cuh)
            if (nvec == 5 || nvec == 6) return launch_twalk_kernel(twalk_direct_kernel<Plugin, 2, 3>, tp, g, st);
            if (nvec > 64) return launch_twalk_kernel(twalk_direct_kernel<Plugin, 32, 4>, tp, g, st);
            if (nvec > 32) return launch_twalk_kernel(twalk_direct_kernel<Plugin, 32, 2>, tp, g, st);
            if (nvec > 16) return launch_twalk_kernel(twalk_direct_kernel<Plugin, 32, 1>, tp, g, st);
            if (nvec > 4) return launch_twalk_kernel(twalk_direct_kernel<Plugin, 8, 1>, tp, g, st);
            if (nvec > 2) return launch_twalk_kernel(twalk_direct_kernel<Plugin, 4, 1>, tp, g, st);
            return launch_twalk_kernel(twalk_direct_kernel<Plugin, 2, 1>, tp, g, st);
        }
    }
    if (nvec == 5 || nvec == 6) return launch_twalk_kernel(twalk_half_kernel<Plugin, 2, 3>, tp, g, st);
    if (nvec > 64) return launch_twalk_kernel(twalk_half_kernel<Plugin, 32, 4>, tp, g, st);
    if (nvec > 32) return launch_twalk_kernel(twalk_half_kernel<Plugin, 32, 2>, tp, g, st);
    uint32_t grp = 2;
    while (grp < 32 && grp < nvec) grp <<= 1;
    switch (grp) {
    case 2: return launch_twalk_kernel(twalk_half_kernel<Plugin, 2, 1>, tp, g, st);
    case 4: return launch_twalk_kernel(twalk_half_kernel<Plugin, 4, 1>, tp, g, st);
    case 8: return launch_twalk_kernel(twalk_half_kernel<Plugin, 8, 1>, tp, g, st);
    case 16: return launch_twalk_kernel(twalk_half_kernel<Plugin, 16, 1>, tp, g, st);
    default: return launch_twalk_kernel(twalk_half_kernel<Plugin, 32, 1>, tp, g, st);
    }
}

cudaError_t SCORUS_CAT(launch_twalk_, SCORUS_PLUGIN)(const TwalkParams &tp, const Geometry &g, cudaStream_t st)
{
    return launch_twalk_impl<SCORUS_PLUGIN>(tp, g, st);
}

}  // namespace scorus
