// launch_plugin.cu -- instantiates the step / init kernels for ONE log-density plugin and exports
// their launchers.  Compiled once per plugin with -DSCORUS_PLUGIN=<struct name> (see Makefile).
#include "launch.cuh"
#include "step_coop_kernel.cuh"
#include "step_kernel.cuh"
#include "step_reg_kernel.cuh"
#ifndef SCORUS_F32  // the f32 instantiation (gen_f32.py) covers the register kernel, init_logprob and the swap
#include "step_small_kernel.cuh"
#endif

#ifndef SCORUS_PLUGIN
#error "compile with -DSCORUS_PLUGIN=LpGaussian (or another plugin of logprob.cuh)"
#endif
#define SCORUS_CAT2(a, b) a##b
#define SCORUS_CAT(a, b) SCORUS_CAT2(a, b)

namespace scorus {

template <class K>
static cudaError_t launch_kernel(K kernel, const StepParams &p, const Geometry &g, cudaStream_t st)
{
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem);
    if (e != cudaSuccess) return e;
    if (g.pdl) {
        // programmatic dependent launch: this step's blocks are scheduled while the previous launch
        // drains and block in griddepcontrol.wait before touching its results (step_reg_kernel.cuh)
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
        return cudaLaunchKernelEx(&cfg, kernel, p);
    }
    kernel<<<g.grid, g.warps * 32, g.smem, st>>>(p);
    return cudaGetLastError();
}

// lanes sharing a pair in the cooperative kernel: the power of two covering a row's double2 count
static int coop_group(uint32_t pitch)
{
    uint32_t nvec = pitch / 2, g = 2;
    while (g < 32 && g < nvec) g <<= 1;
    return (int)g;
}

// the register kernel in its three flavours: single GPU / gathered copies (0), one-sided with both rows through
// registers (1), one-sided with the partner rows staged by TMA (2: direct mode, when the host laid out the stage)
template <class Plugin, int G, int ITERS>
static cudaError_t launch_reg(const StepParams &p, const Geometry &g, cudaStream_t st)
{
    if (p.peer_n) {
        if constexpr (reg_kernel_direct<Plugin, G>()) {
            if (p.peer_stage) return launch_kernel(step_reg_kernel<Plugin, G, ITERS, 2>, p, g, st);
        }
        return launch_kernel(step_reg_kernel<Plugin, G, ITERS, 1>, p, g, st);
    }
#ifndef SCORUS_F32
    if (p.remap || p.host_ens) {  // the step after a ladder swap / the first step of a host-buffer call (the host checked step_can_remap)
        if constexpr (reg_kernel_direct<Plugin, G>()) return launch_kernel(step_reg_kernel<Plugin, G, ITERS, 0, true>, p, g, st);
        else return cudaErrorNotSupported;
    }
#endif
    return launch_kernel(step_reg_kernel<Plugin, G, ITERS, 0>, p, g, st);
}

cudaError_t SCORUS_CAT(launch_step_, SCORUS_PLUGIN)(const StepParams &p, const Geometry &g, bool all_flags,
                                                    cudaStream_t st)
{
    using Plugin = SCORUS_PLUGIN;
#ifdef SCORUS_F32
    if (g.kernel != KERNEL_REG) return cudaErrorNotSupported;
#else
    if (g.kernel == KERNEL_SEQ) {
        if (all_flags) return launch_kernel(step_seq_kernel<Plugin, true>, p, g, st);
        return launch_kernel(step_seq_kernel<Plugin, false>, p, g, st);
    }
#endif
    const uint32_t nvec = p.pitch / 2;
    if (g.kernel == KERNEL_REG) {
        // rows of 5 or 6 double2 (D = 9..12): two lanes x three chunks wastes 1/6 of the lanes, eight lanes
        // x one chunk would waste 3/8 -- and these targets (10-D Rastrigin) are bound by fp64 arithmetic
        if (nvec == 5 || nvec == 6) return launch_reg<Plugin, 2, 3>(p, g, st);
        if (nvec > 64) return launch_reg<Plugin, 32, 4>(p, g, st);
        if (nvec > 32) return launch_reg<Plugin, 32, 2>(p, g, st);
        switch (coop_group(p.pitch)) {
        case 2: return launch_reg<Plugin, 2, 1>(p, g, st);
        case 4: return launch_reg<Plugin, 4, 1>(p, g, st);
        case 8: return launch_reg<Plugin, 8, 1>(p, g, st);
        case 16: return launch_reg<Plugin, 16, 1>(p, g, st);
        default: return launch_reg<Plugin, 32, 1>(p, g, st);
        }
    }
#ifdef SCORUS_F32
    return cudaErrorNotSupported;
#else
    // cooperative kernel with the old rows staged by TMA: any dimension that fits a tile
    if (nvec >= 32) return launch_kernel(step_coop_kernel<Plugin, 32>, p, g, st);
    return launch_kernel(step_coop_kernel<Plugin, 8>, p, g, st);
#endif
}

cudaError_t SCORUS_CAT(launch_init_, SCORUS_PLUGIN)(const StepParams &p, uint32_t grid, size_t smem, cudaStream_t st)
{
    using Plugin = SCORUS_PLUGIN;
    cudaError_t e = cudaFuncSetAttribute(init_logprob_kernel<Plugin>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    init_logprob_kernel<Plugin><<<grid, p.warps * 32, smem, st>>>(p);
    return cudaGetLastError();
}

#ifndef SCORUS_F32
cudaError_t SCORUS_CAT(launch_small_, SCORUS_PLUGIN)(const StepParams &p, const Geometry &g, uint32_t inner_steps,
                                                     cudaStream_t st)
{
    using Plugin = SCORUS_PLUGIN;
    cudaError_t e = cudaFuncSetAttribute(step_small_kernel<Plugin>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem);
    if (e != cudaSuccess) return e;
    step_small_kernel<Plugin><<<1, g.warps * 32, g.smem, st>>>(p, inner_steps);
    return cudaGetLastError();
}
#endif

}  // namespace scorus
