// launch.cuh -- kernel launchers, one translation unit per log-density plugin (launch_plugin.cu
// compiled with -DSCORUS_PLUGIN=...), so the 60-odd kernel instantiations build in parallel.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

// Shape of the register-streaming stretch kernel (step_reg_kernel.cuh), shared with the host geometry.  Rows wider
// than 16 double2 and up to 64 (G = 32, one or two chunks per lane) run 4 rows per software-pipeline block under a 96-register
// cap, 20 warps per SM: the kernel is latency-bound and more resident warps pay (2^22 x 64-D: 0.64 -> 0.61 ms);
// narrower rows keep 8 rows per block and 16 warps (10-D lost 5 % with the tighter cap).
#define SCORUS_REG_ROWS_WIDE 4
#define SCORUS_REG_WARPS_WIDE 20
#ifndef SCORUS_REG_ROWS
#define SCORUS_REG_ROWS 8
#endif
#ifndef SCORUS_REG_WARPS
#define SCORUS_REG_WARPS 16
#endif
// rows of 9..16 double2 (16 lanes per row, tile mode): 1 = the wide shape (4 rows per block, 20 warps) for them too
#ifndef SCORUS_G16_DIRECT
#define SCORUS_G16_DIRECT 0
#endif
#ifndef SCORUS_G16_WIDE
#define SCORUS_G16_WIDE 1
#endif

namespace scorus {

struct StepParams;
struct TwalkParams;
struct HmcParams;

enum : int { KERNEL_SEQ = 0, KERNEL_COOP = 1, KERNEL_REG = 2, KERNEL_SMALL = 3 };

struct Geometry {
    uint32_t row_bytes, row_stride, flag_words, stages, warps, grid;
    size_t smem;
    int kernel;  // KERNEL_*
    int pdl;     // launch with programmatic stream serialization (register kernel only)
};

#define SCORUS_DECLARE_LAUNCHERS(NAME)                                                                   \
    cudaError_t launch_step_##NAME(const StepParams &p, const Geometry &g, bool all_flags, \
                                   cudaStream_t st);                                                     \
    cudaError_t launch_init_##NAME(const StepParams &p, uint32_t grid, size_t smem, cudaStream_t st);         \
    cudaError_t launch_small_##NAME(const StepParams &p, const Geometry &g, uint32_t inner_steps, cudaStream_t st); \
    cudaError_t launch_twalk_##NAME(const TwalkParams &tp, const Geometry &g, cudaStream_t st);             \
    cudaError_t launch_hmc_##NAME(const HmcParams &hp, uint32_t num_sms, cudaStream_t st);

SCORUS_DECLARE_LAUNCHERS(LpGaussian)
SCORUS_DECLARE_LAUNCHERS(LpBoundedGaussian)
SCORUS_DECLARE_LAUNCHERS(LpRosenbrock)
SCORUS_DECLARE_LAUNCHERS(LpRastrigin)
SCORUS_DECLARE_LAUNCHERS(LpBimodal2d)
SCORUS_DECLARE_LAUNCHERS(LpStep2d)
SCORUS_DECLARE_LAUNCHERS(LpCorrGaussian)
SCORUS_DECLARE_LAUNCHERS(LpMixture)

// whether the stretch step of this geometry has the row-moving variant (step_reg_kernel.cuh kRemap: register kernel,
// direct mode -- every target but the dense one, rows outside 9..16 chunks); mirrors launch_step's dispatch
bool step_can_remap(int kind, uint32_t pitch, const Geometry &g);

// dispatch on the plugin id of include/scorus_b200.h (SCORUS_LP_*)
cudaError_t launch_step(int kind, const StepParams &p, const Geometry &g, bool all_flags, cudaStream_t st);
cudaError_t launch_init(int kind, const StepParams &p, uint32_t grid, size_t smem, cudaStream_t st);
// one-block ensembles: inner_steps consecutive steps in one launch (step_small_kernel.cuh)
cudaError_t launch_small(int kind, const StepParams &p, const Geometry &g, uint32_t inner_steps, cudaStream_t st);
cudaError_t launch_twalk(int kind, const TwalkParams &tp, const Geometry &g, cudaStream_t st);
cudaError_t launch_hmc(int kind, const HmcParams &hp, uint32_t num_sms, cudaStream_t st);  // cudaErrorNotSupported: no gradient

}  // namespace scorus
