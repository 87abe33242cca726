// launch_dispatch.cu -- plugin id (SCORUS_LP_*) -> per-plugin launcher.
#include "../../include/scorus_b200.h"
#include "launch.cuh"

namespace scorus {

bool step_can_remap(int kind, uint32_t pitch, const Geometry &g)
{
    if (g.kernel != KERNEL_REG || kind == SCORUS_LP_CORR_GAUSSIAN) return false;  // kNeedsRow: tile mode
    const uint32_t nvec = pitch / 2;
    return SCORUS_G16_DIRECT || nvec < 9 || nvec > 16;  // 16 lanes per row: tile mode (launch_plugin.cu, reg_kernel_direct)
}

cudaError_t launch_step(int kind, const StepParams &p, const Geometry &g, bool all_flags, cudaStream_t st)
{
    switch (kind) {
    case SCORUS_LP_GAUSSIAN: return launch_step_LpGaussian(p, g, all_flags, st);
    case SCORUS_LP_BOUNDED_GAUSSIAN: return launch_step_LpBoundedGaussian(p, g, all_flags, st);
    case SCORUS_LP_ROSENBROCK: return launch_step_LpRosenbrock(p, g, all_flags, st);
    case SCORUS_LP_RASTRIGIN: return launch_step_LpRastrigin(p, g, all_flags, st);
    case SCORUS_LP_BIMODAL2D: return launch_step_LpBimodal2d(p, g, all_flags, st);
    case SCORUS_LP_STEP2D: return launch_step_LpStep2d(p, g, all_flags, st);
    case SCORUS_LP_CORR_GAUSSIAN: return launch_step_LpCorrGaussian(p, g, all_flags, st);
    case SCORUS_LP_MIXTURE: return launch_step_LpMixture(p, g, all_flags, st);
    default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_init(int kind, const StepParams &p, uint32_t grid, size_t smem, cudaStream_t st)
{
    switch (kind) {
    case SCORUS_LP_GAUSSIAN: return launch_init_LpGaussian(p, grid, smem, st);
    case SCORUS_LP_BOUNDED_GAUSSIAN: return launch_init_LpBoundedGaussian(p, grid, smem, st);
    case SCORUS_LP_ROSENBROCK: return launch_init_LpRosenbrock(p, grid, smem, st);
    case SCORUS_LP_RASTRIGIN: return launch_init_LpRastrigin(p, grid, smem, st);
    case SCORUS_LP_BIMODAL2D: return launch_init_LpBimodal2d(p, grid, smem, st);
    case SCORUS_LP_STEP2D: return launch_init_LpStep2d(p, grid, smem, st);
    case SCORUS_LP_CORR_GAUSSIAN: return launch_init_LpCorrGaussian(p, grid, smem, st);
    case SCORUS_LP_MIXTURE: return launch_init_LpMixture(p, grid, smem, st);
    default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_twalk(int kind, const TwalkParams &tp, const Geometry &g, cudaStream_t st)
{
    switch (kind) {
    case SCORUS_LP_GAUSSIAN: return launch_twalk_LpGaussian(tp, g, st);
    case SCORUS_LP_BOUNDED_GAUSSIAN: return launch_twalk_LpBoundedGaussian(tp, g, st);
    case SCORUS_LP_ROSENBROCK: return launch_twalk_LpRosenbrock(tp, g, st);
    case SCORUS_LP_RASTRIGIN: return launch_twalk_LpRastrigin(tp, g, st);
    case SCORUS_LP_BIMODAL2D: return launch_twalk_LpBimodal2d(tp, g, st);
    case SCORUS_LP_STEP2D: return launch_twalk_LpStep2d(tp, g, st);
    case SCORUS_LP_CORR_GAUSSIAN: return launch_twalk_LpCorrGaussian(tp, g, st);
    case SCORUS_LP_MIXTURE: return launch_twalk_LpMixture(tp, g, st);
    default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_hmc(int kind, const HmcParams &hp, uint32_t num_sms, cudaStream_t st)
{
    switch (kind) {
    case SCORUS_LP_GAUSSIAN: return launch_hmc_LpGaussian(hp, num_sms, st);
    case SCORUS_LP_BOUNDED_GAUSSIAN: return launch_hmc_LpBoundedGaussian(hp, num_sms, st);
    case SCORUS_LP_ROSENBROCK: return launch_hmc_LpRosenbrock(hp, num_sms, st);
    case SCORUS_LP_RASTRIGIN: return launch_hmc_LpRastrigin(hp, num_sms, st);
    case SCORUS_LP_BIMODAL2D: return launch_hmc_LpBimodal2d(hp, num_sms, st);
    case SCORUS_LP_STEP2D: return launch_hmc_LpStep2d(hp, num_sms, st);
    case SCORUS_LP_CORR_GAUSSIAN: return launch_hmc_LpCorrGaussian(hp, num_sms, st);
    case SCORUS_LP_MIXTURE: return launch_hmc_LpMixture(hp, num_sms, st);
    default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_small(int kind, const StepParams &p, const Geometry &g, uint32_t inner_steps, cudaStream_t st)
{
    switch (kind) {
    case SCORUS_LP_GAUSSIAN: return launch_small_LpGaussian(p, g, inner_steps, st);
    case SCORUS_LP_BOUNDED_GAUSSIAN: return launch_small_LpBoundedGaussian(p, g, inner_steps, st);
    case SCORUS_LP_ROSENBROCK: return launch_small_LpRosenbrock(p, g, inner_steps, st);
    case SCORUS_LP_RASTRIGIN: return launch_small_LpRastrigin(p, g, inner_steps, st);
    case SCORUS_LP_BIMODAL2D: return launch_small_LpBimodal2d(p, g, inner_steps, st);
    case SCORUS_LP_STEP2D: return launch_small_LpStep2d(p, g, inner_steps, st);
    case SCORUS_LP_CORR_GAUSSIAN: return launch_small_LpCorrGaussian(p, g, inner_steps, st);
    case SCORUS_LP_MIXTURE: return launch_small_LpMixture(p, g, inner_steps, st);
    default: return cudaErrorInvalidValue;
    }
}

}  // namespace scorus
