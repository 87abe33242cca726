"""scorus_b200 -- B200-native drop-in for scorus's ensemble-sampler hot path.

Only what the path needs: csrc/ (CUDA kernels + the C ABI of include/scorus_b200.h) and the
host-side mirror of the reference modules `mcmc::ensemble_sample`, `mcmc::utils` and (ensemble entry point
only) `mcmc::twalk` and `mcmc::hmc::naive`.
"""
from . import ensemble_sample, hmc, init_ensemble, linear_space, pso, twalk, utils  # noqa: F401
from .ensemble_sample import (  # noqa: F401
    BoundedGaussian, Bimodal2d, CorrGaussian, DeviceRng, Gaussian, LogProb, McmcErr, Mixture, Rastrigin,
    Rosenbrock, Sampler, Step2d, UpdateFlagSpec, init_logprob, pinned_empty, propose_move, sample, sample_pt)
from .hmc import HmcParam  # noqa: F401
from .twalk import TWalkKernal, TWalkParams  # noqa: F401
from .utils import draw_z, scale_vec, swap_walkers  # noqa: F401

__all__ = [
    "ensemble_sample", "utils", "init_ensemble", "linear_space", "twalk", "TWalkParams", "TWalkKernal", "hmc", "HmcParam", "pso", "Sampler", "DeviceRng", "LogProb", "McmcErr", "UpdateFlagSpec",
    "Gaussian", "BoundedGaussian", "Rosenbrock", "Rastrigin", "Bimodal2d", "Step2d", "CorrGaussian", "Mixture",
    "init_logprob", "sample", "sample_pt", "propose_move", "swap_walkers", "draw_z", "scale_vec", "pinned_empty",
]
