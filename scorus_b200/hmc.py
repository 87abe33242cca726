"""Host-side mirror of the reference module `mcmc::hmc::naive` (src/mcmc/hmc/naive.rs) for its ensemble /
parallel-tempering entry point: `HmcParam` (:17-56) and `sample_ensemble_pt` (:122-158).  The trajectories run
on the device (scorus_b200/csrc/hmc_kernel.cuh); nothing here computes.

The reference takes the gradient as a second closure, `grad_logprob: &G`.  Here the gradient belongs to the
log-density plugin (Gaussian, BoundedGaussian, Rosenbrock, Rastrigin, Mixture have one); the `grad_logprob` argument is kept
for the call shape and must be None or the same plugin.  The single-chain `sample` (:58-120) is the same arithmetic
for one walker at beta = 1 on the calling thread -- not a data-parallel path, not mirrored.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import _lib as L
from .ensemble_sample import DeviceRng, LogProb, _check_arrays


@dataclass(frozen=True)
class HmcParam:
    """naive.rs:17-23"""
    target_accept_ratio: float
    adj_factor: float

    @staticmethod
    def new(target_accept_ratio: float, adj_factor: float) -> "HmcParam":   # :29-34
        return HmcParam(float(target_accept_ratio), float(adj_factor))

    @staticmethod
    def quick_adj(target_accept_ratio: float) -> "HmcParam":                # :36-38
        return HmcParam(float(target_accept_ratio), 0.01)

    @staticmethod
    def slow_adj(target_accept_ratio: float) -> "HmcParam":                 # :40-42
        return HmcParam(float(target_accept_ratio), 0.000_001)

    @staticmethod
    def fixed(target_accept_ratio: float) -> "HmcParam":                    # :44-46
        return HmcParam(float(target_accept_ratio), 0.0)

    @staticmethod
    def default() -> "HmcParam":                                            # :49-56
        return HmcParam(0.99, 0.000_001)

    def _c(self) -> L.HmcParam:
        c = L.HmcParam()
        c.target_accept_ratio, c.adj_factor = self.target_accept_ratio, self.adj_factor
        return c


def sample_ensemble_pt(flogprob: LogProb, grad_logprob, q0: np.ndarray, lp: np.ndarray, rng: DeviceRng,
                       epsilon: np.ndarray, beta_list, l: int, param: HmcParam):
    """naive.rs:122-158, in place on the host arrays q0 [n][dim], lp [n] and epsilon [nbeta]; returns the
    accepted count per rung (the reference's Vec<usize>)."""
    if grad_logprob is not None and grad_logprob is not flogprob and grad_logprob != flogprob:
        raise ValueError("the gradient is the plugin's own: pass grad_logprob=None (or the same LogProb)")
    _check_arrays(q0, lp)
    if q0.shape[0] != lp.shape[0]:
        raise AssertionError("assert_eq!(q0.len(), lp.len())")          # :177
    s = rng.sampler(flogprob, *q0.shape)
    return [int(c) for c in s.hmc_sample_ensemble_pt_host(q0, lp, epsilon, np.ascontiguousarray(beta_list, dtype=np.float64),
                                                          l, param)]
