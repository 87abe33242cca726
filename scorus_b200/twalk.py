"""Host-side mirror of the reference module `mcmc::twalk` (src/mcmc/twalk.rs) for its ensemble /
parallel-tempering entry point: `TWalkParams` (:75-130), `TWalkKernal` (:32-73) and `sample` (:586-648).
The moves run on the device (scorus_b200/csrc/twalk_kernel.cuh); nothing here computes.

The single-state `sample_st` (:534-584, one (x, x') pair on the calling thread) is not a data-parallel path and
is not mirrored; an ensemble of two walkers through `sample` covers the same move set.
"""
from __future__ import annotations

import enum
from dataclasses import dataclass, replace

import numpy as np

from . import _lib as L
from .ensemble_sample import DeviceRng, LogProb, _check_arrays, _raise


class TWalkKernal(enum.IntEnum):
    """twalk.rs:32-37 (spelling as in the reference); `random` (:40-63) only ever returns the first two."""
    Walk = 0
    Traverse = 1
    Blow = 2
    Hop = 3

    def to_usize(self) -> int:
        return int(self)


@dataclass(frozen=True)
class TWalkParams:
    """twalk.rs:75-83.  `fw` holds CUMULATIVE weights, as the reference stores them."""
    aw: float
    at: float
    pphi: float
    fw: tuple

    @staticmethod
    def new(n: int) -> "TWalkParams":
        """TWalkParams::new(n), :89-110: aw = 1.5, at = 6, pphi = min(n, 4) / n, fw = cumsum [0.5, 0.5]."""
        c = L.TwalkParams()
        _raise(L.load().scorus_twalk_params_default(n, c))
        return TWalkParams(c.aw, c.at, c.pphi, (c.fw[0], c.fw[1]))

    def with_fw(self, fw) -> "TWalkParams":
        """:112-124: plain weights in, cumulative sums stored."""
        return replace(self, fw=(0.0 + float(fw[0]), 0.0 + float(fw[0]) + float(fw[1])))

    def with_pphi(self, pphi: float) -> "TWalkParams":
        """:126-128"""
        return replace(self, pphi=float(pphi))

    def _c(self) -> L.TwalkParams:
        c = L.TwalkParams()
        c.aw, c.at, c.pphi = self.aw, self.at, self.pphi
        c.fw[0], c.fw[1] = self.fw
        return c


def sample(flogprob: LogProb, ensemble_logprob, param: TWalkParams, rng: DeviceRng, beta_list, nthreads: int = 1):
    """twalk::sample, :586-648, in place on the host pair (ensemble [n][dim], logprob [n]).  `nthreads` (the
    reference's worker count for the log-density evaluations) is accepted and ignored."""
    ensemble, logprob = ensemble_logprob
    _check_arrays(ensemble, logprob)
    nb = len(beta_list)
    _raise(L.load().scorus_check_sample_args(ensemble.shape[0], logprob.shape[0], nb))
    s = rng.sampler(flogprob, *ensemble.shape)
    s.twalk_sample_host(ensemble, logprob, param, np.ascontiguousarray(beta_list, dtype=np.float64))
