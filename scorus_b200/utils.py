"""Host-side mirror of scorus::mcmc::utils (src/mcmc/utils.rs) over the C ABI.

    draw_z(rng, a)                                        # :33
    scale_vec(x1, x2, z)                                  # :47
    swap_walkers(ensemble, logprob, rng, beta_list)       # :72
"""
from __future__ import annotations

import numpy as np

from . import _lib as L
from .ensemble_sample import DeviceRng, LogProb, _check_arrays, _raise, propose_move


def swap_walkers(ensemble: np.ndarray, logprob: np.ndarray, rng: DeviceRng, beta_list, flogprob: LogProb = None):
    """src/mcmc/utils.rs:72-112, in place on the host arrays.  A length mismatch between ensemble
    and logprob is the reference's silent no-op (:88)."""
    _check_arrays(ensemble, logprob)
    n, dim = ensemble.shape
    nb = len(beta_list)
    if nb == 0 or (n // nb) * nb != n:
        _raise(L.ERR_NWALKERS_MISMATCHES_NBETA if nb else L.ERR_INVALID_ARG)
    if logprob.shape[0] != n:
        return
    s = rng.sampler(flogprob or LogProb(L.LP_GAUSSIAN, (0.5,)), n, dim)
    s.swap_walkers_host(ensemble, logprob, beta_list)


def scale_vec(x1, x2, z: float) -> np.ndarray:
    """src/mcmc/utils.rs:47-56: (x1*z) + (x2*(1-z)), evaluated by the step kernel."""
    return propose_move(x1, x2, z, None)


def draw_z(rng: DeviceRng, a: float) -> float:
    """src/mcmc/utils.rs:33-45: one stretch factor from the device stream; each call advances
    the stream's step counter."""
    s = rng.sampler(LogProb(L.LP_GAUSSIAN, (0.5,)), 2, 2)
    return float(s.draw_z(a, 1)[0])
