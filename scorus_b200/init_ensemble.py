"""Host-side mirror of scorus::mcmc::init_ensemble (src/mcmc/init_ensemble.rs) over the C ABI.

    get_one_init_realization(y1, y2, rng)   # :19-32  one walker, x[d] ~ U[y1[d], y2[d])

Every call advances rng's init counter, so successive calls give successive walkers of the device
stream; Sampler.init_ensemble(lo, hi) fills a whole ensemble in one launch.
"""
from __future__ import annotations

import numpy as np

from .ensemble_sample import DeviceRng, Gaussian, Sampler


def get_one_init_realization(y1, y2, rng: DeviceRng) -> np.ndarray:
    y1 = np.ascontiguousarray(y1, dtype=np.float64)
    y2 = np.ascontiguousarray(y2, dtype=np.float64)
    k = getattr(rng, "_init_calls", 0)
    rng._init_calls = k + 1
    with Sampler(Gaussian(), 2, y1.size, seed=rng.seed, device=rng.device) as s:
        s.set_shard(k, 0, k + 2)          # walker 0 of this context is global walker k of the stream
        s.init_ensemble(y1, y2)
        ens, _ = s.download()
    return ens[0]
