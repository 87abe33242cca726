"""Hand-off from an MCMC ensemble to the reference's particle-swarm maximiser:
`ParticleSwarmMaximizer::from_ensemble` / `init_swarm_from_ensemble` (src/opt/pso.rs:108-132, 177-192) turn an
ensemble into particles with zero velocity and fitness = func(position).  When `func` is the sampler's own
log-density the fitness is the cached log-prob the device already holds, so the hand-off is one download.
The swarm iteration itself (src/opt/pso.rs:194-) is outside the hot path and is not built here.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import numpy as np

from .ensemble_sample import DeviceRng, LogProb, Sampler, init_logprob


@dataclass
class Particle:
    """src/opt/pso.rs:16-36"""
    position: np.ndarray
    velocity: np.ndarray
    fitness: float
    pbest: Optional["Particle"] = None


def init_swarm_from_ensemble(func: LogProb, ensemble: np.ndarray, rng: Optional[DeviceRng] = None) -> List[Particle]:
    """src/opt/pso.rs:177-192: position = walker, velocity = walker * 0, fitness = func(walker) (evaluated on the
    device by init_logprob)."""
    ensemble = np.ascontiguousarray(ensemble, dtype=np.float64)
    fitness = np.empty(ensemble.shape[0])
    init_logprob(func, ensemble, fitness, rng)
    return [Particle(ensemble[i].copy(), ensemble[i] * 0.0, float(fitness[i])) for i in range(ensemble.shape[0])]


def swarm_from_sampler(sampler: Sampler) -> List[Particle]:
    """The same hand-off straight from a device-resident sampler: its cached log-probs are the fitness."""
    ens, lp = sampler.download()
    return [Particle(ens[i].copy(), ens[i] * 0.0, float(lp[i])) for i in range(ens.shape[0])]
