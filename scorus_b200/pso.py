"""Hand-off from an MCMC ensemble to the reference's particle-swarm maximiser.

`ParticleSwarmMaximizer::from_ensemble` / `init_swarm_from_ensemble` (src/opt/pso.rs:108-132, 177-192) turn an
ensemble into particles with zero velocity and fitness = func(position); the maximiser then iterates `sample`
(:214-272) and tests `converged` (:274-331).  What is data-parallel in that loop is `update_fitness` (:194-212, rayon
`par_iter` in the reference): here it is one `init_logprob` launch over all particles when `func` is a device
log-density plugin -- and when the swarm comes straight from a device-resident sampler (`from_sampler`) the first
fitness is the cached log-prob the device already holds, so the hand-off is one download.  The velocity / position
update is a few vector operations per iteration on the host (the reference runs it on the calling thread); the swarm is
kept as arrays, `swarm` / `gbest` give the reference's `Particle` view.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Union

import numpy as np

from .ensemble_sample import DeviceRng, LogProb, Sampler, init_logprob

Func = Union[LogProb, Callable[[np.ndarray], float]]


@dataclass
class Particle:
    """src/opt/pso.rs:16-36"""
    position: np.ndarray
    velocity: np.ndarray
    fitness: float
    pbest: Optional["Particle"] = None


def _evaluate(func: Func, positions: np.ndarray, rng: Optional[DeviceRng] = None) -> np.ndarray:
    """func over every row: a device plugin goes through init_logprob (one launch), a Python callable is called per row."""
    positions = np.ascontiguousarray(positions, dtype=np.float64)
    if isinstance(func, LogProb):
        fitness = np.empty(positions.shape[0])
        init_logprob(func, positions, fitness, rng)
        return fitness
    return np.array([float(func(p)) for p in positions], dtype=np.float64)


def init_swarm_from_ensemble(func: Func, ensemble: np.ndarray, rng: Optional[DeviceRng] = None) -> List[Particle]:
    """src/opt/pso.rs:177-192: position = walker, velocity = walker * 0, fitness = func(walker)."""
    ensemble = np.ascontiguousarray(ensemble, dtype=np.float64)
    fitness = _evaluate(func, ensemble, rng)
    return [Particle(ensemble[i].copy(), ensemble[i] * 0.0, float(fitness[i])) for i in range(ensemble.shape[0])]


def swarm_from_sampler(sampler: Sampler) -> List[Particle]:
    """The same hand-off straight from a device-resident sampler: its cached log-probs are the fitness."""
    ens, lp = sampler.download()
    return [Particle(ens[i].copy(), ens[i] * 0.0, float(lp[i])) for i in range(ens.shape[0])]


class ParticleSwarmMaximizer:
    """src/opt/pso.rs:38-331 with the swarm as arrays: `position`, `velocity` [particle_count, ndim], `fitness`
    [particle_count], the personal bests (`pbest_position`, `pbest_velocity`, `pbest_fitness`, None before the first
    `sample`) and the global best (`gbest`, a Particle or None)."""

    def __init__(self, func: Func, position: np.ndarray, fitness: np.ndarray, gbest: Optional[Particle],
                 device_rng: Optional[DeviceRng] = None):
        self.func = func
        self.position = np.ascontiguousarray(position, dtype=np.float64)
        self.velocity = np.zeros_like(self.position)                      # &ensemble[i] * T::zero(), :182
        self.fitness = np.ascontiguousarray(fitness, dtype=np.float64)
        self.particle_count, self.ndim = self.position.shape
        self.gbest = gbest
        self.pbest_position = self.pbest_velocity = self.pbest_fitness = None
        self._device_rng = device_rng

    # -- constructors
    @classmethod
    def from_ensemble(cls, func: Func, ensemble: np.ndarray, guess: Optional[np.ndarray] = None,
                      device_rng: Optional[DeviceRng] = None) -> "ParticleSwarmMaximizer":
        """src/opt/pso.rs:108-132"""
        ensemble = np.ascontiguousarray(ensemble, dtype=np.float64)
        if ensemble.ndim != 2 or ensemble.shape[0] == 0:
            raise IndexError("ensemble[0] of an empty ensemble (index panic, :115)")
        fitness = _evaluate(func, ensemble, device_rng)
        return cls(func, ensemble.copy(), fitness, cls._guess(func, guess, ensemble.shape[1], device_rng), device_rng)

    @classmethod
    def from_sampler(cls, sampler: Sampler, func: Optional[Func] = None, guess: Optional[np.ndarray] = None,
                     device_rng: Optional[DeviceRng] = None) -> "ParticleSwarmMaximizer":
        """from_ensemble on the ensemble a sampler holds on the device, with `func` the sampler's own log-density: the
        cached log-probs ARE the fitness, nothing is evaluated again."""
        ens, lp = sampler.download()
        func = sampler.flogprob if func is None else func
        return cls(func, ens, lp, cls._guess(func, guess, ens.shape[1], device_rng), device_rng)

    @staticmethod
    def _guess(func: Func, guess, ndim: int, device_rng) -> Optional[Particle]:
        if guess is None:
            return None
        g = np.ascontiguousarray(guess, dtype=np.float64).reshape(ndim)
        if isinstance(func, LogProb):                                   # (the device evaluates ensembles: two copies of the guess)
            f = float(_evaluate(func, np.stack([g, g]), device_rng)[0])
        else:
            f = float(func(g))
        return Particle(g.copy(), g * 0.0, f)

    # -- the reference's views
    @property
    def swarm(self) -> List[Particle]:
        out = []
        for i in range(self.particle_count):
            pb = None
            if self.pbest_fitness is not None:
                pb = Particle(self.pbest_position[i].copy(), self.pbest_velocity[i].copy(), float(self.pbest_fitness[i]))
            out.append(Particle(self.position[i].copy(), self.velocity[i].copy(), float(self.fitness[i]), pb))
        return out

    # -- iteration
    def update_fitness(self):
        """src/opt/pso.rs:194-212"""
        self.fitness = _evaluate(self.func, self.position, self._device_rng)

    def sample(self, rng: np.random.Generator, w: float, c1: float, c2: float):
        """src/opt/pso.rs:214-272.  The uniforms are drawn in the reference's order (per particle, per coordinate: the
        cognitive one, then the social one) from a numpy Generator."""
        # global best: the first particle when there is none, then every strictly better one in order (:216-232)
        best = int(np.argmax(self.fitness))          # argmax returns the FIRST maximum: ties keep the earlier particle
        if self.gbest is None or self.gbest.fitness < self.fitness[best]:
            self.gbest = Particle(self.position[best].copy(), self.velocity[best].copy(), float(self.fitness[best]))
        # personal bests (:234-252)
        if self.pbest_fitness is None:
            self.pbest_position, self.pbest_velocity = self.position.copy(), self.velocity.copy()
            self.pbest_fitness = self.fitness.copy()
        else:
            better = self.fitness > self.pbest_fitness
            self.pbest_position[better] = self.position[better]
            self.pbest_velocity[better] = self.velocity[better]
            self.pbest_fitness[better] = self.fitness[better]
        # velocities and positions (:255-269)
        u = rng.random((self.particle_count, self.ndim, 2))
        part_vel = w * self.velocity
        cog_vel = c1 * u[:, :, 0] * (self.pbest_position - self.position)
        soc_vel = c2 * u[:, :, 1] * (self.gbest.position[None, :] - self.position)
        self.velocity = part_vel + cog_vel + soc_vel
        self.position = self.position + self.velocity
        self.update_fitness()

    # -- convergence tests
    def converged(self, p: float, m1: float, m2: float) -> bool:
        """src/opt/pso.rs:274-276"""
        return self.converge_dfit(p, m1) and self.converge_dspace(p, m2)

    def converge_dfit(self, p: float, m: float) -> bool:
        """src/opt/pso.rs:278-300: |gbest.fitness - mean of the best fitnesses [1, floor(pc p))| < m"""
        if self.gbest is None:
            return False
        best_sort = np.sort(self.fitness)[::-1]
        i1 = int(np.floor(self.particle_count * p))
        best_mean = best_sort[1:i1].sum() / (i1 - 1) if i1 != 1 else float("nan")
        return bool(abs(self.gbest.fitness - best_mean) < m)

    def converge_dspace(self, p: float, m: float) -> bool:
        """src/opt/pso.rs:302-331: the largest squared distance from gbest among the best floor(pc p) particles < m"""
        if self.gbest is None:
            return False
        order = np.argsort(-self.fitness, kind="stable")
        i1 = int(np.floor(self.particle_count * p))
        d2 = ((self.gbest.position[None, :] - self.position[order[:i1]]) ** 2).sum(axis=1)
        return bool((d2.max() if d2.size else 0.0) < m)
