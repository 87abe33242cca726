"""The host side of the PSO hand-off (scorus_b200/pso.py, src/opt/pso.rs:108-331) with a Python callable as `func`:
no GPU involved, the logic of from_ensemble / sample / converged against a literal restatement of the reference loops."""
import numpy as np
import pytest

from scorus_b200.pso import Particle, ParticleSwarmMaximizer, init_swarm_from_ensemble


def _f(p):
    return -float(np.sum((p - 0.3) ** 2))


def _literal_sample(pos, vel, fit, pbest, gbest, u, w, c1, c2):
    """src/opt/pso.rs:214-269 as written (loops over particles and coordinates); u[i, j, 0/1] = cognitive / social draw."""
    n, d = pos.shape
    for i in range(n):
        if gbest is None:
            gbest = [pos[i].copy(), vel[i].copy(), fit[i]]
        elif gbest[2] < fit[i]:
            gbest = [pos[i].copy(), vel[i].copy(), fit[i]]
        if pbest[i] is None:
            pbest[i] = [pos[i].copy(), vel[i].copy(), fit[i]]
        elif fit[i] > pbest[i][2]:
            pbest[i] = [pos[i].copy(), vel[i].copy(), fit[i]]
    for i in range(n):
        for j in range(d):
            part_vel = w * vel[i, j]
            cog_vel = c1 * u[i, j, 0] * (pbest[i][0][j] - pos[i, j])
            soc_vel = c2 * u[i, j, 1] * (gbest[0][j] - pos[i, j])
            vel[i, j] = part_vel + cog_vel + soc_vel
            pos[i, j] = pos[i, j] + vel[i, j]
    return gbest


def test_from_ensemble_and_guess():
    rng = np.random.default_rng(0)
    ens = rng.normal(size=(12, 3))
    m = ParticleSwarmMaximizer.from_ensemble(_f, ens, guess=np.array([0.3, 0.3, 0.3]))
    assert m.particle_count == 12 and m.ndim == 3
    assert np.array_equal(m.position, ens) and not m.velocity.any()
    assert np.allclose(m.fitness, [_f(p) for p in ens])
    assert isinstance(m.gbest, Particle) and m.gbest.fitness == 0.0 and not m.gbest.velocity.any()
    sw = m.swarm
    assert len(sw) == 12 and sw[5].pbest is None and np.array_equal(sw[5].position, ens[5])
    assert ParticleSwarmMaximizer.from_ensemble(_f, ens).gbest is None
    parts = init_swarm_from_ensemble(_f, ens)
    assert [p.fitness for p in parts] == [p.fitness for p in sw]
    with pytest.raises(IndexError):
        ParticleSwarmMaximizer.from_ensemble(_f, np.zeros((0, 3)))


def test_sample_follows_the_reference_loops():
    rng = np.random.default_rng(1)
    ens = rng.normal(size=(9, 4))
    ens[4] = ens[2]                      # a tie in fitness: the earlier particle stays the global best
    m = ParticleSwarmMaximizer.from_ensemble(_f, ens)
    pos, vel = ens.copy(), np.zeros_like(ens)
    fit = np.array([_f(p) for p in pos])
    pbest, gbest = [None] * 9, None
    for it in range(6):
        u = np.random.default_rng(100 + it).random((9, 4, 2))
        gbest = _literal_sample(pos, vel, fit, pbest, gbest, u, 0.7, 1.4, 1.3)
        fit = np.array([_f(p) for p in pos])
        m.sample(np.random.default_rng(100 + it), 0.7, 1.4, 1.3)
        assert np.allclose(m.position, pos, rtol=0, atol=1e-15) and np.allclose(m.velocity, vel, rtol=0, atol=1e-15)
        assert np.allclose(m.fitness, fit) and m.gbest.fitness == gbest[2] and np.array_equal(m.gbest.position, gbest[0])
        assert np.allclose(m.pbest_fitness, [pb[2] for pb in pbest])
    assert m.gbest.fitness >= max(_f(p) for p in ens)       # the best ever seen
    assert m.swarm[0].pbest is not None


def test_convergence_tests():
    ens = np.tile(np.array([[0.3, 0.3]]), (10, 1)) + 1e-4 * np.arange(10)[:, None]
    m = ParticleSwarmMaximizer.from_ensemble(_f, ens)
    assert not m.converged(0.5, 1.0, 1.0)                   # no global best yet (:296-299, :327-330)
    m.sample(np.random.default_rng(3), 0.0, 0.0, 0.0)       # w = c1 = c2 = 0: nothing moves, the bests are recorded
    assert np.array_equal(m.position, ens)
    best = np.sort(m.fitness)[::-1]
    i1 = 5
    want_dfit = abs(m.gbest.fitness - best[1:i1].sum() / (i1 - 1))
    assert m.converge_dfit(0.5, want_dfit * 1.01) and not m.converge_dfit(0.5, want_dfit * 0.99)
    order = np.argsort(-m.fitness, kind="stable")[:i1]
    want_d2 = max(float(np.sum((m.gbest.position - ens[k]) ** 2)) for k in order)
    assert m.converge_dspace(0.5, want_d2 * 1.01) and not m.converge_dspace(0.5, want_d2 * 0.99)
    assert m.converged(0.5, want_dfit * 1.01, want_d2 * 1.01)
