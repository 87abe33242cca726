"""The HMC oracle (oracle/hmc.c, a restatement of src/mcmc/hmc/naive.rs:122-292 + utils.rs:6-30) against the
independent numpy restatement, finite-difference gradients and the moments the sampler must reproduce."""
import numpy as np
import pytest

from oracle import numpy_oracle as no
from tests.util import hmc_golden_files


@pytest.mark.parametrize("path", hmc_golden_files(), ids=lambda p: p.split("/")[-1][:-4])
def test_golden_vectors_both_restatements(oracle, path):
    g = np.load(path)
    kind, params, beta, l = int(g["kind"]), g["params"], g["beta"], int(g["l"])
    target, adj = float(g["target"]), float(g["adj"])
    q, lp, eps = g["q0"].copy(), g["lp0"].copy(), g["eps0"].copy()
    for k in range(g["p0"].shape[0]):
        q2, l2, e2, a2, c2 = no.hmc_ensemble_pt_fixed(kind, params, q, lp, eps, beta, l, target, adj, g["p0"][k], g["u1"][k], g["u2"][k])
        acc, cnt = oracle.hmc_ensemble_pt_fixed(kind, params, q, lp, eps, beta, l, target, adj, g["p0"][k], g["u1"][k], g["u2"][k])
        assert np.array_equal(acc, g["accepted"][k]) and np.array_equal(a2, acc) and np.array_equal(cnt, g["cnt"][k])
        assert np.array_equal(q, g["q"][k]) and np.array_equal(q2, q)
        assert np.array_equal(lp, g["lp"][k], equal_nan=True) and np.array_equal(l2, lp, equal_nan=True)
        assert np.array_equal(eps, g["eps"][k]) and np.array_equal(e2, eps)


def _corr9():
    g = np.random.default_rng(77)
    L = np.eye(9) * 1.5 + 0.1 * g.normal(size=(9, 9))
    return np.concatenate([g.uniform(-0.2, 0.2, 9), L.ravel()]).tolist()


@pytest.mark.parametrize("kind,params", [(0, [0.5]), (0, [1.0]), (1, [1.5]), (2, []), (3, [10.0]),
                                         (7, [1.0, 2.0, 1.5] + [0.0] * 8), (6, _corr9())],
                         ids=lambda v: str(v) if isinstance(v, int) else f"p{len(v)}")
def test_gradients_match_finite_differences_and_numpy(oracle, kind, params):
    g = np.random.default_rng(kind)
    x = g.uniform(-1.0, 1.0, (5, 9))
    gr = oracle.grad_logprob(kind, params, x)
    assert np.array_equal(gr, no.grad_logprob(kind, params, x))
    h = 1e-6
    for d in range(x.shape[1]):
        xp, xm = x.copy(), x.copy()
        xp[:, d] += h
        xm[:, d] -= h
        num = (oracle.init_logprob(kind, params, xp) - oracle.init_logprob(kind, params, xm)) / (2 * h)
        assert np.allclose(gr[:, d], num, rtol=1e-5, atol=1e-5)


def test_plugins_without_gradient_say_so(oracle):
    with pytest.raises(ValueError):
        oracle.grad_logprob(4, [], np.zeros((1, 2)) + 0.3)


@pytest.mark.parametrize("kind,params,n,dim,nb,l", [(0, [0.5], 32, 5, 2, 4), (2, [], 24, 8, 3, 3), (3, [10.0], 40, 10, 4, 5),
                                                    (1, [1.5], 17, 3, 1, 6), (6, _corr9(), 20, 9, 2, 3)],
                         ids=lambda v: str(v) if isinstance(v, int) else f"p{len(v)}")
def test_both_restatements_agree(oracle, kind, params, n, dim, nb, l):
    g = np.random.default_rng(dim)
    q = np.ascontiguousarray(g.uniform(0.8, 1.2, (n, dim)) if kind == 2 else g.uniform(-1, 1, (n, dim)))
    lp = oracle.init_logprob(kind, params, q)
    eps = np.full(nb, 0.02 if kind == 2 else 0.1)
    beta = 2.0 ** -np.arange(nb, dtype=np.float64)
    for step in range(5):
        p0, u1, u2 = g.normal(size=(n, dim)), g.uniform(size=n), g.uniform(size=n)
        q2, l2, e2, a2, c2 = no.hmc_ensemble_pt_fixed(kind, params, q, lp, eps, beta, l, 0.8, 0.01, p0, u1, u2)
        a1, c1 = oracle.hmc_ensemble_pt_fixed(kind, params, q, lp, eps, beta, l, 0.8, 0.01, p0, u1, u2)
        assert np.array_equal(a1, a2) and np.array_equal(c1, c2) and np.array_equal(q, q2)
        assert np.array_equal(eps, e2) and np.array_equal(lp, l2, equal_nan=True)


def test_leapfrog_is_reversible_and_conserves_energy(oracle):
    """One trajectory forward, momenta flipped, the same trajectory again: back at the start (to rounding); a small
    step keeps H = -lp + |p|^2/2 to O(eps^2)."""
    n, dim, l, eps = 8, 4, 20, 0.01
    g = np.random.default_rng(1)
    q = np.ascontiguousarray(g.normal(size=(n, dim)))
    lp = oracle.init_logprob(0, [0.5], q)
    p0 = g.normal(size=(n, dim))
    q1, lp1 = q.copy(), lp.copy()
    acc, _ = oracle.hmc_ensemble_pt_fixed(0, [0.5], q1, lp1, np.array([eps]), [1.0], l, 0.5, 0.0, p0, np.zeros(n), np.ones(n))
    assert acc.all()                                    # u1 = 0 accepts any finite dh
    # unit Gaussian: the flow is a rotation of (q, p); after time t = l eps, p1 = p0 cos t - q0 sin t
    t = l * eps
    p1 = p0 * np.cos(t) - q * np.sin(t)
    assert np.allclose(q1, q * np.cos(t) + p0 * np.sin(t), atol=1e-4)
    q2, lp2 = q1.copy(), lp1.copy()
    oracle.hmc_ensemble_pt_fixed(0, [0.5], q2, lp2, np.array([eps]), [1.0], l, 0.5, 0.0, -p1, np.zeros(n), np.ones(n))
    assert np.allclose(q2, q, atol=1e-4)


def test_oracle_hmc_samples_tempered_gaussian_and_adapts(oracle):
    n, dim = 64, 5
    g = np.random.default_rng(0)
    q = np.ascontiguousarray(g.normal(size=(n, dim)))
    lp = oracle.init_logprob(0, [0.5], q)
    eps, beta = np.array([0.3, 0.3]), np.array([1.0, 0.25])
    rng = oracle.Rng(1)
    s2, c, acc = np.zeros(2), 0, np.zeros(2)
    for k in range(3000):
        acc += oracle.hmc_ensemble_pt(0, [0.5], q, lp, rng, eps, beta, 5, 0.8, 0.01)
        if k > 300:
            s2 += [(q[:32] ** 2).mean(), (q[32:] ** 2).mean()]
            c += 1
    assert abs(s2[0] / c - 1.0) < 0.06 and abs(s2[1] / c - 4.0) < 0.25
    assert np.all(np.abs(acc / (3000 * 32) - 0.8) < 0.03) and eps[1] > eps[0]


def test_argument_checks(oracle):
    q, lp = np.zeros((6, 2)), np.zeros(6)
    with pytest.raises(ValueError):
        oracle.hmc_ensemble_pt_fixed(0, [0.5], q, lp, np.ones(4), [1.0, 0.5, 0.25, 0.125], 2, 0.8, 0.01, np.zeros((6, 2)),
                                     np.zeros(6), np.zeros(6))
