"""Size- and seed-independent properties of the oracle restatements (hypothesis): what must hold for ANY input, on
top of the golden vectors.  These are the invariants the GPU suites rely on when they compare against the oracle."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import pyoracle as po

po.build()

small = settings(max_examples=40, deadline=None)


@small
@given(seed=st.integers(0, 2 ** 31), nb=st.integers(1, 5), half_npb=st.integers(1, 6), dim=st.integers(1, 9),
       prob=st.sampled_from([0.2, 0.5, 1.0]))
def test_stretch_streams_and_step_invariants(seed, nb, half_npb, dim, prob):
    """partner is a fixed-point-free involution inside each rung; z lies in [1/a, a); a walker either keeps its row
    and log-prob or takes the proposal, and an accepted row's log-prob is the log-density of the new row."""
    n = nb * 2 * half_npb
    rng = po.Rng(seed)
    partner, z, flags, u, _ = po.draw_sample_streams(rng, n, dim, nb, 2.0, po.UFS_PROB, prob)
    npb = n // nb
    idx = np.arange(n)
    assert np.array_equal(partner[partner], idx) and np.all(partner != idx) and np.array_equal(partner // npb, idx // npb)
    assert np.all(z >= 0.5) and np.all(z < 2.0) and np.all(flags.any(axis=1)) and np.all((u >= 0) & (u < 1))
    g = np.random.default_rng(seed)
    ens = np.ascontiguousarray(g.normal(size=(n, dim)))
    lp = po.init_logprob(po.LP_GAUSSIAN, [0.5], ens)
    e0, l0 = ens.copy(), lp.copy()
    beta = 2.0 ** -np.arange(nb, dtype=np.float64)
    rc, acc = po.sample_pt_fixed(po.LP_GAUSSIAN, [0.5], ens, lp, beta, partner, z, flags, u)
    assert rc == 0
    keep = acc == 0
    assert np.array_equal(ens[keep], e0[keep]) and np.array_equal(lp[keep], l0[keep])
    assert np.array_equal(lp[~keep], po.init_logprob(po.LP_GAUSSIAN, [0.5], ens)[~keep])
    # unflagged coordinates never move
    moved = ens != e0
    assert not np.any(moved & (flags == 0))


@small
@given(seed=st.integers(0, 2 ** 31), nb=st.integers(2, 6), npb=st.integers(1, 7), dim=st.integers(1, 5))
def test_swap_is_a_permutation_of_row_logprob_pairs(seed, nb, npb, dim):
    n = nb * npb
    g = np.random.default_rng(seed)
    ens = np.ascontiguousarray(g.normal(size=(n, dim)))
    lp = po.init_logprob(po.LP_GAUSSIAN, [0.5], ens)
    before = sorted(map(tuple, np.c_[ens, lp].tolist()))
    jvec, r = po.draw_swap_streams(po.Rng(seed), n, nb)
    assert all(sorted(row.tolist()) == list(range(npb)) for row in jvec)
    rc, sw = po.swap_walkers_fixed(ens, lp, 2.0 ** -np.arange(nb, dtype=np.float64), jvec, r)
    assert rc == 0 and sorted(map(tuple, np.c_[ens, lp].tolist())) == before
    assert np.array_equal(lp, po.init_logprob(po.LP_GAUSSIAN, [0.5], ens))


@small
@given(seed=st.integers(0, 2 ** 31), nb=st.integers(1, 4), half_npb=st.integers(1, 5), dim=st.integers(1, 8), half=st.integers(0, 1))
def test_twalk_half_pass_touches_only_its_movers(seed, nb, half_npb, dim, half):
    n = nb * 2 * half_npb
    tw = po.twalk_params(dim)
    order, kernel, b, flags, wu, u = po.philox_twalk_streams(seed, 2 + half, 2, n, dim, nb, tw, half)
    npb = n // nb
    assert sorted(order.tolist()) == list(range(n)) and np.array_equal(order // npb, np.arange(n) // npb)
    g = np.random.default_rng(seed)
    ens = np.ascontiguousarray(g.normal(size=(n, dim)))
    lp = po.init_logprob(po.LP_ROSENBROCK, [], ens)
    e0, l0 = ens.copy(), lp.copy()
    acc = po.twalk_half_fixed(po.LP_ROSENBROCK, [], ens, lp, 2.0 ** -np.arange(nb, dtype=np.float64), tw, order, half, kernel, b,
                              flags, wu, u)
    movers = order[half::2]
    still = np.setdiff1d(np.arange(n), movers[acc == 1])
    assert np.array_equal(ens[still], e0[still]) and np.array_equal(lp[still], l0[still])
    assert np.array_equal(lp[movers[acc == 1]], po.init_logprob(po.LP_ROSENBROCK, [], ens)[movers[acc == 1]])
    # a move never changes an unflagged coordinate (Walk adds (x - xp) * 0, Traverse copies x)
    changed = ens[movers] != e0[movers]
    assert not np.any(changed & (flags == 0))


@small
@given(seed=st.integers(0, 2 ** 31), nb=st.integers(1, 4), npb=st.integers(1, 6), dim=st.integers(1, 6), l=st.integers(0, 4))
def test_hmc_rejecting_everything_changes_nothing_but_epsilon(seed, nb, npb, dim, l):
    n = nb * npb
    g = np.random.default_rng(seed)
    q = np.ascontiguousarray(g.normal(size=(n, dim)))
    lp = po.init_logprob(po.LP_GAUSSIAN, [0.5], q)
    q0, l0 = q.copy(), lp.copy()
    eps = np.full(nb, 0.1)
    beta = 2.0 ** -np.arange(nb, dtype=np.float64)
    # an accept draw no exp(dh) can beat rejects every trajectory; u2 = 0 < target shrinks epsilon once per walker
    acc, cnt = po.hmc_ensemble_pt_fixed(po.LP_GAUSSIAN, [0.5], q, lp, eps, beta, l, 0.7, 0.01, g.normal(size=(n, dim)),
                                        np.full(n, 1e300), np.zeros(n))
    assert not acc.any() and not cnt.any() and np.array_equal(q, q0) and np.array_equal(lp, l0)
    want = 0.1
    for _ in range(npb):
        want = want / 1.01
    assert np.all(eps == want)
    # u1 = 0 accepts every finite dh; with l = 0 the "trajectory" is the start point: positions unchanged, all accepted
    acc, cnt = po.hmc_ensemble_pt_fixed(po.LP_GAUSSIAN, [0.5], q, lp, eps, beta, 0, 0.7, 0.0, g.normal(size=(n, dim)),
                                        np.zeros(n), np.ones(n))
    assert acc.all() and np.all(cnt == npb) and np.array_equal(q, q0) and np.allclose(lp, l0, rtol=1e-15)
