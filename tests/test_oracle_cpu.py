"""CPU tests of the oracle itself (no GPU): the C restatement against the independent numpy
restatement, against the committed golden vectors, and against the reference's edge semantics
(SURVEY App. A.4).  The reference has no tests for this path; these pin what parity means."""
import math

import numpy as np
import pytest

from oracle import numpy_oracle as no
from tests.util import golden_files


def test_philox4x32_10_known_answers(oracle):
    # Random123 kat_vectors for philox4x32-10
    assert oracle.philox4x32_10([0, 0, 0, 0], [0, 0]) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert oracle.philox4x32_10([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert oracle.philox4x32_10([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]) == [
        0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


CASES = [(0, [0.5], 2, 1), (0, [1.0], 10, 4), (1, [1.5], 3, 1), (2, [], 6, 2), (2, [], 1, 1), (3, [10.0], 5, 4),
         (4, [], 2, 2), (5, [], 2, 1), (6, None, 5, 1), (7, [1.0, 2.0, 5, 0, 0, 0], 4, 1)]


def _setup(kind, params, d, nb, seed):
    g = np.random.default_rng(seed)
    n = 16 * nb
    if kind in (4, 5):
        ens = np.ascontiguousarray(np.c_[g.uniform(-6, 6, n), g.uniform(-0.1, 1.1, n)])
    else:
        ens = np.ascontiguousarray(g.uniform(-1.6, 1.6, (n, d)))
    if kind == 6:
        L = np.triu(g.normal(size=(d, d))) + np.eye(d)
        params = np.r_[g.normal(size=d), L.ravel()]
    return n, ens, np.asarray(params, dtype=np.float64), 2.0 ** -np.arange(nb, dtype=np.float64)


@pytest.mark.parametrize("kind,params,d,nb", CASES, ids=[f"k{c[0]}_d{c[2]}_b{c[3]}" for c in CASES])
def test_c_oracle_equals_numpy_oracle(oracle, kind, params, d, nb):
    n, ens, params, beta = _setup(kind, params, d, nb, kind)
    lp = oracle.init_logprob(kind, params, ens)
    assert np.array_equal(lp, no.logprob(kind, params, ens), equal_nan=True)
    rng = oracle.Rng(42 + kind)
    for step in range(12):
        for ufs, prob in [(oracle.UFS_PROB, 0.5), (oracle.UFS_ALL, 0.0), (oracle.UFS_ONE_HOT_CYCLE, 0.0)]:
            partner, z, flags, u, _ = oracle.draw_sample_streams(rng, n, d, nb, 2.0, ufs, prob, step)
            e2, l2, a2 = no.sample_pt_fixed(kind, params, ens, lp, beta, partner, z, flags, u)
            rc, a1 = oracle.sample_pt_fixed(kind, params, ens, lp, beta, partner, z, flags, u)
            assert rc == 0
            assert np.array_equal(a1, a2) and np.array_equal(ens, e2) and np.array_equal(lp, l2, equal_nan=True)
        if nb > 1:
            jvec, r = oracle.draw_swap_streams(rng, n, nb)
            e2, l2, s2 = no.swap_walkers_fixed(ens, lp, beta, jvec, r)
            rc, s1 = oracle.swap_walkers_fixed(ens, lp, beta, jvec, r)
            assert rc == 0 and np.array_equal(s1, s2) and np.array_equal(ens, e2)
            assert np.array_equal(lp, l2, equal_nan=True)


@pytest.mark.parametrize("path", golden_files(), ids=lambda p: p.split("/")[-1][:-4])
def test_golden_vectors_reproduced_by_both_oracles(oracle, path):
    g = np.load(path)
    kind, params, beta = int(g["kind"]), g["params"], g["beta"]
    ens, lp = np.ascontiguousarray(g["ens0"]), np.ascontiguousarray(g["lp0"])
    nb, nswap = beta.size, 0
    for k in range(g["z"].shape[0]):
        if nb > 1 and k % 10 == 0:
            e2, l2, s2 = no.swap_walkers_fixed(ens, lp, beta, g["jvec"][nswap], g["r"][nswap])
            rc, sw = oracle.swap_walkers_fixed(ens, lp, beta, g["jvec"][nswap], g["r"][nswap])
            assert rc == 0 and np.array_equal(sw, g["swapped"][nswap]) and np.array_equal(s2, sw)
            assert np.array_equal(ens, e2)
            nswap += 1
        e2, l2, a2 = no.sample_pt_fixed(kind, params, ens, lp, beta, g["partner"][k], g["z"][k], g["flags"][k], g["u"][k])
        rc, acc = oracle.sample_pt_fixed(kind, params, ens, lp, beta, g["partner"][k], g["z"][k], g["flags"][k], g["u"][k])
        assert rc == 0
        assert np.array_equal(acc, g["accepted"][k]) and np.array_equal(a2, acc)
        assert np.array_equal(ens, g["ens"][k]) and np.array_equal(e2, ens)
        assert np.array_equal(lp, g["lp"][k], equal_nan=True) and np.array_equal(l2, lp, equal_nan=True)


def test_draw_z_formula_and_range(oracle):
    # src/mcmc/utils.rs:39-44
    a = 2.0
    sa = math.sqrt(a)
    assert oracle.z_range(a) == 2.0 * (sa - 1.0 / sa)
    for p in [0.0, 0.3, oracle.z_range(a) * (1 - 2 ** -53)]:
        y = 1.0 / sa + p / 2.0
        assert oracle.z_from_p(p, a) == y * y
    assert abs(oracle.z_from_p(0.0, a) - 1.0 / a) < 1e-15
    assert oracle.z_from_p(oracle.z_range(a) * (1 - 2 ** -53), a) < a


def test_propose_move_rounding_and_short_flags(oracle):
    p1, p2 = np.array([0.1, 0.2, 0.3, 0.7]), np.array([1.5, -2.5, 3.25, 1e-3])
    z = 1.2345678901234567
    want = p1 * z + p2 * (1.0 - z)  # two rounded products, one rounded sum
    assert np.array_equal(oracle.propose_move(p1, p2, z, None), want)
    out = oracle.propose_move(p1, p2, z, np.array([0, 1], dtype=np.uint8))  # fewer flags than dims (:68)
    assert out[0] == p1[0] and out[1] == want[1] and np.array_equal(out[2:], want[2:])


def test_accept_edge_semantics(oracle):
    """-inf / NaN handling in the accept test (ensemble_sample.rs:183-188)."""
    kind, params = 1, [1.5]  # bounded Gaussian: -inf outside |x| <= 1.5
    ens = np.array([[0.1], [1.45], [1.6], [1.7], [0.2], [1.4]])  # walkers 2,3 start at -inf
    lp = oracle.init_logprob(kind, params, ens)
    assert np.isneginf(lp[2]) and np.isneginf(lp[3])
    partner = np.array([1, 0, 3, 2, 5, 4], dtype=np.uint32)
    flags = np.ones((6, 1), dtype=np.uint8)
    # z = 2 pushes walker 1 away from walker 0: 1.45*2 - 0.1 = 2.8 -> -inf -> q = 0 -> reject even with u = 0
    # walkers 2,3: old = new = -inf -> NaN -> reject.  walker 4: finite -> accept with u = 0
    z = np.array([1.0, 2.0, 1.0, 1.0, 0.9, 1.0])
    u = np.zeros(6)
    e, l = ens.copy(), lp.copy()
    rc, acc = oracle.sample_pt_fixed(kind, params, e, l, [1.0], partner, z, flags, u)
    assert rc == 0
    assert acc.tolist() == [1, 0, 0, 0, 1, 1]
    # old = -inf, new finite -> q = +inf -> accept even with u just below 1
    ens2 = np.array([[1.6], [0.0]])
    lp2 = oracle.init_logprob(kind, params, ens2)
    rc, acc = oracle.sample_pt_fixed(kind, params, ens2, lp2, [1.0], [1, 0], [0.5, 1.0], np.ones((2, 1), np.uint8),
                                     [1 - 2 ** -53, 0.5])
    assert rc == 0 and acc[0] == 1 and ens2[0, 0] == 0.8
    # beta = 0 with an infinite delta -> NaN -> reject
    ens3 = np.array([[1.6], [0.0]])
    lp3 = oracle.init_logprob(kind, params, ens3)
    rc, acc = oracle.sample_pt_fixed(kind, params, ens3, lp3, [0.0], [1, 0], [0.5, 1.0], np.ones((2, 1), np.uint8),
                                     [0.0, 0.0])
    assert rc == 0 and acc[0] == 0


def test_swap_edge_semantics(oracle):
    # NaN probability (both -inf) -> no swap; hotter walker with higher lp always moves down
    ens = np.array([[0.0], [1.0], [2.0], [3.0]])
    lp = np.array([-np.inf, -5.0, -np.inf, -1.0])
    beta = [1.0, 0.5]
    jvec = np.array([[0, 1]], dtype=np.uint32)
    rc, sw = oracle.swap_walkers_fixed(ens, lp, beta, jvec, np.array([[0.0, 0.999]]))
    assert rc == 0 and sw.tolist() == [[0, 1]]
    assert ens[:, 0].tolist() == [0.0, 3.0, 2.0, 1.0] and lp.tolist() == [-np.inf, -1.0, -np.inf, -5.0]
    assert oracle.exchange_prob(-1.0, -5.0, 0.5, 1.0) == 1.0
    assert oracle.exchange_prob(-5.0, -1.0, 0.5, 1.0) == math.exp(0.5 * -4.0)
    assert math.isnan(oracle.exchange_prob(-np.inf, -np.inf, 0.5, 1.0))


def test_error_codes_follow_assert_order(oracle):
    d = 2
    streams = lambda n: (np.arange(n, dtype=np.uint32) ^ 1, np.ones(n), np.ones((n, d), np.uint8), np.zeros(n))

    def call(n, nbeta, lp_len=None):
        ens, lp = np.zeros((n, d)), np.zeros(n if lp_len is None else lp_len)
        p, z, f, u = streams(max(n, 2))
        return oracle.sample_pt_fixed(0, [0.5], ens, lp, np.ones(nbeta), p[:n], z[:n], f[:n], u[:n])[0]

    assert call(12, 5) == oracle.ERR_NWALKERS_MISMATCHES_NBETA   # ensemble_sample.rs:130
    assert call(0, 1) == oracle.ERR_NWALKERS_IS_ZERO              # :131
    assert call(12, 4) == oracle.ERR_NWALKERS_IS_NOT_EVEN         # :132
    assert call(12, 2, lp_len=11) == oracle.ERR_LENGTH_MISMATCH   # :133
    ens, lp = np.zeros((4, d)), np.zeros(4)
    assert oracle.sample_pt_fixed(0, [0.5], ens, lp, [1.0], [1, 0, 3, 2], np.ones(4), np.zeros((4, d), np.uint8),
                                  np.zeros(4))[0] == oracle.ERR_NPHI_ZERO
    assert oracle.sample_pt_fixed(0, [0.5], ens, lp, [1.0, 0.5], [2, 3, 0, 1], np.ones(4), np.ones((4, d), np.uint8),
                                  np.zeros(4))[0] == oracle.ERR_INVALID_ARG  # partner leaves the rung
    # swap: decreasing order enforced, equal tolerated, length mismatch is a silent no-op (utils.rs:88,93)
    jvec, r = np.zeros((1, 2), np.uint32), np.zeros((1, 2))
    jvec[0] = [0, 1]
    assert oracle.swap_walkers_fixed(ens, lp, [0.5, 1.0], jvec, r)[0] == oracle.ERR_BETA_NOT_IN_DECR_ORD
    assert oracle.swap_walkers_fixed(ens, lp, [1.0, 1.0], jvec, r)[0] == 0
    before = ens.copy()
    assert oracle.swap_walkers_fixed(ens, np.zeros(3), [1.0, 0.5], jvec, r)[0] == 0
    assert np.array_equal(ens, before)


def test_reference_order_stream_drawing(oracle):
    n, d, nb = 48, 5, 4
    rng = oracle.Rng(7)
    partner, z, flags, u, oh = oracle.draw_sample_streams(rng, n, d, nb, 2.0, oracle.UFS_PROB, 0.3)
    npb = n // nb
    assert np.array_equal(partner[partner], np.arange(n)) and np.all(partner != np.arange(n))
    assert np.array_equal(partner // npb, np.arange(n) // npb)      # never leaves the rung
    assert np.all(z >= 0.5) and np.all(z < 2.0) and np.all((u >= 0) & (u < 1))
    assert np.all(flags.sum(axis=1) >= 1)                            # redrawn until one is set (:47-49)
    # the cycling closure of examples/sample_high_dim.rs:59-65: one call per walker, continuing across steps
    _, _, fl, _, oh = oracle.draw_sample_streams(rng, n, d, nb, 2.0, oracle.UFS_ONE_HOT_CYCLE, 0.0, 0)
    assert [int(np.argmax(f)) for f in fl[:7]] == [1, 2, 3, 4, 0, 1, 2] and oh == n % d
    jvec, r = oracle.draw_swap_streams(rng, n, nb)
    assert jvec.shape == (nb - 1, npb) and all(sorted(row) == list(range(npb)) for row in jvec.tolist())


def test_ref_structure_port_equals_oracle(oracle):
    """The timed CPU baseline (per-walker heap vectors, threaded log-density) computes exactly what
    orc_sample_pt computes for the same seed."""
    n, d, nb = 256, 8, 2
    g = np.random.default_rng(3)
    ens = np.ascontiguousarray(g.uniform(0.9, 1.1, (n, d)))
    lp = oracle.init_logprob(2, [], ens)
    beta = np.array([1.0, 0.5])
    e1, l1 = ens.copy(), lp.copy()
    t, stage = oracle.ref_structure_run(2, [], e1, l1, beta, 2.0, oracle.UFS_PROB, 0.5, 99, 5, 4)
    assert t > 0 and stage.sum() > 0
    e2, l2 = ens.copy(), lp.copy()
    rng, oh = oracle.Rng(99), 0
    for _ in range(5):
        rc, oh = oracle.sample_pt(2, [], e2, l2, rng, 2.0, oracle.UFS_PROB, 0.5, beta, oh)
        assert rc == 0
    assert np.array_equal(e1, e2) and np.array_equal(l1, l2)
    assert not np.array_equal(e1, ens)


def test_flat_threaded_port_equals_philox_replay(oracle):
    n, d = 512, 6
    g = np.random.default_rng(4)
    ens = np.ascontiguousarray(g.uniform(-1, 1, (n, d)))
    lp = oracle.init_logprob(0, [0.5], ens)
    e1, l1 = ens.copy(), lp.copy()
    assert oracle.flat_threaded_run(0, [0.5], e1, l1, [1.0], 2.0, oracle.UFS_PROB, 0.5, 5, 3, 4, 4) > 0
    e2, l2 = ens.copy(), lp.copy()
    for s in range(3, 7):
        partner, z, flags, u = oracle.philox_sample_streams(5, s, n, d, 1, 2.0, oracle.UFS_PROB, 0.5, 0)
        assert oracle.sample_pt_fixed(0, [0.5], e2, l2, [1.0], partner, z, flags, u)[0] == 0
    assert np.array_equal(e1, e2) and np.array_equal(l1, l2)


@pytest.mark.parametrize("n,nb", [(4, 1), (6, 1), (64, 4), (2048, 1), (4096, 2), (3000, 1)])
def test_philox_pairing_is_a_within_rung_matching(oracle, n, nb):
    for step in range(3):
        partner, z, flags, u = oracle.philox_sample_streams(11, step, n, 3, nb, 2.0, oracle.UFS_PROB, 0.2, 0)
        npb = n // nb
        assert np.array_equal(partner[partner], np.arange(n)) and np.all(partner != np.arange(n))
        assert np.array_equal(partner // npb, np.arange(n) // npb)
        assert np.all(z >= 0.5) and np.all(z < 2.0) and np.all(flags.sum(axis=1) >= 1)
        jvec, r = oracle.philox_swap_streams(11, step, n, nb)
        assert all(sorted(row) == list(range(npb)) for row in jvec.tolist())


def test_philox_small_rung_matchings_are_uniform(oracle):
    """Rungs of <= 1024 walkers get an exact shuffle (sorted Philox keys): with 4 walkers the three
    perfect matchings must be equally likely (the reference's examples run 4 walkers per rung)."""
    counts = {}
    T = 6000
    for step in range(T):
        partner, *_ = oracle.philox_sample_streams(3, step, 4, 2, 1, 2.0, oracle.UFS_ALL, 0.0, 0)
        counts[int(partner[0])] = counts.get(int(partner[0]), 0) + 1
    exp = T / 3
    chi2 = sum((c - exp) ** 2 / exp for c in counts.values())
    assert set(counts) == {1, 2, 3} and chi2 < 13.8  # 2 dof, p = 0.001


def test_philox_large_rung_partner_is_uniform(oracle):
    n = 2048  # > 1024: keyed bijection
    hist = np.zeros(16)
    for step in range(300):
        partner, *_ = oracle.philox_sample_streams(9, step, n, 2, 1, 2.0, oracle.UFS_ALL, 0.0, 0)
        hist += np.histogram((partner.astype(np.int64) - np.arange(n)) % n, bins=16, range=(0, n))[0]
    exp = hist.sum() / 16
    assert ((hist - exp) ** 2 / exp).sum() < 39.3  # 15 dof, p = 0.0006


def test_philox_flag_probability_and_z_moments(oracle):
    n, d = 4096, 16
    for prob in (0.5, 0.25, 0.2):
        _, z, flags, u = oracle.philox_sample_streams(1, 0, n, d, 1, 2.0, oracle.UFS_PROB, prob, 0)
        frac = flags.mean()
        assert abs(frac - prob) < 4 * math.sqrt(prob * (1 - prob) / (n * d)) + 1e-3
    # g(z) ~ 1/sqrt(z) on [1/a, a]: E[z] = (a^1.5 - a^-1.5) / (3 (a^.5 - a^-.5))
    a = 2.0
    ez = (a ** 1.5 - a ** -1.5) / (3 * (a ** 0.5 - a ** -0.5))
    assert abs(z.mean() - ez) < 0.03 and abs(u.mean() - 0.5) < 0.02


def test_mean_and_cov_match_reference_estimators(oracle):
    g = np.random.default_rng(0)
    x = g.normal(size=(200, 4))
    assert np.allclose(oracle.mean(x), x.mean(axis=0), rtol=1e-13, atol=1e-15)
    assert np.allclose(oracle.cov(x), np.cov(x.T, bias=True), rtol=1e-12, atol=1e-14)  # biased, /N (utils.rs:36)
