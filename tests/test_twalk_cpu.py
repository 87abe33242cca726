"""The t-walk oracle (oracle/twalk.c, a restatement of src/mcmc/twalk.rs:586-762) against the committed golden
vectors, the independent numpy restatement, and the moments the move must reproduce."""
import numpy as np
import pytest

from oracle import numpy_oracle as no
from tests.util import twalk_golden_files


def _tw(oracle, g):
    t = oracle.twalk_params(int(g["ens0"].shape[1]))
    t.aw, t.at, t.pphi = g["tw"][0], g["tw"][1], g["tw"][2]
    t.fw[0], t.fw[1] = g["tw"][3], g["tw"][4]
    return t


@pytest.mark.parametrize("path", twalk_golden_files(), ids=lambda p: p.split("/")[-1][:-4])
def test_golden_vectors_both_restatements(oracle, path):
    g = np.load(path)
    kind, params, beta = int(g["kind"]), g["params"], g["beta"]
    tw = _tw(oracle, g)
    ens, lp = g["ens0"].copy(), g["lp0"].copy()
    for k in range(g["order"].shape[0]):
        for half in (0, 1):
            args = (g["order"][k], half, g["kernel"][k, half], g["b"][k, half], g["flags"][k, half], g["walk_u"][k, half],
                    g["u"][k, half])
            e2, l2, a2 = no.twalk_half_fixed(kind, params, ens, lp, beta, tw.aw, *args)
            acc = oracle.twalk_half_fixed(kind, params, ens, lp, beta, tw, *args)
            assert np.array_equal(acc, g["accepted"][k, half]) and np.array_equal(a2, acc)
            assert np.array_equal(ens, g["ens"][k, half]) and np.array_equal(e2, ens)
            assert np.array_equal(lp, g["lp"][k, half], equal_nan=True) and np.array_equal(l2, lp, equal_nan=True)


def test_params_default_and_b(oracle):
    for dim, pphi in ((1, 1.0), (3, 1.0), (4, 1.0), (10, 0.4), (64, 0.0625)):
        t = oracle.twalk_params(dim)
        assert (t.aw, t.at, t.pphi, t.fw[0], t.fw[1]) == (1.5, 6.0, pphi, 0.5, 1.0)   # TWalkParams::new, twalk.rs:89-110
    L = oracle.lib()
    assert L.orc_twalk_b_from_uniforms(0.0, 0.3, 6.0) == 0.0                          # :315-316
    assert L.orc_twalk_b_from_uniforms(0.5, 0.1, 6.0) == 0.5 ** (1.0 / 7.0)           # v < 5/12 -> x^(1/(at+1)), :317-318
    assert L.orc_twalk_b_from_uniforms(0.5, 0.9, 6.0) == 0.5 ** (1.0 / -5.0)          # else x^(1/(1-at)) > 1, :319-320
    for x, v in ((0.2, 0.2), (0.7, 0.6)):
        assert L.orc_twalk_b_from_uniforms(x, v, 6.0) == no.twalk_b(x, v, 6.0)


def test_reference_order_streams(oracle):
    """orc_twalk_draw_streams consumes the generator as twalk::sample does: shuffles, then per half-pass all proposal
    draws and then all accept uniforms; pairs stay inside their rung; Walk pairs carry uniforms on flagged
    coordinates only, Traverse pairs none."""
    n, dim, nb = 48, 7, 3
    tw = oracle.twalk_params(dim)
    order, kernel, b, flags, wu, u = oracle.twalk_draw_streams(oracle.Rng(5), n, dim, nb, tw)
    npb = n // nb
    assert sorted(order.tolist()) == list(range(n))
    assert np.array_equal(order // npb, np.arange(n) // npb)
    assert flags.any(axis=2).all()                                   # gen_update_flags redraws until one is set
    walk = kernel == 0
    assert np.all((wu != 0.0) <= (flags.astype(bool) & walk[:, :, None]))
    assert np.all(b[walk] == 0.0) and np.all(b[~walk] > 0.0)
    assert np.all((u >= 0.0) & (u < 1.0)) and 0.2 < walk.mean() < 0.8
    # same seed, same draws; the RNG-driven wrapper equals draw + two fixed half-passes
    g = np.random.default_rng(0)
    ens = np.ascontiguousarray(g.normal(size=(n, dim)))
    lp = oracle.init_logprob(0, [0.5], ens)
    beta = 2.0 ** -np.arange(nb, dtype=np.float64)
    e1, l1 = ens.copy(), lp.copy()
    assert oracle.twalk_sample(0, [0.5], e1, l1, oracle.Rng(5), tw, beta) == 0
    for half in (0, 1):
        oracle.twalk_half_fixed(0, [0.5], ens, lp, beta, tw, order, half, kernel[half], b[half], flags[half], wu[half], u[half])
    assert np.array_equal(e1, ens) and np.array_equal(l1, lp)


def test_oracle_twalk_samples_tempered_gaussian(oracle):
    """-x^2/2 at rungs beta = 1, 1/4: variance 1 and 4 (averaged over the chain)."""
    n, dim = 32, 4
    tw = oracle.twalk_params(dim)
    g = np.random.default_rng(0)
    ens = np.ascontiguousarray(g.normal(size=(n, dim)))
    lp = oracle.init_logprob(0, [0.5], ens)
    rng = oracle.Rng(3)
    beta = [1.0, 0.25]
    s2, cnt = np.zeros(2), 0
    for k in range(6000):
        oracle.twalk_sample(0, [0.5], ens, lp, rng, tw, beta)
        if k > 300:
            s2 += [(ens[:16] ** 2).mean(), (ens[16:] ** 2).mean()]
            cnt += 1
    assert abs(s2[0] / cnt - 1.0) < 0.08 and abs(s2[1] / cnt - 4.0) < 0.35
    assert np.allclose(lp, oracle.init_logprob(0, [0.5], ens), rtol=1e-12)


def test_argument_checks(oracle):
    tw = oracle.twalk_params(2)
    ens, lp = np.zeros((6, 2)), np.zeros(6)
    assert oracle.twalk_sample(0, [0.5], ens, lp, oracle.Rng(1), tw, [1.0, 0.5, 0.25, 0.125]) == oracle.ERR_NWALKERS_MISMATCHES_NBETA
    assert oracle.twalk_sample(0, [0.5], ens, lp, oracle.Rng(1), tw, [1.0, 0.5]) == oracle.ERR_NWALKERS_IS_NOT_EVEN


def test_literal_index_reading_of_the_reference_parts_from_the_shipped_one_at_two_rungs(oracle):
    """twalk.rs:605-612 builds every rung's pairs from rung-LOCAL indices; sample1 reads walkers and log-probs with
    them unshifted (:686-690, :741-749) and writes shifted by ibeta * nwalkers_per_beta (:757-758).  The shipped
    reading (oracle/twalk.c: orc_twalk_half_fixed, the device kernels) shifts both sides; the literal one
    (orc_twalk_half_fixed_literal) follows the Rust as written.  This pins where they part:
      * one rung: bit-identical -- parity with the reference as written holds there;
      * two rungs, same draws: rung 0 evolves identically in both readings, rung 1 does not; under the literal
        reading every row accepted into rung 1 is a proposal built from rung 0's walkers, and rung 1 never moves off
        them: its own starting positions play no part."""
    rng = np.random.default_rng(4)
    dim = 6
    tw = oracle.twalk_params(dim)
    for nb in (1, 2):
        npb = 64
        n = nb * npb
        beta = 2.0 ** -np.arange(nb, dtype=np.float64)
        ens0 = np.ascontiguousarray(rng.normal(size=(n, dim)))
        if nb == 2:
            ens0[npb:] += 50.0      # rung 1 starts far away from rung 0: the two readings cannot agree by accident
        lp0 = oracle.init_logprob(0, [0.5], ens0)
        g = oracle.Rng(11)
        e_s, l_s, e_l, l_l = ens0.copy(), lp0.copy(), ens0.copy(), lp0.copy()
        for step in range(6):
            order, kernel, b, flags, wu, u = oracle.twalk_draw_streams(g, n, dim, nb, tw)
            for half in (0, 1):
                a = (order, half, kernel[half], b[half], flags[half], wu[half], u[half])
                acc_s, prop_s, _ = oracle.twalk_half_fixed(0, [0.5], e_s, l_s, beta, tw, *a, want_proposals=True)
                pre_l = e_l.copy()
                acc_l, prop_l, _ = oracle.twalk_half_fixed(0, [0.5], e_l, l_l, beta, tw, *a, want_proposals=True,
                                                           literal_index=True)
                if nb == 2:
                    # literal proposals of rung 1's pairs are functions of rung 0's rows only
                    movers = order[half::2]
                    r1 = movers >= npb
                    src = pre_l[movers[r1] - npb]
                    assert np.all(np.abs(prop_l[r1] - src).max(axis=1) < 40.0), "literal proposals come from rung 0"
                    assert np.all(np.abs(prop_s[r1]).max(axis=1) > 40.0), "shipped proposals stay with rung 1's walkers"
        if nb == 1:
            assert np.array_equal(e_s, e_l) and np.array_equal(l_s, l_l)
        else:
            assert np.array_equal(e_s[:npb], e_l[:npb]) and np.array_equal(l_s[:npb], l_l[:npb])   # rung 0: the same chain
            assert not np.array_equal(e_s[npb:], e_l[npb:])
            moved = np.any(e_l[npb:] != ens0[npb:], axis=1)
            assert moved.any() and np.all(np.abs(e_l[npb:][moved]).max(axis=1) < 40.0)    # accepted rows sit in rung 0's cloud
            assert np.all(np.abs(e_s[npb:]).max(axis=1) > 40.0)                            # shipped: rung 1 stays its own chain
