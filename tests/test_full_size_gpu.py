"""BASELINE.json's full-size configurations (c3, c4, c5) through size-independent properties: the oracle cannot
replay 2^22 walkers in seconds, so these check invariants that hold at any size -- the cached log-prob is the
log-density of the resident position for every walker, a ladder swap permutes (never alters) walkers, counters add
up, the chain is reproducible from (seed, counters), and the leading rungs of a tempered ensemble evolve exactly as
they do alone.  Ensembles are initialised and summarised on the device; only log-probs (8 bytes per walker) and
small statistics cross PCIe."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

C3 = dict(kind=7, n=1 << 20, dim=32, nb=1, ufs=0.5, box=(-6.0, 6.0))
C4 = dict(kind=3, n=1 << 20, dim=10, nb=64, ufs=0.2, box=(-0.01, 0.01))
C5 = dict(kind=2, n=1 << 22, dim=64, nb=1, ufs=0.5, box=(0.9, 1.1))


def _logprob(sb, cfg):
    if cfg["kind"] == 7:
        m = [5.0] + [0.0] * (cfg["dim"] - 1)
        return sb.Mixture(1.0, 2.0, m)
    return {2: sb.Rosenbrock(), 3: sb.Rastrigin()}[cfg["kind"]]


def _beta(nb):
    return 2.0 ** (-np.arange(nb, dtype=np.float64) / 8.0)


def _lp_tol(cfg, lp):
    # 1e-12 of the sum of |terms| (tests/util.py): bounded here by |lp| plus the targets' constant offsets
    return 1e-12 * (np.abs(lp) + 40.0 * cfg["dim"] + 1.0)


@pytest.mark.parametrize("cfg", [C3, C4, C5], ids=["c3", "c4", "c5"])
def test_cached_logprob_is_the_logdensity_of_the_resident_position(sb, cfg):
    """After stretch steps, swaps, t-walk steps and (where the plugin has a gradient) HMC steps, recomputing the
    log-density of every resident walker (K2) reproduces the cached value: every accepted row and its log-prob went
    out together, no rejected row leaked, for all 2^20 / 2^22 walkers."""
    n, dim, nb = cfg["n"], cfg["dim"], cfg["nb"]
    beta = _beta(nb)
    with sb.Sampler(_logprob(sb, cfg), n, dim, seed=5) as s:
        s.init_ensemble(*cfg["box"])
        for k in range(3):
            if nb > 1:
                s.swap_walkers(beta)
            s.sample_pt(2.0, sb.UpdateFlagSpec.Prob(cfg["ufs"]), beta, 3)
        s.twalk_sample(sb.TWalkParams.new(dim), beta, 2)
        eps = np.full(nb, {2: 0.002, 3: 0.02, 7: 0.2}[cfg["kind"]])
        s.hmc_sample_ensemble_pt(eps, beta, 3, sb.HmcParam.fixed(0.8), 2)
        st = s.stats()
        assert st["proposed"] == n * (9 + 2 + 2) and 0.02 * st["proposed"] < st["accepted"] < st["proposed"]
        cached = s.download_logprob()
        s.init_logprob()
        fresh = s.download_logprob()
    assert np.all(np.isfinite(cached))
    assert np.all(np.abs(cached - fresh) <= _lp_tol(cfg, fresh)), np.abs(cached - fresh).max()


def test_ladder_swap_permutes_walkers_at_full_size(sb):
    """c4: 64 rungs x 16384 walkers.  swap_walkers moves (row, log-prob) pairs between adjacent rungs and alters
    nothing: the sorted log-probs and the column sums of the ensemble are unchanged, the cached log-probs still
    belong to their rows, and proposals = (rungs - 1) x walkers per rung."""
    cfg = C4
    n, dim, nb = cfg["n"], cfg["dim"], cfg["nb"]
    beta = _beta(nb)
    with sb.Sampler(_logprob(sb, cfg), n, dim, seed=9) as s:
        s.init_ensemble(-0.5, 0.5)
        s.sample_pt(2.0, sb.UpdateFlagSpec.Prob(0.2), beta, 5)
        before, mean_before = s.download_logprob(), s.mean()
        s.swap_walkers(beta)
        after, mean_after = s.download_logprob(), s.mean()
        st = s.rung_stats(nb)
        s.init_logprob()
        fresh = s.download_logprob()
    assert np.array_equal(np.sort(before), np.sort(after)) and not np.array_equal(before, after)
    assert np.allclose(mean_before, mean_after, rtol=0, atol=1e-12)
    assert np.all(st["swap_proposed"] == n // nb) and np.all(st["swapped"] > 0) and np.all(st["swapped"] <= n // nb)
    assert np.all(np.abs(after - fresh) <= _lp_tol(cfg, fresh))


def test_leading_rungs_evolve_as_they_do_alone(sb):
    """c4 at full size against its first four rungs as an ensemble of their own (same seed): streams are keyed by
    walker and rung, the matching never leaves a rung, so the leading rungs must be bit-identical -- until a swap
    couples them to rung 4."""
    cfg = C4
    n, dim, nb, sub = cfg["n"], cfg["dim"], cfg["nb"], 4
    npb = n // nb
    beta = _beta(nb)
    f = _logprob(sb, cfg)
    with sb.Sampler(f, n, dim, seed=21) as big, sb.Sampler(f, sub * npb, dim, seed=21) as small:
        for s, b in ((big, beta), (small, beta[:sub])):
            s.init_ensemble(-0.5, 0.5)
            s.sample_pt(2.0, sb.UpdateFlagSpec.Prob(0.2), b, 6)
            s.twalk_sample(sb.TWalkParams.new(dim), b, 2)
        eb, lb = big.download()
        es, ls = small.download()
    assert np.array_equal(eb[:sub * npb], es) and np.array_equal(lb[:sub * npb], ls)


def test_full_size_chain_is_reproducible_from_seed_and_counters(sb):
    """c5: two contexts with the same seed agree bit for bit on the device-side summaries after the same calls;
    restoring (counters) on a third one and replaying the tail from a downloaded state does too (checkpoint)."""
    cfg = C5
    n, dim = cfg["n"], cfg["dim"]
    f = _logprob(sb, cfg)
    outs = []
    for _ in range(2):
        with sb.Sampler(f, n, dim, seed=77) as s:
            s.init_ensemble(*cfg["box"])
            s.sample(2.0, sb.UpdateFlagSpec.Prob(0.5), 4)
            outs.append((s.mean(), s.stats()["accepted"], s.download_logprob()))
    assert np.array_equal(outs[0][0], outs[1][0]) and outs[0][1] == outs[1][1] and np.array_equal(outs[0][2], outs[1][2])
    assert 0.05 * 4 * n < outs[0][1] < 0.9 * 4 * n
