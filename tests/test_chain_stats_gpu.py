"""Chain statistics on the device against the oracle: per-rung acceptance and
exchange counts, running per-rung moments (src/linear_space/utils.rs:4-41 over steps x walkers), thinned
capture and the integrated autocorrelation time of the captured series."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ladder(nb):
    return 2.0 ** (-np.arange(nb, dtype=np.float64) / 2.0)


@pytest.mark.parametrize("nb,npb,dim,kernel", [(4, 64, 5, None), (3, 48, 7, None), (5, 16, 3, None), (4, 64, 5, "seq"),
                                               (3, 48, 300, None)])
def test_rung_counts_match_accept_and_swap_masks(sb, oracle, nb, npb, dim, kernel, monkeypatch):
    """Per-rung accepted moves == sums of the accept mask, per-stage exchanges == sums of the swap mask
    (npb = 48: tiles that straddle a rung boundary; dim = 300: the cooperative kernel)."""
    if kernel:
        monkeypatch.setenv("SCORUS_KERNEL", kernel)
    n = nb * npb
    g = np.random.default_rng(nb * 1000 + npb)
    beta = _ladder(nb)
    ens = np.ascontiguousarray(g.uniform(-0.5, 0.5, (n, dim)))
    with sb.Sampler(sb.Rastrigin(), n, dim, seed=9) as s:
        s.upload(ens, None)
        acc_want, sw_want = np.zeros(nb, dtype=np.uint64), np.zeros(nb - 1, dtype=np.uint64)
        rng = oracle.Rng(77)
        for step in range(6):
            partner, z, flags, u, _ = oracle.draw_sample_streams(rng, n, dim, nb, 2.0, oracle.UFS_PROB, 0.5)
            acc = s.sample_pt_fixed(beta, partner, z, flags, u)
            acc_want += acc.reshape(nb, npb).sum(axis=1).astype(np.uint64)
            if step % 2 == 1:
                jvec, r = oracle.draw_swap_streams(rng, n, nb)
                sw = s.swap_walkers_fixed(beta, jvec, r).reshape(nb - 1, npb)
                # stage 0 is the hottest pair (rungs nb-1, nb-2); entry k of the counts = rungs k, k+1
                sw_want += sw.sum(axis=1)[::-1].astype(np.uint64)
        st = s.rung_stats(nb)
        tot = s.stats()
    assert np.array_equal(st["accepted"], acc_want)
    assert np.array_equal(st["proposed"], np.full(nb, 6 * npb, dtype=np.uint64))
    assert np.array_equal(st["swapped"], sw_want)
    assert np.array_equal(st["swap_proposed"], np.full(nb - 1, 3 * npb, dtype=np.uint64))
    assert st["accepted"].sum() == tot["accepted"] and st["swapped"].sum() == tot["swapped"]
    assert 0 < acc_want.sum() < 6 * n


def test_rung_counts_in_device_stream_mode_and_single_rung(sb):
    nb, npb, dim = 8, 256, 10
    n = nb * npb
    beta = _ladder(nb)
    with sb.Sampler(sb.Rastrigin(), n, dim, seed=3) as s:
        s.init_ensemble(-0.01, 0.01)
        for k in range(30):
            if k % 10 == 0:
                s.swap_walkers(beta)
            s.sample_pt(2.0, sb.UpdateFlagSpec.Prob(0.2), beta, 1)
        st, tot = s.rung_stats(nb), s.stats()
        assert st["accepted"].sum() == tot["accepted"] and np.all(st["proposed"] == 30 * npb)
        assert st["swapped"].sum() == tot["swapped"] and np.all(st["swap_proposed"] == 3 * npb)
        rate = st["accepted"] / st["proposed"]
        assert np.all(rate > 0.02) and np.all(rate < 0.98)
        # single-rung calls land in rung 0
        before = s.rung_stats(1)["accepted"][0]
        s.sample(2.0, sb.UpdateFlagSpec.All(), 5)
        after = s.rung_stats(1)
        assert after["accepted"][0] - before == s.stats()["accepted"] - tot["accepted"]
        assert after["proposed"][0] == 30 * npb + 5 * n


@pytest.mark.parametrize("nb,npb,dim,every,with_cov", [(1, 512, 6, 1, True), (4, 96, 9, 2, True), (2, 64, 130, 3, False),
                                                       (64, 32, 10, 1, True)])
def test_running_moments_match_oracle_over_snapshots(sb, oracle, nb, npb, dim, every, with_cov):
    n, steps = nb * npb, 12
    beta = _ladder(nb)
    g = np.random.default_rng(dim)
    ens = np.ascontiguousarray(g.normal(size=(n, dim)) * 0.3 + 50.0)   # far from the origin: the shift matters
    snaps = []
    with sb.Sampler(sb.Gaussian(1.0), n, dim, seed=11) as s:
        s.upload(ens - 50.0, None)
        s.moments_enable(nb, every, with_cov)
        with pytest.raises(ValueError):
            s.moments(0)                       # nothing accumulated yet
        for k in range(steps):
            s.sample_pt(2.0, sb.UpdateFlagSpec.Prob(0.5), beta, 1)
            if (k + 1) % every == 0:
                snaps.append(s.download()[0])
        snaps = np.array(snaps)
        for r in range(nb):
            m, c, cnt = s.moments(r)
            sub = np.ascontiguousarray(snaps[:, r * npb:(r + 1) * npb, :])
            wm, wc = oracle.running_moments(sub, with_cov)
            assert cnt == len(snaps) * npb
            sd = sub.reshape(-1, dim).std(axis=0)
            assert np.all(np.abs(m - wm) <= 1e-12 * (np.abs(wm) + sd))
            if with_cov:
                assert np.all(np.abs(c - wc) <= 1e-11 * np.outer(sd, sd))
            else:
                assert c is None
        with pytest.raises(ValueError):
            s.moments(nb)
        s.moments_enable(0)
        with pytest.raises(ValueError):
            s.moments(0)


def test_running_moments_survive_a_large_offset(sb, oracle):
    """Sums are taken about the first snapshot's rung mean: a mean of 1e6 with unit spread keeps 1e-9 accuracy
    on the covariance (raw second moments would lose it all)."""
    n, dim = 1024, 4
    g = np.random.default_rng(1)
    ens = np.ascontiguousarray(g.normal(size=(n, dim)) + 1e6)
    with sb.Sampler(sb.Gaussian(1.0), n, dim, seed=1) as s:
        s.upload(ens, None)
        s.moments_enable(1, 1, True)
        snaps = []
        for _ in range(3):
            s.sample(1.01, sb.UpdateFlagSpec.All(), 1)
            snaps.append(s.download()[0])
        m, c, cnt = s.moments(0)
        snaps = np.array(snaps)
    wm, wc = oracle.running_moments(snaps, True)
    assert cnt == 3 * n
    assert np.allclose(m, wm, rtol=1e-14, atol=0) and np.allclose(c, wc, rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("stride", [1, 3])
def test_iat_on_device_is_bit_identical_to_oracle(sb, oracle, stride):
    nb, npb, dim, steps = 2, 64, 4, 900
    n = nb * npb
    beta = _ladder(nb)
    walkers = [0, 1, npb, npb + 5, n - 1]
    coords = [0, 3, 0, 2, 1]
    with sb.Sampler(sb.Gaussian(1.0), n, dim, seed=21) as s:
        s.init_ensemble(-1.0, 1.0)
        s.trace_enable(walkers, coords, steps)
        s.trace_stride(stride)
        s.sample_pt(2.0, sb.UpdateFlagSpec.Prob(0.5), beta, steps)
        got = s.trace_iat(5.0, 0)
        got_short = s.trace_iat(4.0, 50)
        trace = s.trace_download(steps)
    assert trace.shape == (steps // stride, len(walkers))
    for res, (c, lag) in ((got, (5.0, 0)), (got_short, (4.0, 50))):
        tau, window, mean, var = oracle.trace_iat(trace, c, lag)
        assert np.array_equal(res["tau"], tau) and np.array_equal(res["window"], window)
        assert np.array_equal(res["mean"], mean) and np.array_equal(res["var"], var)
    assert np.all(got["tau"] > 1.0) and np.all(got["window"] < trace.shape[0] - 1)
    # thinning by 3 shortens the autocorrelation time measured in recorded rows
    if stride == 3:
        assert np.median(got["tau"]) < 30


def test_trace_stride_keeps_every_kth_step(sb):
    n, dim, steps = 32, 3, 20
    with sb.Sampler(sb.Gaussian(), n, dim, seed=2) as a, sb.Sampler(sb.Gaussian(), n, dim, seed=2) as b:
        for s in (a, b):
            s.init_ensemble(-1.0, 1.0)
            s.trace_enable([0, 5], [1, 2], steps)
        b.trace_stride(4)
        a.sample(2.0, sb.UpdateFlagSpec.All(), steps)
        b.sample(2.0, sb.UpdateFlagSpec.All(), steps)
        full, thin = a.trace_download(steps), b.trace_download(steps)
    assert np.array_equal(thin, full[3::4])


def test_chain_statistics_errors(sb):
    with sb.Sampler(sb.Gaussian(), 32, 3) as s:
        s.init_ensemble(-1.0, 1.0)
        with pytest.raises(ValueError):
            s._trace_ncap = 1
            s.trace_iat()                       # no trace enabled
        with pytest.raises(sb.McmcErr):
            s.moments_enable(5)                 # 32 walkers do not split into 5 rungs
        with pytest.raises(ValueError):
            s.moments_enable(2, 0)
        s.trace_enable([0], [0], 4)
        s.sample(2.0, sb.UpdateFlagSpec.All(), 1)
        with pytest.raises(ValueError):
            s.trace_iat()                       # one row only
    with sb.Sampler(sb.Gaussian(), 32, 200) as s:
        with pytest.raises(ValueError):
            s.moments_enable(1, 1, True)        # covariance limited to dim <= 128
