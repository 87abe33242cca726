"""The oracle's chain statistics (oracle/chain_stats.c) against independent numpy restatements."""
import numpy as np


def _ar1(T, ncap, phi, seed):
    g = np.random.default_rng(seed)
    x = np.zeros((T, ncap))
    e = g.standard_normal((T, ncap))
    for t in range(1, T):
        x[t] = phi * x[t - 1] + e[t]
    return x + g.normal(size=ncap) * 3.0


def _numpy_iat(x, c, max_lag):
    T = x.shape[0]
    L = T - 1 if max_lag == 0 or max_lag > T - 1 else max_lag
    xc = x - x.mean()
    acov = np.array([np.dot(xc[:T - k], xc[k:]) / T for k in range(L + 1)])
    tau = 1.0 + 2.0 * np.cumsum(acov[1:] / acov[0])
    k = np.arange(1, L + 1)
    hit = np.nonzero(k >= c * tau)[0]
    m = hit[0] if hit.size else L - 1
    return tau[m], k[m], x.mean(), acov[0]


def test_iat_matches_numpy_and_theory(oracle):
    x = _ar1(6000, 4, 0.8, 0)
    tau, window, mean, var = oracle.trace_iat(x, 5.0, 0)
    for c in range(4):
        wt, wk, wm, wv = _numpy_iat(x[:, c], 5.0, 0)
        assert abs(tau[c] - wt) <= 1e-10 * wt and window[c] == wk
        assert abs(mean[c] - wm) <= 1e-12 * (abs(wm) + 1) and abs(var[c] - wv) <= 1e-12 * wv
    # AR(1): tau = (1 + phi) / (1 - phi) = 9
    assert np.all(np.abs(tau - 9.0) < 2.5)
    tau_s, window_s, _, _ = oracle.trace_iat(x, 5.0, 10)     # a window that cannot close in 10 lags
    assert np.all(window_s == 10) and np.all(tau_s < 10.0 * 5.0)


def test_iat_of_white_noise_is_one(oracle):
    x = np.random.default_rng(5).standard_normal((20000, 3))
    tau, window, _, _ = oracle.trace_iat(x, 5.0, 0)
    assert np.all(np.abs(tau - 1.0) < 0.15) and np.all(window <= 8)


def test_running_moments_match_numpy(oracle):
    g = np.random.default_rng(2)
    s = g.normal(size=(7, 40, 5)) * g.uniform(0.5, 2.0, 5) + 10.0
    m, c = oracle.running_moments(s, True)
    flat = s.reshape(-1, 5)
    assert np.allclose(m, flat.mean(axis=0), rtol=1e-13)
    assert np.allclose(c, np.cov(flat.T, bias=True), rtol=1e-11)
    # the same thing the reference's mean / cov give on the concatenated rows
    assert np.allclose(m, oracle.mean(flat), rtol=1e-13) and np.allclose(c, oracle.cov(flat), rtol=1e-12)
    m2, c2 = oracle.running_moments(s, False)
    assert c2 is None and np.array_equal(m, m2)
