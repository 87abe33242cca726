"""GPU parity of the ensemble / parallel-tempering HMC (src/mcmc/hmc/naive.rs:122-292, src/mcmc/hmc/utils.rs:6-30;
SURVEY.md 8(f) rank 3) against oracle/hmc.c: with identical momenta and uniforms the accept decisions and the step-size
adaptation are bit-identical, positions are bit-identical for transcendental-free forces (1e-12 with sin() in the
force: Rastrigin), log-probs agree within 1e-12 of their sum of |terms|."""
import numpy as np
import pytest

from tests.util import assert_lp_close, hmc_golden_files

pytestmark = pytest.mark.gpu

CASES = [
    # kind, params, n, dim, nbeta, l, epsilon, start
    (0, [0.5], 64, 2, 1, 5, 0.2, (-1, 1)),
    (0, [1.0], 3, 1, 1, 3, 0.2, (-1, 1)),          # odd ensemble, one coordinate: nothing pairs up in HMC
    (0, [0.5], 1000, 7, 4, 4, 0.15, (-1, 1)),      # odd dim
    (2, [], 512, 2, 1, 2, 0.005, (0.8, 1.2)),      # examples/test_hmc.rs: 2-D Rosenbrock, l = 2, epsilon = 0.005
    (2, [], 2048, 64, 2, 3, 0.002, (0.9, 1.1)),    # one chunk per lane, neighbour shuffles
    (2, [], 96, 200, 4, 2, 0.001, (0.9, 1.1)),     # four chunks per lane
    (2, [], 70, 33, 1, 3, 0.003, (0.9, 1.1)),      # odd dim, ragged last trip
    (2, [], 300, 100, 3, 2, 0.002, (0.9, 1.1)),    # two chunks per lane
    (3, [10.0], 4096, 10, 16, 5, 0.02, (-0.5, 0.5)),
    (1, [1.5], 512, 5, 2, 6, 0.3, (-1.4, 1.4)),    # trajectories that leave the box: lp = -inf, dh not finite
    (7, [1.0, 2.0, 3.0] + [0.0] * 31, 1024, 32, 2, 4, 0.4, (-4.0, 4.0)),   # mixture: row-wide sums before every force
    (7, [0.5, 1.5, 2.0, -1.0, 0.5], 256, 3, 1, 5, 0.3, (-3.0, 3.0)),
    # the dense Gaussian -1/2 |L (x - mu)|^2 (BASELINE config 2's target): two matrix-vector products per force
    (6, "corr", 256, 12, 2, 4, 0.1, (-1.0, 1.0)),      # 8 lanes per walker
    (6, "corr", 96, 100, 1, 3, 0.05, (-1.0, 1.0)),     # 32 lanes, two chunks per lane
    (6, "corr", 70, 7, 1, 5, 0.1, (-1.0, 1.0)),        # odd dim: the last chunk is half empty
]


def _corr_params(dim):
    """mu and a well-conditioned dense factor L (row-major) for SCORUS_LP_CORR_GAUSSIAN."""
    g = np.random.default_rng(1000 + dim)
    L = np.eye(dim) * 1.5 + 0.3 * g.normal(size=(dim, dim)) / np.sqrt(dim)
    return np.concatenate([g.uniform(-0.2, 0.2, dim), L.ravel()]).tolist()


@pytest.mark.parametrize("path", hmc_golden_files(), ids=lambda p: p.split("/")[-1][:-4])
def test_golden_fixed_stream(sb, path):
    g = np.load(path)
    kind, params, beta, l = int(g["kind"]), g["params"], g["beta"], int(g["l"])
    param = sb.HmcParam.new(float(g["target"]), float(g["adj"]))
    n, dim = g["q0"].shape
    eps = g["eps0"].copy()
    with sb.Sampler(sb.LogProb(kind, tuple(params.tolist())), n, dim, seed=0) as s:
        s.upload(np.ascontiguousarray(g["q0"]), np.ascontiguousarray(g["lp0"]))
        for k in range(g["p0"].shape[0]):
            acc, cnt = s.hmc_sample_ensemble_pt_fixed(eps, beta, l, param, g["p0"][k], g["u1"][k], g["u2"][k])
            q, lp = s.download()
            assert np.array_equal(acc, g["accepted"][k]) and np.array_equal(cnt, g["cnt"][k]), f"decisions differ at step {k}"
            assert np.array_equal(q, g["q"][k]), f"positions differ at step {k}"
            assert np.array_equal(eps, g["eps"][k]), f"epsilon differs at step {k}"
            assert_lp_close(kind, params, q, lp, g["lp"][k], exact=False)


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"k{c[0]}_n{c[2]}_d{c[3]}_b{c[4]}_l{c[5]}")
def test_fixed_stream_vs_oracle(sb, oracle, case):
    kind, params, n, dim, nb, l, eps0, box = case
    if params == "corr":
        params = _corr_params(dim)
    g = np.random.default_rng(n + dim)
    beta = 2.0 ** (-np.arange(nb, dtype=np.float64) / 2.0)
    q = np.ascontiguousarray(g.uniform(box[0], box[1], (n, dim)))
    lp = oracle.init_logprob(kind, params, q)
    eps_o = np.full(nb, eps0)
    eps_d = eps_o.copy()
    param = sb.HmcParam.quick_adj(0.7)
    total = 0
    with sb.Sampler(sb.LogProb(kind, tuple(params)), n, dim, seed=1) as s:
        s.upload(q, lp)
        for step in range(4):
            p0, u1, u2 = g.normal(size=(n, dim)), g.uniform(size=n), g.uniform(size=n)
            acc, cnt = s.hmc_sample_ensemble_pt_fixed(eps_d, beta, l, param, p0, u1, u2)
            want, want_cnt = oracle.hmc_ensemble_pt_fixed(kind, params, q, lp, eps_o, beta, l, 0.7, 0.01, p0, u1, u2)
            got_q, got_lp = s.download()
            assert np.array_equal(acc, want), f"accept mask differs at step {step}"
            assert np.array_equal(cnt, want_cnt)
            if kind in (3, 7):   # sin() / exp() in the force: the device's and glibc's differ by an ulp now and then
                assert np.allclose(got_q, q, rtol=1e-12, atol=1e-13), f"positions differ at step {step}: {np.abs(got_q - q).max()}"
            else:
                assert np.array_equal(got_q, q), f"positions differ at step {step}: {np.abs(got_q - q).max()}"
            assert np.array_equal(eps_d, eps_o), f"epsilon differs at step {step}"
            assert_lp_close(kind, params, q, got_lp, lp, exact=False)
            total += int(acc.sum())
        assert 0 < total <= 4 * n
        if kind == 1:
            assert total < 4 * n          # some trajectories left the box


@pytest.mark.parametrize("kind,params,n,dim,nb,l,eps0", [(0, [0.5], 256, 6, 2, 5, 0.2), (2, [], 1024, 20, 2, 3, 0.003),
                                                          (3, [10.0], 2048, 10, 8, 4, 0.02)])
def test_device_streams_replay_through_oracle(sb, oracle, kind, params, n, dim, nb, l, eps0):
    """sample_ensemble_pt on the device streams == the oracle fed the device's own momenta (they go through
    log / sqrt / sincos) and the CPU restatement of the uniform stream."""
    g = np.random.default_rng(3)
    beta = 2.0 ** (-np.arange(nb, dtype=np.float64))
    q = np.ascontiguousarray(g.uniform(0.8, 1.2, (n, dim)) if kind == 2 else g.uniform(-0.5, 0.5, (n, dim)))
    lp = oracle.init_logprob(kind, params, q)
    eps_o = np.full(nb, eps0)
    eps_d = eps_o.copy()
    param = sb.HmcParam.new(0.8, 0.005)
    seed = 77
    with sb.Sampler(sb.LogProb(kind, tuple(params)), n, dim, seed=seed) as s:
        s.upload(q, lp)
        for step in range(4):
            c0 = s.counters[0]
            p0 = s.hmc_draw_momenta(c0)
            cnt = s.hmc_sample_ensemble_pt(eps_d, beta, l, param, 1)
            assert s.counters[0] == c0 + 1
            u1, u2 = _zu_uniforms(oracle, seed, c0, n)
            _, want_cnt = oracle.hmc_ensemble_pt_fixed(kind, params, q, lp, eps_o, beta, l, 0.8, 0.005, p0, u1, u2)
            got_q, got_lp = s.download()
            assert np.array_equal(cnt, want_cnt) and np.array_equal(eps_d, eps_o)
            assert np.allclose(got_q, q, rtol=1e-12, atol=1e-13) if kind == 3 else np.array_equal(got_q, q)
            assert_lp_close(kind, params, q, got_lp, lp, exact=False)


def test_device_momenta_are_standard_normal(sb):
    with sb.Sampler(sb.Gaussian(), 65536, 5, seed=6) as s:
        m = s.hmc_draw_momenta(123)
        again = s.hmc_draw_momenta(123)
        other = s.hmc_draw_momenta(124)
    assert np.array_equal(m, again) and not np.array_equal(m, other)
    assert abs(m.mean()) < 0.01 and abs(m.var() - 1.0) < 0.01 and abs((m ** 4).mean() - 3.0) < 0.06
    assert abs(np.corrcoef(m[:, 0], m[:, 1])[0, 1]) < 0.02 and abs(np.corrcoef(m[:-1, 4], m[1:, 0])[0, 1]) < 0.02


def _zu_uniforms(oracle, seed, step, n):
    """Stream ZU = 1 of the device spec at this counter: (r0, r1) -> accept uniform,
    (r2, r3) -> adaptation uniform, keyed by the walker."""
    u1, u2 = np.empty(n), np.empty(n)
    for w in range(n):
        r = oracle.philox4x32_10([w, 0, step & 0xFFFFFFFF, ((step >> 32) & 0xFFFFFF) | (1 << 24)], [seed & 0xFFFFFFFF, seed >> 32])
        u1[w] = float(((int(r[1]) << 32) | int(r[0])) >> 11) * 2.0 ** -53
        u2[w] = float(((int(r[3]) << 32) | int(r[2])) >> 11) * 2.0 ** -53
    return u1, u2


def test_large_rungs_adapt_by_net_count(sb, oracle):
    """Rungs above 4096 walkers: epsilon (1 + adj)^(grown - shrunk) instead of the walker-by-walker replay; same
    decisions, epsilon within 1e-12 of the sequential loop's."""
    n, dim, nb, l = 2 * 8192, 4, 2, 3
    g = np.random.default_rng(11)
    beta = np.array([1.0, 0.5])
    q = np.ascontiguousarray(g.uniform(-1, 1, (n, dim)))
    lp = oracle.init_logprob(0, [0.5], q)
    eps_o = np.array([0.9, 1.3])
    eps_d = eps_o.copy()
    with sb.Sampler(sb.Gaussian(0.5), n, dim, seed=1) as s:
        s.upload(q, lp)
        for step in range(3):
            p0, u1, u2 = g.normal(size=(n, dim)), g.uniform(size=n), g.uniform(size=n)
            acc, cnt = s.hmc_sample_ensemble_pt_fixed(eps_d, beta, l, sb.HmcParam.new(0.7, 1e-3), p0, u1, u2)
            want, want_cnt = oracle.hmc_ensemble_pt_fixed(0, [0.5], q, lp, eps_o, beta, l, 0.7, 1e-3, p0, u1, u2)
            assert np.array_equal(acc, want) and np.array_equal(cnt, want_cnt) and np.array_equal(s.download()[0], q)
            assert np.allclose(eps_d, eps_o, rtol=1e-12, atol=0) and not np.array_equal(eps_d, [0.9, 1.3])
            eps_o[:] = eps_d          # keep the two chains on the same step size


def test_fixed_step_size_and_single_rung_counts(sb):
    n, dim = 512, 4
    eps = np.array([0.3])
    with sb.Sampler(sb.Gaussian(0.5), n, dim, seed=4) as s:
        s.init_ensemble(-1.0, 1.0)
        cnt = s.hmc_sample_ensemble_pt(eps, [1.0], 5, sb.HmcParam.fixed(0.8), 20)
        assert eps[0] == 0.3                                  # HmcParam::fixed never adapts
        st = s.stats()
        assert cnt[0] == st["accepted"] and st["proposed"] == 20 * n and 0.5 * 20 * n < cnt[0] <= 20 * n
        assert s.launch_count >= 20


def test_errors(sb):
    eps = np.array([0.1])
    with sb.Sampler(sb.Bimodal2d(), 8, 2) as s:
        s.init_ensemble(0.1, 0.9)
        with pytest.raises(ValueError):
            s.hmc_sample_ensemble_pt(eps, [1.0], 3, sb.HmcParam.default(), 1)     # no gradient for this plugin
    with sb.Sampler(sb.Gaussian(), 9, 2) as s:
        s.init_ensemble(-1.0, 1.0)
        with pytest.raises(sb.McmcErr):
            s.hmc_sample_ensemble_pt(np.array([0.1, 0.1]), [1.0, 0.5], 3, sb.HmcParam.default(), 1)   # 9 walkers, 2 rungs
        with pytest.raises(AssertionError):
            s.hmc_sample_ensemble_pt(np.array([0.1, 0.1]), [1.0], 3, sb.HmcParam.default(), 1)        # epsilon.len() != nbeta
        s.hmc_sample_ensemble_pt(np.array([0.1, 0.1, 0.1]), [1.0, 0.5, 0.25], 3, sb.HmcParam.default(), 1)   # odd rungs are fine


def test_module_level_call_matches_resident_chain(sb):
    n, dim, nb, l, steps = 96, 5, 3, 4, 6
    g = np.random.default_rng(9)
    q0 = np.ascontiguousarray(g.uniform(-1, 1, (n, dim)))
    beta = [1.0, 0.5, 0.25]
    f = sb.Rastrigin()
    param = sb.HmcParam.quick_adj(0.8)
    eps_a = np.full(nb, 0.02)
    with sb.Sampler(f, n, dim, seed=8) as s:
        s.upload(q0, None)
        want_cnt = np.zeros(nb, dtype=np.uint64)
        for _ in range(steps):
            want_cnt += s.hmc_sample_ensemble_pt(eps_a, beta, l, param, 1)
        want, want_lp = s.download()
    q, lp = q0.copy(), np.empty(n)
    rng = sb.DeviceRng(seed=8)
    sb.init_logprob(f, q, lp, rng)
    eps_b = np.full(nb, 0.02)
    got_cnt = np.zeros(nb, dtype=np.uint64)
    for _ in range(steps):
        got_cnt += np.array(sb.hmc.sample_ensemble_pt(f, None, q, lp, rng, eps_b, beta, l, param), dtype=np.uint64)
    assert np.array_equal(q, want) and np.array_equal(lp, want_lp) and np.array_equal(eps_a, eps_b)
    assert np.array_equal(got_cnt, want_cnt)


def test_hmc_samples_the_tempered_gaussian_and_adapts(sb):
    nb, npb, dim = 4, 512, 6
    beta = 2.0 ** -np.arange(nb, dtype=np.float64)
    eps = np.full(nb, 0.05)
    param = sb.HmcParam.new(0.8, 2e-4)   # 512 walkers adapt a rung's epsilon in every step: keep each nudge small
    with sb.Sampler(sb.Gaussian(0.5), nb * npb, dim, seed=2) as s:
        s.init_ensemble(-1.0, 1.0)
        s.hmc_sample_ensemble_pt(eps, beta, 8, param, 600)
        s.moments_enable(nb, 1, True)
        cnt = s.hmc_sample_ensemble_pt(eps, beta, 8, sb.HmcParam.fixed(0.8), 1500)
        for r in range(nb):
            m, c, k = s.moments(r)
            sd = 1.0 / np.sqrt(beta[r])
            assert k == 1500 * npb
            assert np.all(np.abs(m) < 0.05 * sd) and np.all(np.abs(np.diag(c) * beta[r] - 1.0) < 0.05), (r, m, np.diag(c))
        rate = cnt / (1500.0 * npb)
        assert np.all(np.abs(rate - 0.8) < 0.08), rate       # the adaptation drove every rung to the target
        assert eps[0] < eps[1] < eps[2] < eps[3]             # hotter rungs take longer steps
