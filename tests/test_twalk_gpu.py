"""GPU parity of the ensemble / parallel-tempering t-walk (src/mcmc/twalk.rs:586-762)
against oracle/twalk.c: with identical matching / kernel choice / b / flags / walk uniforms / accept uniforms the
accept decisions and positions are bit-identical and log-probs agree within 1e-12 of their sum of |terms|."""
import numpy as np
import pytest

from tests.test_parity_gpu import _params_for
from tests.util import assert_lp_close, twalk_golden_files

pytestmark = pytest.mark.gpu

CASES = [
    # kind, params, n, dim, nbeta, init box
    (0, [0.5], 64, 2, 1, (-1, 1)),
    (0, [0.5], 2, 1, 1, (-1, 1)),           # one pair, one coordinate
    (2, [], 2048, 64, 1, (0.9, 1.1)),       # the headline shape: pphi = 1/16, four flag bits per coordinate
    (2, [], 70, 33, 1, (0.9, 1.1)),         # odd dim, ragged tail tile
    (2, [], 96, 200, 4, (0.9, 1.1)),        # four chunks per lane
    (3, [10.0], 1024, 10, 16, (-0.5, 0.5)), # rows of five double2: the two-lane x three-chunk mapping
    (1, [1.5], 512, 5, 2, (-1.6, 1.6)),     # -inf proposals and -inf starts
    (4, [], 256, 2, 8, None),
    (7, None, 512, 32, 1, (-6, 6)),
    (6, None, 256, 12, 1, (-1, 1)),         # dense target: the tile evaluation on the fp64 tensor cores
    (6, None, 512, 100, 2, (-1, 1)),
    (0, [0.5], 1024, 128, 1, (-1, 1)),
    (2, [], 34, 3, 17, (0.9, 1.1)),         # two walkers per rung
]


@pytest.mark.parametrize("path", twalk_golden_files(), ids=lambda p: p.split("/")[-1][:-4])
def test_golden_fixed_stream(sb, path):
    g = np.load(path)
    kind, params, beta = int(g["kind"]), g["params"], g["beta"]
    n, dim = g["ens0"].shape
    tw = sb.TWalkParams(g["tw"][0], g["tw"][1], g["tw"][2], (g["tw"][3], g["tw"][4]))
    with sb.Sampler(sb.LogProb(kind, tuple(params.tolist())), n, dim, seed=0) as s:
        s.upload(np.ascontiguousarray(g["ens0"]), np.ascontiguousarray(g["lp0"]))
        for k in range(g["order"].shape[0]):
            for half in (0, 1):
                acc = s.twalk_half_fixed(tw, beta, g["order"][k], half, g["kernel"][k, half], g["b"][k, half],
                                         g["flags"][k, half], g["walk_u"][k, half], g["u"][k, half])
                ens, lp = s.download()
                assert np.array_equal(acc, g["accepted"][k, half]), f"accept mask differs at step {k} half {half}"
                assert np.array_equal(ens, g["ens"][k, half]), f"positions differ at step {k} half {half}"
                assert_lp_close(kind, params, ens, lp, g["lp"][k, half], exact=False)


def _start(oracle, kind, params, n, dim, box, g):
    if box is None:      # the reference's bimodal start table, roughly: inside and outside the support
        ens = np.ascontiguousarray(np.column_stack([g.uniform(-16, 16, n), g.uniform(-0.1, 1.1, n)]))
    else:
        ens = np.ascontiguousarray(g.uniform(box[0], box[1], (n, dim)))
    return ens, oracle.init_logprob(kind, params, ens)


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"k{c[0]}_n{c[2]}_d{c[3]}_b{c[4]}")
@pytest.mark.parametrize("pphi", [None, 1.0], ids=["pphi_default", "pphi_1"])
@pytest.mark.parametrize("kernel", ["default", "tile", "direct"])
def test_half_pass_fixed_stream_vs_oracle(sb, oracle, case, pphi, kernel, monkeypatch):
    """kernel: the default choice, the shared-memory-tile kernel forced (twalk_kernel.cuh), or the register kernel
    forced wherever it exists (twalk_direct_kernel.cuh; pphi = 1 exercises its in-loop draws: more flagged coordinates
    than the per-walker z list holds)."""
    if kernel != "default":
        monkeypatch.setenv("SCORUS_TWALK_DIRECT", "0" if kernel == "tile" else "2")
    kind, params, n, dim, nb, box = case
    g = np.random.default_rng(n * 31 + dim)
    params = _params_for(kind, params, dim, g)
    beta = 2.0 ** (-np.arange(nb, dtype=np.float64) / 2.0)
    ens, lp = _start(oracle, kind, params, n, dim, box, g)
    tw_o = oracle.twalk_params(dim, pphi=pphi)
    tw = sb.TWalkParams.new(dim) if pphi is None else sb.TWalkParams.new(dim).with_pphi(pphi)
    with sb.Sampler(sb.LogProb(kind, tuple(params.tolist())), n, dim, seed=1) as s:
        s.upload(ens, lp)
        n_acc = 0
        for step in range(3):
            for half in (0, 1):
                order, kernel, b, flags, wu, u = oracle.philox_twalk_streams(99, 2 * step + half, 2 * step, n, dim, nb, tw_o, half)
                acc = s.twalk_half_fixed(tw, beta, order, half, kernel, b, flags, wu, u)
                want = oracle.twalk_half_fixed(kind, params, ens, lp, beta, tw_o, order, half, kernel, b, flags, wu, u)
                got_ens, got_lp = s.download()
                assert np.array_equal(acc, want), f"accept mask differs at step {step} half {half}"
                assert np.array_equal(got_ens, ens), f"positions differ at step {step} half {half}"
                assert_lp_close(kind, params, ens, got_lp, lp, exact=False)
                n_acc += int(acc.sum())
                assert set(np.unique(kernel)) <= {0, 1}
        if kind not in (1, 4) and pphi is None:
            assert 0 < n_acc < 3 * n


@pytest.mark.parametrize("kind,params,n,dim,nb", [(0, [0.5], 256, 6, 2), (2, [], 4096, 64, 1), (3, [10.0], 2048, 10, 8),
                                                  (2, [], 4096, 20, 2),
                                                  (3, [10.0], 8192, 10, 2),     # rungs above 1024 walkers: keyed-bijection matching
                                                  (0, [0.5], 4096, 32, 1), (0, [0.5], 2048, 130, 1)])
@pytest.mark.parametrize("kernel", ["default", "tile", "direct"])
def test_device_streams_replay_through_oracle(sb, oracle, kind, params, n, dim, nb, kernel, monkeypatch):
    """twalk::sample on the device streams == the oracle fed the CPU restatement of those streams (b taken from
    the device: the one draw that goes through pow()); with the default kernel choice and with either kernel forced."""
    if kernel != "default":
        monkeypatch.setenv("SCORUS_TWALK_DIRECT", "0" if kernel == "tile" else "2")
    g = np.random.default_rng(5)
    beta = 2.0 ** (-np.arange(nb, dtype=np.float64))
    ens, lp = _start(oracle, kind, params, n, dim, (0.5, 1.5), g)
    tw_o, tw = oracle.twalk_params(dim), sb.TWalkParams.new(dim)
    seed = 1234
    with sb.Sampler(sb.LogProb(kind, tuple(params)), n, dim, seed=seed) as s:
        s.upload(ens, lp)
        for step in range(4):
            c0 = s.counters[0]
            s.twalk_sample(tw, beta, 1)
            assert s.counters[0] == c0 + 2
            for half in (0, 1):
                order, kernel, b, flags, wu, u = oracle.philox_twalk_streams(seed, c0 + half, c0, n, dim, nb, tw_o, half)
                b_dev = s.twalk_draw_b(tw.at, c0 + half)[order[half::2]]
                trav = kernel == 1
                assert np.all(np.abs(b_dev[trav] - b[trav]) <= 4 * np.spacing(b[trav]))
                b = np.where(trav, b_dev, b)
                oracle.twalk_half_fixed(kind, params, ens, lp, beta, tw_o, order, half, kernel, b, flags, wu, u)
            got_ens, got_lp = s.download()
            assert np.array_equal(got_ens, ens), f"positions differ at step {step}"
            assert_lp_close(kind, params, ens, got_lp, lp, exact=False)
        st = s.stats()
        assert st["proposed"] == 4 * n and 0 < st["accepted"] < 4 * n
        if nb > 1:
            rs = s.rung_stats(nb)
            assert rs["accepted"].sum() == st["accepted"] and np.all(rs["proposed"] == 4 * n // nb)


def test_calc_a_edge_cases(sb, oracle):
    """calc_a (twalk.rs:480-501): no flag set or a proposed coordinate equal to the partner's -> a = 0; b = 0 in a
    Traverse move -> ln b = -inf (reject, or NaN with nphi = 2: reject); u is compared strictly."""
    n, dim = 8, 3
    ens = np.ascontiguousarray(np.arange(n * dim, dtype=np.float64).reshape(n, dim) / 10.0)
    ens[1, 2] = ens[0, 2]                      # pair (0, 1): coordinate 2 coincides and is not moved
    lp = oracle.init_logprob(0, [0.5], ens)
    order = np.arange(n, dtype=np.uint32)
    kernel = np.array([0, 0, 1, 1], dtype=np.uint8)
    b = np.array([1.0, 1.0, 0.0, 0.0])
    flags = np.array([[1, 1, 0], [0, 0, 0], [1, 1, 0], [1, 1, 1]], dtype=np.uint8)
    wu = np.full((4, dim), 0.75)
    u = np.zeros(4)                            # u = 0 accepts anything with a > 0
    tw_o, tw = oracle.twalk_params(dim), sb.TWalkParams.new(dim)
    with sb.Sampler(sb.Gaussian(), n, dim) as s:
        s.upload(ens, lp)
        acc = s.twalk_half_fixed(tw, [1.0], order, 0, kernel, b, flags, wu, u)
        want = oracle.twalk_half_fixed(0, [0.5], ens, lp, [1.0], tw_o, order, 0, kernel, b, flags, wu, u)
        got, got_lp = s.download()
    assert np.array_equal(acc, want) and not acc.any()
    assert np.array_equal(got, ens) and np.array_equal(got_lp, lp)


def test_errors_mirror_reference_panics(sb):
    tw = sb.TWalkParams.new(3)
    with sb.Sampler(sb.Gaussian(), 12, 3) as s:
        s.init_ensemble(-1.0, 1.0)
        with pytest.raises(sb.McmcErr) as e:
            s.twalk_sample(tw, [1.0, 0.5, 0.25, 0.125, 0.0625], 1)   # 12 walkers, 5 rungs: assert at twalk.rs:675
        assert e.value.code == 3
        with pytest.raises(sb.McmcErr) as e:
            s.twalk_sample(tw, [1.0, 0.5, 0.25, 0.125], 1)           # three walkers per rung: a.chunks(2) ... a[1], :610
        assert e.value.code == 2
        with pytest.raises(ValueError):
            s.twalk_sample(tw.with_pphi(0.0), [1.0], 1)
        with pytest.raises(ValueError):
            s.twalk_half_fixed(tw, [1.0], np.zeros(12, dtype=np.uint32), 0, np.zeros(6), np.ones(6), np.ones((6, 3)),
                               np.zeros((6, 3)), np.zeros(6))        # not a permutation
    with sb.Sampler(sb.Gaussian(), 8, 300) as s:
        s.init_ensemble(-1.0, 1.0)
        with pytest.raises(ValueError):
            s.twalk_sample(sb.TWalkParams.new(300), [1.0], 1)        # dim > 256
    with sb.Sampler(sb.Gaussian(), 8, 3, sequential_sum=True) as s:
        s.init_ensemble(-1.0, 1.0)
        with pytest.raises(ValueError):
            s.twalk_sample(tw, [1.0], 1)


def test_module_level_sample_matches_resident_chain(sb):
    """twalk::sample(flogprob, &mut (ensemble, logprob), param, rng, beta_list, nthreads) on host arrays, pinned
    and pageable, equals the device-resident chain with the same seed."""
    n, dim, nb, steps = 128, 5, 2, 6
    g = np.random.default_rng(3)
    ens0 = np.ascontiguousarray(g.uniform(-1, 1, (n, dim)))
    beta = [1.0, 0.5]
    f = sb.Rosenbrock()
    tw = sb.TWalkParams.new(dim)
    with sb.Sampler(f, n, dim, seed=8) as s:
        s.upload(ens0, None)
        s.twalk_sample(tw, beta, steps)
        want, want_lp = s.download()
    for pinned in (False, True):
        ens = sb.pinned_empty((n, dim)) if pinned else np.empty((n, dim))
        lp = sb.pinned_empty((n,)) if pinned else np.empty(n)
        ens[:] = ens0
        rng = sb.DeviceRng(seed=8)
        sb.init_logprob(f, ens, lp, rng)
        for _ in range(steps):
            sb.twalk.sample(f, (ens, lp), tw, rng, beta, 4)
        assert np.array_equal(ens, want) and np.array_equal(lp, want_lp)


def test_twalk_samples_the_tempered_gaussian(sb):
    """Per-rung moments accumulated on the device: N(0, 1/beta) in every coordinate for -x^2/2."""
    nb, npb, dim, steps = 4, 512, 6, 3000
    beta = 2.0 ** -np.arange(nb, dtype=np.float64)
    with sb.Sampler(sb.Gaussian(0.5), nb * npb, dim, seed=2) as s:
        s.init_ensemble(-1.0, 1.0)
        tw = sb.TWalkParams.new(dim)
        s.twalk_sample(tw, beta, 500)
        s.moments_enable(nb, 1, True)
        for k in range(steps // 10):
            s.swap_walkers(beta)
            s.twalk_sample(tw, beta, 10)
        for r in range(nb):
            m, c, cnt = s.moments(r)
            assert cnt == steps * npb
            sd = 1.0 / np.sqrt(beta[r])
            assert np.all(np.abs(m) < 0.05 * sd), (r, m)
            assert np.all(np.abs(np.diag(c) * beta[r] - 1.0) < 0.05), (r, np.diag(c) * beta[r])
            off = c - np.diag(np.diag(c))
            assert np.all(np.abs(off) * beta[r] < 0.05)
        st = s.rung_stats(nb)
        rate = st["accepted"] / st["proposed"]
        assert np.all(rate > 0.1) and np.all(rate < 0.9)
