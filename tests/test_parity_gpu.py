"""GPU parity: the CUDA path through the C ABI against the CPU oracle and the golden vectors.

Bar (BASELINE.json north_star): fed identical partner / z / flags / u, accept decisions and
partner handling match bit for bit, positions and log-probs agree within 1e-12 relative in fp64
(positions are in fact bit-identical: the kernel rounds x1*z, x2*(1-z) and their sum separately,
as src/mcmc/utils.rs:55 does; log-probs of the transcendental-free targets are bit-identical too).
"""
import numpy as np
import pytest

from tests.util import assert_lp_close, golden_files

pytestmark = pytest.mark.gpu


def _box(rng, n, d, lo, hi):
    return np.ascontiguousarray(rng.uniform(lo, hi, (n, d)))


@pytest.mark.parametrize("variant", ["default", "reg", "seq"])
@pytest.mark.parametrize("path", golden_files(), ids=lambda p: p.split("/")[-1][:-4])
def test_golden_fixed_stream(sb, path, variant, monkeypatch):
    """default: what a caller gets (these ensembles fit one block: step_small_kernel.cuh); reg: the streaming
    register kernel forced; seq: SCORUS_CFG_SEQUENTIAL_SUM."""
    seq = variant == "seq"
    if variant == "reg":
        monkeypatch.setenv("SCORUS_KERNEL", "reg")
    g = np.load(path)
    kind, params, beta = int(g["kind"]), g["params"], g["beta"]
    n, dim = g["ens0"].shape
    nb = beta.size
    with sb.Sampler(sb.LogProb(kind, tuple(params.tolist())), n, dim, seed=0, sequential_sum=seq) as s:
        s.upload(np.ascontiguousarray(g["ens0"]), np.ascontiguousarray(g["lp0"]))
        nswap = 0
        for k in range(g["z"].shape[0]):
            if nb > 1 and k % 10 == 0:
                sw = s.swap_walkers_fixed(beta, g["jvec"][nswap], g["r"][nswap])
                assert np.array_equal(sw, g["swapped"][nswap]), f"swap decisions differ at step {k}"
                nswap += 1
            acc = s.sample_pt_fixed(beta, g["partner"][k], g["z"][k], g["flags"][k], g["u"][k])
            ens, lp = s.download()
            assert np.array_equal(acc, g["accepted"][k]), f"accept mask differs at step {k}"
            assert np.array_equal(ens, g["ens"][k]), f"positions differ at step {k}"
            assert_lp_close(kind, params, ens, lp, g["lp"][k], exact=seq)


SHAPES = [
    # kind, params, n, dim, nbeta, init box
    (0, [0.5], 64, 2, 1, (-1, 1)),
    (0, [1.0], 4096, 64, 1, (-1, 1)),
    (2, [], 2048, 64, 1, (0.9, 1.1)),
    (2, [], 96, 200, 4, (0.9, 1.1)),      # examples/test_emcee.rs: 200-D rosenbrock
    (2, [], 70, 33, 1, (0.9, 1.1)),       # odd dim, ragged tail tile
    (3, [10.0], 4096, 10, 64, (-0.5, 0.5)),
    (1, [1.5], 512, 5, 2, (-1.6, 1.6)),   # -inf proposals and -inf starts
    (4, [], 256, 2, 8, None),
    (5, [], 128, 3, 2, (-0.1, 1.1)),
    (7, None, 1024, 32, 1, (-6, 6)),
    (6, None, 256, 12, 1, (-1, 1)),
    (6, None, 2048, 100, 1, (-1, 1)),     # BASELINE config 2 shape: DMMA tile path, two chunks per lane
    (2, [], 64, 300, 2, (0.9, 1.1)),      # dim > 256: the TMA-staged cooperative kernel is the default
    (0, [0.5], 4096, 128, 1, (-1, 1)),    # two chunks per lane in the register kernel
    (3, [10.0], 1024, 200, 1, (-0.5, 0.5)),  # four chunks per lane
    (0, [0.5], 2, 1, 1, (-1, 1)),         # the smallest legal ensemble: one pair, one coordinate
    (2, [], 34, 3, 17, (0.9, 1.1)),       # two walkers per rung, a two-walker tail tile
    (3, [10.0], 66, 1, 1, (-0.5, 0.5)),   # one coordinate (odd dim, padded row), ragged tail
    (1, [1.5], 4, 2, 2, (-1.6, 1.6)),
]


def _params_for(kind, params, dim, rng):
    if kind == 7:
        m = np.zeros(dim); m[0] = 5.0
        return np.concatenate([[1.0, 2.0], m])
    if kind == 6:
        L = np.triu(rng.normal(size=(dim, dim))) * 0.3 + np.eye(dim)
        return np.concatenate([rng.normal(size=dim) * 0.1, L.ravel()])
    return np.asarray(params, dtype=np.float64)


@pytest.mark.parametrize("case", SHAPES, ids=lambda c: f"k{c[0]}_n{c[2]}_d{c[3]}_b{c[4]}")
@pytest.mark.parametrize("variant", ["default", "reg", "seq", "coop"])
@pytest.mark.parametrize("ufs", ["prob", "all", "short"])
def test_fixed_stream_vs_oracle(sb, oracle, case, ufs, variant, monkeypatch):
    """variant: the default dispatch (one-block ensembles: the shared-memory kernel; else register streaming, or
    the cooperative kernel above 256-D), the register-streaming kernel forced, the sequential-sum kernel
    (SCORUS_CFG_SEQUENTIAL_SUM), or the TMA-staged cooperative kernel forced."""
    seq = variant == "seq"
    if variant in ("coop", "reg"):
        monkeypatch.setenv("SCORUS_KERNEL", variant)
    kind, params, n, dim, nb, box = case
    rng = np.random.default_rng(1000 * kind + dim)
    params = _params_for(kind, params, dim, rng)
    if box is None:
        ens = np.ascontiguousarray(np.c_[rng.uniform(-16, 16, n), rng.uniform(-0.1, 1.1, n)])
    else:
        ens = _box(rng, n, dim, *box)
    beta = 2.0 ** -(np.arange(nb) / max(1, nb // 8))
    g = oracle.Rng(kind * 7 + dim)
    lp = oracle.init_logprob(kind, params, ens)
    with sb.Sampler(sb.LogProb(kind, tuple(params.tolist())), n, dim, seed=0, sequential_sum=seq) as s:
        s.upload(ens, None)                       # device init_logprob (K2)
        _, lp_dev = s.download()
        assert_lp_close(kind, params, ens, lp_dev, lp)
        s.upload(ens, lp)
        for step in range(6):
            ukind = oracle.UFS_ALL if ufs == "all" else oracle.UFS_PROB
            partner, z, flags, u, _ = oracle.draw_sample_streams(g, n, dim, nb, 2.0, ukind, 0.3)
            if ufs == "short" and dim > 1:
                flags = np.ascontiguousarray(flags[:, : max(1, dim // 2)])  # Func returning fewer flags (:68)
                flags[:, 0] |= (flags.sum(axis=1) == 0).astype(np.uint8)
            want_e, want_lp = ens.copy(), lp.copy()
            rc, want_acc = oracle.sample_pt_fixed(kind, params, want_e, want_lp, beta, partner, z, flags, u)
            assert rc == 0
            acc = s.sample_pt_fixed(beta, partner, z, flags, u)
            got_e, got_lp = s.download()
            assert np.array_equal(acc, want_acc), f"accept mask differs at step {step}: {np.flatnonzero(acc != want_acc)[:8]}"
            assert np.array_equal(got_e, want_e), f"positions differ at step {step}"
            assert_lp_close(kind, params, got_e, got_lp, want_lp, exact=seq)
            ens, lp = want_e, want_lp
            if nb > 1:
                jvec, r = oracle.draw_swap_streams(g, n, nb)
                rc, want_sw = oracle.swap_walkers_fixed(ens, lp, beta, jvec, r)
                assert rc == 0
                sw = s.swap_walkers_fixed(beta, jvec, r)
                got_e, got_lp = s.download()
                assert np.array_equal(sw, want_sw), f"swap decisions differ at step {step}"
                assert np.array_equal(got_e, ens), f"positions after swap differ at step {step}"
                assert_lp_close(kind, params, got_e, got_lp, lp, exact=seq)


@pytest.mark.parametrize("case", [(0, [0.5], 16, 2, 1), (2, [], 4096, 64, 1), (3, [10.0], 2048, 10, 8),
                                  (2, [], 8192, 16, 2), (0, [1.0], 320, 10, 4)],
                         ids=lambda c: f"k{c[0]}_n{c[2]}_d{c[3]}_b{c[4]}")
@pytest.mark.parametrize("seq", [False, True], ids=["reg", "seq"])
@pytest.mark.parametrize("ufs", ["prob0.5", "prob0.2", "prob0.25", "all", "onehot"])
def test_philox_mode_replays_through_oracle(sb, oracle, case, ufs, seq):
    """The device draws its own streams (Philox); the oracle regenerates the same streams on the
    CPU (oracle/philox_streams.c) and replays the step: results must agree as in fixed mode."""
    kind, params, n, dim, nb = case
    rng = np.random.default_rng(5)
    params = np.asarray(params, dtype=np.float64)
    ens = _box(rng, n, dim, 0.5, 1.5) if kind == 2 else _box(rng, n, dim, -1, 1)
    lp = oracle.init_logprob(kind, params, ens)
    beta = 2.0 ** -np.arange(nb, dtype=np.float64)
    seed = 12345 + n
    prob = float(ufs[4:]) if ufs.startswith("prob") else 0.5
    spec = (sb.UpdateFlagSpec.Prob(prob) if ufs.startswith("prob") else
            {"all": sb.UpdateFlagSpec.All(), "onehot": sb.UpdateFlagSpec.OneHotCycle()}[ufs])
    okind = oracle.UFS_PROB if ufs.startswith("prob") else {"all": oracle.UFS_ALL, "onehot": oracle.UFS_ONE_HOT_CYCLE}[ufs]
    with sb.Sampler(sb.LogProb(kind, tuple(params.tolist())), n, dim, seed=seed, sequential_sum=seq) as s:
        s.upload(ens, lp)
        for k in range(8):
            if nb > 1 and k % 4 == 0:
                _, swap_call, _ = s.counters
                jvec, r = oracle.philox_swap_streams(seed, swap_call, n, nb)
                rc, _ = oracle.swap_walkers_fixed(ens, lp, beta, jvec, r)
                assert rc == 0
                s.swap_walkers(beta)
            step, _, onehot = s.counters
            partner, z, flags, u = oracle.philox_sample_streams(seed, step, n, dim, nb, 2.0, okind, prob, onehot)
            rc, _ = oracle.sample_pt_fixed(kind, params, ens, lp, beta, partner, z, flags, u)
            assert rc == 0
            s.sample_pt(2.0, spec, beta, 1)
            got_e, got_lp = s.download()
            assert np.array_equal(got_e, ens), f"positions differ at step {k}"
            assert_lp_close(kind, params, got_e, got_lp, lp, exact=seq)
        st = s.stats()
        assert st["proposed"] == 8 * n and 0 < st["accepted"] <= st["proposed"]


def test_errors_mirror_mcmc_err(sb):
    with sb.Sampler(sb.Gaussian(), 12, 2) as s:
        s.upload(np.zeros((12, 2)) + 0.1, None)
        with pytest.raises(sb.McmcErr) as e:
            s.sample_pt(2.0, sb.UpdateFlagSpec.All(), [1.0, 0.5, 0.25, 0.125, 0.0625], 1)  # 12 % 5 != 0
        assert e.value.variant == "NWalkersMismatchesNBeta"
        with pytest.raises(sb.McmcErr) as e:
            s.sample_pt(2.0, sb.UpdateFlagSpec.All(), [1.0, 0.5, 0.25, 0.125], 1)          # 3 per rung
        assert e.value.variant == "NWalkersIsNotEven"
        with pytest.raises(sb.McmcErr) as e:
            s.swap_walkers([0.5, 1.0])
        assert e.value.variant == "BetaNotInDecrOrd"
        s.swap_walkers([1.0, 1.0])  # equal betas are tolerated (utils.rs:93)
        with pytest.raises(ValueError):
            s.sample_pt_fixed([1.0], np.arange(12), np.ones(12), np.ones((12, 2)), np.zeros(12))  # not an involution
        with pytest.raises(ValueError):
            s.sample_pt_fixed([1.0], np.arange(12) ^ 1, np.ones(12), np.zeros((12, 2)), np.zeros(12))  # nphi == 0


def test_free_functions_reference_call_shape(sb, oracle):
    """sample / sample_pt / swap_walkers / init_logprob with the reference's argument order."""
    n, dim, nb = 64, 4, 4
    rng = np.random.default_rng(0)
    ens = _box(rng, n, dim, -1, 1)
    lp = np.empty(n)
    dev = sb.DeviceRng(seed=99)
    f = sb.Gaussian(1.0)
    sb.init_logprob(f, ens, lp, dev)
    assert np.array_equal(lp, oracle.init_logprob(0, [1.0], ens))
    beta = 2.0 ** -np.arange(nb, dtype=np.float64)
    ref_e, ref_lp = ens.copy(), lp.copy()
    for k in range(5):
        smp = dev.sampler(f, n, dim)
        step, swap_call, onehot = smp.counters
        if k % 2 == 0:
            jvec, r = oracle.philox_swap_streams(99, swap_call, n, nb)
            oracle.swap_walkers_fixed(ref_e, ref_lp, beta, jvec, r)
            sb.swap_walkers(ens, lp, dev, beta)
        partner, z, flags, u = oracle.philox_sample_streams(99, step, n, dim, nb, 2.0, oracle.UFS_PROB, 0.5, onehot)
        oracle.sample_pt_fixed(0, [1.0], ref_e, ref_lp, beta, partner, z, flags, u)
        sb.sample_pt(f, ens, lp, dev, 2.0, sb.UpdateFlagSpec.Prob(0.5), beta)
        assert np.array_equal(ens, ref_e)
        assert_lp_close(0, [1.0], ens, lp, ref_lp, exact=False)
    # length mismatch: sample_pt asserts (ensemble_sample.rs:133), swap_walkers silently returns (utils.rs:88)
    with pytest.raises(AssertionError):
        sb.sample_pt(f, ens, lp[:-1].copy(), dev, 2.0, sb.UpdateFlagSpec.All(), beta)
    before = ens.copy()
    sb.swap_walkers(ens, lp[:-2].copy(), dev, beta)
    assert np.array_equal(ens, before)
    # propose_move / scale_vec / draw_z
    p1, p2 = rng.normal(size=6), rng.normal(size=6)
    fl = np.array([1, 0, 1, 1, 0, 1], dtype=np.uint8)
    want = oracle.propose_move(p1, p2, 1.37, fl)
    assert np.array_equal(sb.propose_move(p1, p2, 1.37, fl), want)
    assert np.array_equal(sb.scale_vec(p1, p2, 0.81), p1 * 0.81 + p2 * (1.0 - 0.81))
    zs = np.array([sb.draw_z(dev, 2.0) for _ in range(200)])
    assert np.all(zs >= 0.5) and np.all(zs < 2.0)


def test_cpp_host_layer_matches_python_mirror(sb, tmp_path):
    """include/scorus_b200.hpp (reference names / argument order in C++) drives the same C ABI:
    the example-shaped loops of tests/cpp/driver_like_reference.cpp must print what the Python
    mirror computes for the same seeds."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "driver")
    libdir = os.path.join(root, "scorus_b200", "lib")
    env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
    subprocess.run(["g++", "-std=c++17", "-O1", "-I", os.path.join(root, "include"),
                    os.path.join(root, "tests", "cpp", "driver_like_reference.cpp"), "-o", exe,
                    "-L", libdir, "-lscorus_b200", f"-Wl,-rpath,{libdir}"], check=True, env=env)
    niter = 30
    out = subprocess.run([exe, str(niter)], check=True, capture_output=True, text=True, env=env).stdout.splitlines()

    st = [12345]

    def lcg():
        st[0] = (st[0] * 6364136223846793005 + 1442695040888963407) % (1 << 64)
        return float(st[0] >> 11) * (1.0 / 9007199254740992.0)

    # A: test_emcee.rs shape
    ens = np.array([[0.9 + 0.2 * lcg() for _ in range(8)] for _ in range(16)])
    lp = np.empty(16)
    dev = sb.DeviceRng(seed=7)
    f = sb.Rosenbrock()
    sb.init_logprob(f, ens, lp, dev)
    for _ in range(niter):
        sb.sample(f, ens, lp, dev, 2.0, sb.UpdateFlagSpec.Prob(0.5))
    got = [ln.split() for ln in out if ln.startswith("A ")]
    assert len(got) == 16
    for i, t in enumerate(got):
        assert (float(t[2]), float(t[3]), float(t[4])) == (ens[i, 0], ens[i, 7], lp[i])
    # B: sample_rastrigin.rs shape
    blist = [2.0 ** -i for i in range(8)]
    ens = np.array([[-0.01 + 0.02 * lcg() for _ in range(20)] for _ in range(32)])
    lp = np.empty(32)
    dev = sb.DeviceRng(seed=11)
    f = sb.Rastrigin()
    sb.init_logprob(f, ens, lp, dev)
    for k in range(niter):
        if k % 10 == 0:
            sb.swap_walkers(ens, lp, dev, blist, f)
        sb.sample_pt(f, ens, lp, dev, 2.0, sb.UpdateFlagSpec.Prob(0.2), blist)
    got = [ln.split() for ln in out if ln.startswith("B ")]
    assert len(got) == 32
    for i, t in enumerate(got):
        assert (float(t[2]), float(t[3]), float(t[4])) == (ens[i, 0], ens[i, 19], lp[i])
    # T: the t-walk through mcmc::twalk::sample
    ens = np.array([[-1.0 + 2.0 * lcg() for _ in range(6)] for _ in range(16)])
    lp = np.empty(16)
    dev = sb.DeviceRng(seed=13)
    f = sb.Rosenbrock()
    sb.init_logprob(f, ens, lp, dev)
    param = sb.TWalkParams.new(6).with_pphi(0.5)
    for _ in range(niter):
        sb.twalk.sample(f, (ens, lp), param, dev, [1.0, 0.5], 4)
    got = [ln.split() for ln in out if ln.startswith("T ")]
    assert len(got) == 16
    for i, t in enumerate(got):
        assert (float(t[2]), float(t[3]), float(t[4])) == (ens[i, 0], ens[i, 5], lp[i])
    # H: ensemble-PT HMC through mcmc::hmc::naive::sample_ensemble_pt
    q0 = np.array([[0.9 + 0.2 * lcg() for _ in range(4)] for _ in range(10)])
    lp = np.empty(10)
    dev = sb.DeviceRng(seed=17)
    f = sb.Rosenbrock()
    sb.init_logprob(f, q0, lp, dev)
    eps = np.array([0.01, 0.01])
    accepted = 0
    for _ in range(niter):
        accepted += sum(sb.hmc.sample_ensemble_pt(f, None, q0, lp, dev, eps, [1.0, 0.5], 3, sb.HmcParam.quick_adj(0.7)))
    got = [ln.split() for ln in out if ln.startswith("H ")]
    assert len(got) == 10
    for i, t in enumerate(got):
        assert (float(t[2]), float(t[3]), float(t[4])) == (q0[i, 0], q0[i, 3], lp[i])
    he = [ln.split() for ln in out if ln.startswith("HE ")][0]
    assert (float(he[1]), float(he[2]), int(he[3])) == (eps[0], eps[1], accepted)
    assert "C 2" in out and "D 4" in out  # NWalkersIsNotEven, BetaNotInDecrOrd


def test_device_init_matches_stream_restatement(sb, oracle):
    """get_one_init_realization on the device (src/mcmc/init_ensemble.rs:19-32) against the CPU
    restatement of the device stream; then init_logprob of the result."""
    n, dim = 1000, 7
    lo, hi = np.linspace(-1.0, 0.5, dim), np.linspace(1.0, 2.5, dim)
    with sb.Sampler(sb.Rosenbrock(), n, dim, seed=321) as s:
        s.init_ensemble(lo, hi)
        ens, lp = s.download()
    want = oracle.philox_init_ensemble(321, n, dim, lo, hi)
    assert np.array_equal(ens, want)
    assert np.all(ens >= lo) and np.all(ens < hi)
    assert np.array_equal(lp, oracle.init_logprob(2, [], want))
    with sb.Sampler(sb.Gaussian(), 4, 2) as s:
        with pytest.raises(ValueError):
            s.init_ensemble([0.0, 1.0], [1.0, 1.0])  # empty range: Uniform::new panics in the reference


@pytest.mark.parametrize("n,dim,nb", [(256, 4, 4), (4096, 64, 1), (1000, 7, 2), (64, 128, 1)])
def test_device_mean_and_cov_match_reference_estimators(sb, oracle, n, dim, nb):
    """linear_space::utils::{mean, cov} (src/linear_space/utils.rs:4-41) per rung, on the device."""
    g = np.random.default_rng(n + dim)
    ens = np.ascontiguousarray(g.normal(size=(n, dim)) * g.uniform(0.5, 3.0, dim) + g.normal(size=dim))
    npb = n // nb
    with sb.Sampler(sb.Gaussian(), n, dim) as s:
        s.upload(ens, None)
        for r in range(nb):
            sub = np.ascontiguousarray(ens[r * npb:(r + 1) * npb])
            m, c = s.mean(r * npb, npb), s.cov(r * npb, npb)
            wm, wc = oracle.mean(sub), oracle.cov(sub)
            # different summation order than the reference's fold: 1e-12 of the sum of |terms|
            assert np.all(np.abs(m - wm) <= 1e-12 * np.abs(sub).mean(axis=0) + 1e-300)
            scale = np.sqrt(np.outer((sub ** 2).mean(axis=0), (sub ** 2).mean(axis=0)))
            assert np.all(np.abs(c - wc) <= 1e-11 * scale)
        with pytest.raises(ValueError):
            s.mean(n - 1, 2)


def test_device_trace_equals_per_step_download(sb):
    """The examples' `results[i].push(ensemble[i*n+0][0])` loop (examples/sample_high_dim.rs:84-86)."""
    nb, per, dim, steps = 4, 20, 10, 25
    g = np.random.default_rng(0)
    ens = np.ascontiguousarray(g.uniform(-1, 1, (nb * per, dim)))
    beta = 2.0 ** -np.arange(nb, dtype=np.float64)
    walkers, coords = [r * per for r in range(nb)], [0] * nb
    with sb.Sampler(sb.Gaussian(1.0), nb * per, dim, seed=5) as a, sb.Sampler(sb.Gaussian(1.0), nb * per, dim, seed=5) as b:
        a.upload(ens, None)
        b.upload(ens, None)
        a.trace_enable(walkers, coords, steps)
        want = []
        for k in range(steps):
            if k % 10 == 0:
                a.swap_walkers(beta)
                b.swap_walkers(beta)
            a.sample_pt(2.0, sb.UpdateFlagSpec.OneHotCycle(), beta, 1)
            b.sample_pt(2.0, sb.UpdateFlagSpec.OneHotCycle(), beta, 1)
            e, _ = b.download()
            want.append([e[w, 0] for w in walkers])
        got = a.trace_download(steps)
    assert got.shape == (steps, nb) and np.array_equal(got, np.array(want))


def test_periphery_mirrors_reference_call_shapes(sb, oracle):
    """init_ensemble::get_one_init_realization(y1, y2, rng) and linear_space::utils::{mean, cov}(x)."""
    dev = sb.DeviceRng(seed=77)
    lo, hi = np.array([0.9, -1.0, 0.0]), np.array([1.1, 1.0, 5.0])
    rows = np.array([sb.init_ensemble.get_one_init_realization(lo, hi, dev) for _ in range(5)])
    assert np.array_equal(rows, oracle.philox_init_ensemble(77, 5, 3, lo, hi))   # successive walkers of the stream
    x = np.ascontiguousarray(np.random.default_rng(3).normal(size=(51, 3)))
    assert np.allclose(sb.linear_space.mean(x), oracle.mean(x), rtol=1e-13, atol=1e-15)
    assert np.allclose(sb.linear_space.cov(x), oracle.cov(x), rtol=1e-12, atol=1e-15)


def test_nan_and_infinite_log_probs_follow_reference_accept_rules(sb, oracle):
    """SURVEY A.4 on the device: NaN / -inf log-probs in the accept test and in the swap."""
    kind, params = 1, [1.5]
    ens = np.array([[0.1], [1.45], [1.6], [1.7], [0.2], [1.4], [1.6], [0.0]])
    beta = [1.0, 0.5]
    lp = oracle.init_logprob(kind, params, ens)
    partner = np.array([1, 0, 3, 2, 5, 4, 7, 6], dtype=np.uint32)
    z = np.array([1.0, 2.0, 1.0, 1.0, 0.9, 1.0, 0.5, 1.0])
    u = np.array([0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1 - 2 ** -53, 0.5])
    flags = np.ones((8, 1), dtype=np.uint8)
    want_e, want_lp = ens.copy(), lp.copy()
    rc, want_acc = oracle.sample_pt_fixed(kind, params, want_e, want_lp, beta, partner, z, flags, u)
    assert rc == 0 and want_acc.tolist() == [1, 0, 0, 0, 1, 1, 1, 1]
    for seq in (False, True):
        with sb.Sampler(sb.BoundedGaussian(1.5), 8, 1, sequential_sum=seq) as s:
            s.upload(ens, None)
            acc = s.sample_pt_fixed(beta, partner, z, flags, u)
            e, l = s.download()
            assert acc.tolist() == want_acc.tolist() and np.array_equal(e, want_e)
            assert np.array_equal(l, want_lp, equal_nan=True)
            # swap: both -inf -> NaN probability -> no swap; finite vs -inf decided by the sign
            jvec, r = np.array([[0, 1, 2, 3]], dtype=np.uint32), np.array([[0.0, 0.0, 0.999, 0.5]])
            we, wl = want_e.copy(), want_lp.copy()
            rc, want_sw = oracle.swap_walkers_fixed(we, wl, beta, jvec, r)
            sw = s.swap_walkers_fixed(beta, jvec, r)
            e, l = s.download()
            assert rc == 0 and np.array_equal(sw, want_sw) and np.array_equal(e, we)
            assert np.array_equal(l, wl, equal_nan=True)


@pytest.mark.parametrize("dim", [6, 7])
def test_host_call_pinned_and_pageable_buffers_agree(sb, dim):
    """scorus_sample_pt_host writes only accepted walkers back into pinned host buffers
    (scatter download); pageable buffers take the full copy.  Same result either way."""
    n, nb = 512, 2
    g = np.random.default_rng(dim)
    ens0 = np.ascontiguousarray(g.uniform(-1, 1, (n, dim)))
    beta = [1.0, 0.5]
    out = []
    for pinned in (False, True):
        with sb.Sampler(sb.Gaussian(0.5), n, dim, seed=3) as s:
            ens = sb.pinned_empty((n, dim)) if pinned else np.empty((n, dim))
            lp = sb.pinned_empty((n,)) if pinned else np.empty(n)
            ens[:] = ens0
            s.upload(ens, None)
            s.download(ens, lp)
            for k in range(4):
                s.sample_pt_host(ens, lp, 2.0, sb.UpdateFlagSpec.Prob(0.5), beta, nsteps=1 + k % 2)
            dev_e, dev_lp = s.download()
            assert np.array_equal(ens, dev_e) and np.array_equal(lp, dev_lp)   # host copy == device copy
            out.append((ens.copy(), lp.copy()))
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
    assert not np.array_equal(out[0][0], ens0)


def test_pso_hand_off_from_ensemble(sb, oracle):
    """ParticleSwarmMaximizer::init_swarm_from_ensemble (src/opt/pso.rs:177-192): zero velocity, fitness = func(position)."""
    g = np.random.default_rng(4)
    ens = np.ascontiguousarray(g.uniform(0.5, 1.5, (20, 6)))
    swarm = sb.pso.init_swarm_from_ensemble(sb.Rosenbrock(), ens, sb.DeviceRng(seed=1))
    want = oracle.init_logprob(2, [], ens)
    assert len(swarm) == 20 and all(p.pbest is None and not p.velocity.any() for p in swarm)
    assert np.array_equal(np.array([p.position for p in swarm]), ens)
    assert np.allclose([p.fitness for p in swarm], want, rtol=1e-12)
    with sb.Sampler(sb.Rosenbrock(), 20, 6, seed=1) as s:
        s.upload(ens, None)
        s.sample(2.0, sb.UpdateFlagSpec.All(), 5)
        e, lp = s.download()
        swarm = sb.pso.swarm_from_sampler(s)
    assert np.array_equal(np.array([p.position for p in swarm]), e) and np.array_equal([p.fitness for p in swarm], lp)


def test_pso_maximizer_from_a_resident_sampler(sb, oracle):
    """ParticleSwarmMaximizer::from_ensemble (src/opt/pso.rs:108-132) on the ensemble a sampler holds: the cached
    log-probs are the first fitness; `sample` (:214-272) re-evaluates the fitness on the device (update_fitness,
    :194-212) and the swarm climbs."""
    g = np.random.default_rng(9)
    n, dim = 64, 6
    ens = np.ascontiguousarray(g.uniform(0.5, 1.5, (n, dim)))
    with sb.Sampler(sb.Rosenbrock(), n, dim, seed=2) as s:
        s.upload(ens, None)
        s.sample(2.0, sb.UpdateFlagSpec.All(), 3)
        e, lp = s.download()
        m = sb.pso.ParticleSwarmMaximizer.from_sampler(s, guess=np.ones(dim))
    assert m.particle_count == n and m.ndim == dim and np.array_equal(m.position, e) and np.array_equal(m.fitness, lp)
    assert m.gbest is not None and abs(m.gbest.fitness) < 1e-12          # Rosenbrock's maximum, the guess
    m2 = sb.pso.ParticleSwarmMaximizer.from_ensemble(sb.Rosenbrock(), e)
    assert np.allclose(m2.fitness, oracle.init_logprob(2, [], e), rtol=1e-12) and m2.gbest is None
    first = float(m2.fitness.max())
    rng = np.random.default_rng(5)
    for _ in range(30):
        m2.sample(rng, 0.6, 1.2, 1.2)
    assert np.allclose(m2.fitness, oracle.init_logprob(2, [], m2.position), rtol=1e-12)   # the device's evaluation of the moved swarm
    assert m2.gbest.fitness >= first and np.all(m2.pbest_fitness >= lp - 1e-300)


@pytest.mark.parametrize("dim", [12, 100, 130])
@pytest.mark.parametrize("shape", ["dense", "upper", "lower"])
def test_corr_gaussian_factor_shapes(sb, oracle, dim, shape):
    """The dense-target tile kernel skips the zero half of an upper-triangular factor (the k loop of column block
    n0 starts at n0); a general or lower-triangular L must take the full loop.  Same answers either way."""
    n = 128
    g = np.random.default_rng(dim)
    M = g.normal(size=(dim, dim)) * 0.2 + np.eye(dim)
    L = {"dense": M, "upper": np.triu(M), "lower": np.tril(M)}[shape]
    params = np.concatenate([g.normal(size=dim) * 0.1, L.ravel()])
    ens = np.ascontiguousarray(g.uniform(-1, 1, (n, dim)))
    lp = oracle.init_logprob(6, params, ens)
    rng = oracle.Rng(3)
    with sb.Sampler(sb.LogProb(6, tuple(params.tolist())), n, dim, seed=0) as s:
        s.upload(ens, lp)
        for step in range(3):
            partner, z, flags, u, _ = oracle.draw_sample_streams(rng, n, dim, 1, 2.0, oracle.UFS_PROB, 0.5)
            rc, want_acc = oracle.sample_pt_fixed(6, params, ens, lp, [1.0], partner, z, flags, u)
            acc = s.sample_pt_fixed([1.0], partner, z, flags, u)
            got_e, got_lp = s.download()
            assert rc == 0 and np.array_equal(acc, want_acc) and np.array_equal(got_e, ens)
            assert_lp_close(6, params, ens, got_lp, lp, exact=False)


@pytest.mark.parametrize("move", ["stretch", "twalk", "hmc"])
def test_checkpoint_resume_is_bit_identical(sb, move):
    """(download, counters, epsilon) taken mid-run and restored into a fresh context with the same seed continues the
    chain exactly: the device streams are pure functions of (seed, counters, walker)."""
    nb, npb, dim = 4, 64, 7
    n = nb * npb
    beta = 2.0 ** -np.arange(nb, dtype=np.float64)
    f = sb.Rosenbrock()
    tw, hp = sb.TWalkParams.new(dim), sb.HmcParam.quick_adj(0.7)

    def advance(s, eps, k):
        for i in range(k):
            if i % 3 == 0:
                s.swap_walkers(beta)
            if move == "stretch":
                s.sample_pt(2.0, sb.UpdateFlagSpec.OneHotCycle(), beta, 1)   # the cycling closure's state is a counter too
            elif move == "twalk":
                s.twalk_sample(tw, beta, 1)
            else:
                s.hmc_sample_ensemble_pt(eps, beta, 3, hp, 1)

    ens0 = np.ascontiguousarray(np.random.default_rng(1).uniform(0.8, 1.2, (n, dim)))
    with sb.Sampler(f, n, dim, seed=99) as a:
        a.upload(ens0, None)
        eps_a = np.full(nb, 0.004)
        advance(a, eps_a, 12)
        want = a.download()
    with sb.Sampler(f, n, dim, seed=99) as b:
        b.upload(ens0, None)
        eps_b = np.full(nb, 0.004)
        advance(b, eps_b, 6)
        ckpt = (b.download(), b.counters, eps_b.copy())
    with sb.Sampler(f, n, dim, seed=99) as c:
        c.upload(*ckpt[0])
        c.counters = ckpt[1]
        eps_c = ckpt[2].copy()
        advance(c, eps_c, 6)
        got = c.download()
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]) and np.array_equal(eps_c, eps_a)


@pytest.mark.parametrize("kind,params,n,dim,nb,ufs", [(0, [0.5], 16, 2, 1, "prob"), (2, [], 96, 20, 4, "onehot"),
                                                      (3, [10.0], 1024, 10, 8, "prob"), (6, None, 64, 12, 1, "all"),
                                                      (1, [1.5], 34, 3, 17, "prob")])
def test_one_block_ensembles_run_many_steps_per_launch(sb, kind, params, n, dim, nb, ufs, monkeypatch):
    """Ensembles that fit a thread block take the shared-memory kernel: k steps in one launch == k launches of one
    step == the same call with the chain capture hook on (one launch per step again) == the streaming register
    kernel's chain (positions bit for bit; log-probs to 1e-12: sequential against tree sums)."""
    g = np.random.default_rng(n)
    params = _params_for(kind, params, dim, g)
    f = sb.LogProb(kind, tuple(params.tolist()))
    ens0 = _box(g, n, dim, 0.5, 1.5) if kind == 2 else _box(g, n, dim, -1, 1)
    beta = 2.0 ** -np.arange(nb, dtype=np.float64)
    spec = {"prob": sb.UpdateFlagSpec.Prob(0.5), "onehot": sb.UpdateFlagSpec.OneHotCycle(), "all": sb.UpdateFlagSpec.All()}[ufs]
    steps = 23

    def run(mode):
        with sb.Sampler(f, n, dim, seed=3) as s:
            s.upload(ens0, None)
            l0 = s.launch_count
            if mode == "one_call":
                s.sample_pt(2.0, spec, beta, steps)
            elif mode == "per_step":
                for _ in range(steps):
                    s.sample_pt(2.0, spec, beta, 1)
            else:  # chain capture on: the hook runs between steps
                s.trace_enable([0], [0], steps)
                s.sample_pt(2.0, spec, beta, steps)
                assert s.trace_download(steps).shape == (steps, 1)
            return s.download(), s.launch_count - l0, s.counters, s.stats()

    (e1, l1), n1, c1, st1 = run("one_call")
    (e2, l2), n2, c2, st2 = run("per_step")
    (e3, l3), n3, c3, st3 = run("hooks")
    assert n1 == 1 and n2 == steps and n3 >= 2 * steps
    for e, l, c, st in ((e2, l2, c2, st2), (e3, l3, c3, st3)):
        assert np.array_equal(e, e1) and np.array_equal(l, l1) and c == c1 and st == st1
    assert st1["proposed"] == steps * n and 0 < st1["accepted"] < steps * n
    monkeypatch.setenv("SCORUS_KERNEL", "reg")
    (e4, l4), n4, c4, st4 = run("one_call")
    assert n4 >= steps and np.array_equal(e4, e1) and c4 == c1 and st4["accepted"] == st1["accepted"]
    assert_lp_close(kind, params, e1, l1, l4, exact=False)
