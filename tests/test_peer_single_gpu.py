"""The one-sided multi-GPU kernels in a world of ONE rank, so that they run in the single-GPU `pytest -m gpu` pass too:
scorus_ipc_attach(world = 1) maps nothing, every partner is "local", and the owner-computes exchange step
(scorus_sample_pt_peer: pair list by inverse bijection, flag barrier, the kPeer variants of the step kernel with the
partner rows staged by TMA), the one-sided ladder swap (scorus_swap_walkers_peer) and the host-buffer form must
reproduce the ordinary single-GPU chain bit for bit.  The real thing -- 2 and 8 ranks -- is tests/multi_gpu_check.py."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES = [
    # logprob factory, n, dim, nbeta, prob
    (lambda sb: sb.Rosenbrock(), 8192, 64, 1, 0.5),                                  # direct mode: kPeer == 2 (TMA stage)
    (lambda sb: sb.Mixture(1.0, 2.0, [5.0] + [0.0] * 31), 4096, 32, 1, 0.5),         # tile mode (16 lanes per row): kPeer == 1
    (lambda sb: sb.Rastrigin(), 4096 * 3, 10, 3, 0.2),                               # tempered, 2 lanes x 3 chunks, loop-free flags
    (lambda sb: sb.Gaussian(0.5), 70, 7, 1, 0.5),                                    # small rung: exact shuffle + scan of the matching
    (lambda sb: sb.Rosenbrock(), 2048, 200, 2, 0.5),                                 # four chunks per lane
]


def _attach_world_of_one(s):
    he, hl = s.ipc_export()
    s.ipc_attach(1, 0, [he], [hl])


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"n{c[1]}_d{c[2]}_b{c[3]}")
def test_exchange_step_and_swap_in_a_world_of_one(sb, case):
    make, n, dim, nb, prob = case
    rng = np.random.default_rng(3)
    ens = np.ascontiguousarray(rng.uniform(0.4, 1.1, (n, dim)))
    beta = 2.0 ** -np.arange(nb, dtype=np.float64)
    ufs = sb.UpdateFlagSpec.Prob(prob)
    with sb.Sampler(make(sb), n, dim, seed=17) as ref, sb.Sampler(make(sb), n, dim, seed=17) as one:
        ref.upload(ens, None)
        one.upload(ens, None)
        _attach_world_of_one(one)
        for k in range(4):
            if nb > 1:
                ref.swap_walkers(beta)
                one.swap_walkers_peer(beta)
            ref.sample_pt(2.0, ufs, beta, 2)
            one.sample_pt_peer(2.0, ufs, beta, 2)
            e0, l0 = ref.download()
            e1, l1 = one.download()
            assert np.array_equal(e1, e0), f"positions differ after round {k}"
            # (one-block ensembles run the sequential-sum kernel on the ordinary path: log-probs to the last ulps)
            assert np.allclose(l1, l0, rtol=1e-12, atol=1e-9, equal_nan=True), f"log-probs differ after round {k}"
        a, b = ref.stats(), one.stats()
        assert a["accepted"] == b["accepted"] and a["proposed"] == b["proposed"] and a["swapped"] == b["swapped"]
        assert one.counters == ref.counters


def _cycling_closure(dim, flen):
    """A stateful Func closure in the spirit of examples/sample_high_dim.rs:59-65: two flags set, advancing by one
    coordinate per call."""
    state = {"i": 0}

    def f():
        v = np.zeros(flen, dtype=bool)
        v[state["i"] % flen] = True
        v[(state["i"] * 3 + 1) % flen] = True
        state["i"] += 1
        return v
    return f


@pytest.mark.parametrize("n,dim,flen", [(4096, 64, 64), (70, 7, 5)])
def test_func_flags_over_the_one_sided_step(sb, n, dim, flen):
    """UpdateFlagSpec::Func (ensemble_sample.rs:31,52,150-153) over the whole-ensemble one-sided step: the mask of the
    whole ensemble, indexed by global walker id, gives the chain the ordinary step gives with the same closure."""
    rng = np.random.default_rng(8)
    ens = np.ascontiguousarray(rng.uniform(0.8, 1.2, (n, dim)))
    with sb.Sampler(sb.Rosenbrock(), n, dim, seed=23) as ref, sb.Sampler(sb.Rosenbrock(), n, dim, seed=23) as one:
        ref.upload(ens, None)
        one.upload(ens, None)
        _attach_world_of_one(one)
        fa, fb = _cycling_closure(dim, flen), _cycling_closure(dim, flen)
        for k in range(3):
            ref.sample_pt(2.0, sb.UpdateFlagSpec.Func(fa), [1.0], 1)
            one.sample_pt_peer(2.0, sb.UpdateFlagSpec.Mask(sb.UpdateFlagSpec.Func(fb).evaluate(n)), [1.0], 1)
            e0, l0 = ref.download()
            e1, l1 = one.download()
            assert np.array_equal(e1, e0), f"positions differ after step {k}"
            assert np.allclose(l1, l0, rtol=1e-12, atol=1e-9)
        assert ref.stats()["accepted"] == one.stats()["accepted"] > 0
        with pytest.raises(Exception):   # a mask covers exactly one step
            one.sample_pt_peer(2.0, sb.UpdateFlagSpec.Mask(np.ones((n, flen), dtype=np.uint8)), [1.0], 2)
        with pytest.raises(Exception):   # NPhiZero, ensemble_sample.rs:171-173
            one.sample_pt_peer(2.0, sb.UpdateFlagSpec.Mask(np.zeros((n, flen), dtype=np.uint8)), [1.0], 1)


def test_host_buffer_form_of_the_exchange_step(sb):
    """scorus_sample_pt_peer_host on pinned buffers: after every call the host arrays hold the chain (only the accepted
    walkers travel back -- marked on their home rank by the owner of the pair)."""
    n, dim = 4096, 64
    rng = np.random.default_rng(5)
    ens = np.ascontiguousarray(rng.uniform(0.9, 1.1, (n, dim)))
    ufs = sb.UpdateFlagSpec.Prob(0.5)
    with sb.Sampler(sb.Rosenbrock(), n, dim, seed=4) as ref, sb.Sampler(sb.Rosenbrock(), n, dim, seed=4) as one:
        ref.upload(ens, None)
        _, lp0 = ref.download()
        _attach_world_of_one(one)
        h_ens, h_lp = sb.pinned_empty((n, dim)), sb.pinned_empty((n,))
        h_ens[:], h_lp[:] = ens, lp0
        for k in range(3):
            ref.sample_pt(2.0, ufs, [1.0], 1)
            one.sample_pt_peer_host(h_ens, h_lp, 2.0, ufs, [1.0])
            e0, l0 = ref.download()
            assert np.array_equal(h_ens, e0) and np.allclose(h_lp, l0, rtol=1e-12, atol=1e-9)


def test_peer_entry_points_need_an_attached_world(sb):
    n, dim = 64, 4
    with sb.Sampler(sb.Gaussian(0.5), n, dim, seed=1) as s:
        s.upload(np.zeros((n, dim)), None)
        for call in (lambda: s.sample_pt_peer(2.0, sb.UpdateFlagSpec.All(), [1.0], 1), lambda: s.peer_barrier(),
                     lambda: s.swap_walkers_peer([1.0, 0.5])):
            with pytest.raises(Exception) as ei:
                call()
            assert "ipc_attach" in str(ei.value)
