"""The row-moving variant of the stretch step (step_reg_kernel.cuh, kRemap) on ensembles large enough that every warp
processes SEVERAL tiles -- what a step right after a ladder swap and the first step of a host-buffer call run on real
workloads, and what the small parity cases (one tile per warp) do not reach: the source rows of a warp's second and
later tiles (blocks after the first, the re-fetch of rejected rows) are found through state carried from tile to tile.

* after a swap: the chain with the rows moved by the following step (the default) against the same chain with the rows
  moved right away (SCORUS_SWAP_FUSE=0: apply_swap_kernel, then the ordinary step);
* host-buffer call on pinned buffers: the step that streams its rows from the caller's ensemble against upload +
  ordinary step + download.
* the one-sided exchange step in a world of one rank (tests/test_peer_single_gpu.py has the small cases): the partner
  stage's mbarrier phases and its refill run from tile to tile too.
A B200 keeps 148 x 16..20 warps resident: 2^18 walkers are 8192 tiles, three per warp.  (The file sorts last on purpose:
these are the largest cases of the suite.)"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES = [
    # logprob factory, n, dim, nbeta, prob, box
    (lambda sb: sb.Rosenbrock(), 1 << 18, 64, 4, 0.5, (0.9, 1.1)),       # 32 lanes per row, eight blocks per tile
    (lambda sb: sb.Rastrigin(), 1 << 18, 10, 16, 0.2, (-0.5, 0.5)),      # two lanes per row: rejected rows are fetched again
    (lambda sb: sb.Rosenbrock(), 1 << 17, 130, 2, 0.5, (0.9, 1.1)),      # 32 lanes, three chunks per lane (ITERS = 4)
    (lambda sb: sb.Gaussian(0.5), 1 << 18, 16, 256, 0.5, (-1.0, 1.0)),   # eight lanes per row; rungs of 1024: exact shuffles
]


def _env(name, value):
    if value is None:
        os.environ.pop(name, None)
    else:
        os.environ[name] = value


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"n{c[1]}_d{c[2]}_b{c[3]}")
def test_rows_moved_by_the_step_after_a_swap(sb, case):
    make, n, dim, nb, prob, box = case
    rng = np.random.default_rng(dim)
    ens = np.ascontiguousarray(rng.uniform(box[0], box[1], (n, dim)))
    beta = 2.0 ** (-np.arange(nb, dtype=np.float64) / 4.0)
    ufs = sb.UpdateFlagSpec.Prob(prob)
    saved = os.environ.get("SCORUS_SWAP_FUSE")
    try:
        with sb.Sampler(make(sb), n, dim, seed=29) as fused, sb.Sampler(make(sb), n, dim, seed=29) as plain:
            fused.upload(ens, None)
            plain.upload(ens, None)
            for k in range(3):
                _env("SCORUS_SWAP_FUSE", "1")
                fused.swap_walkers(beta)
                fused.sample_pt(2.0, ufs, beta, 2)          # the first of the two steps moves the rows
                _env("SCORUS_SWAP_FUSE", "0")
                plain.swap_walkers(beta)                     # rows moved right away
                plain.sample_pt(2.0, ufs, beta, 2)
                e1, l1 = fused.download()
                e0, l0 = plain.download()
                assert np.array_equal(e1, e0), f"positions differ after round {k}: {int(np.any(e1 != e0, axis=1).sum())} walkers"
                assert np.allclose(l1, l0, rtol=1e-12, atol=1e-9, equal_nan=True), f"log-probs differ after round {k}"
            a, b = fused.stats(), plain.stats()
            assert a["accepted"] == b["accepted"] > 0 and a["swapped"] == b["swapped"] > 0
            assert not np.array_equal(e1, ens)
    finally:
        _env("SCORUS_SWAP_FUSE", saved)


@pytest.mark.parametrize("n,dim,nb", [(1 << 18, 64, 1), (1 << 17, 128, 2)])
def test_host_call_that_streams_its_rows(sb, n, dim, nb):
    """scorus_sample_pt_host on pinned buffers with rows of 512 bytes and more: no upload, the step reads the rows from the
    caller's ensemble and writes accepted walkers straight back.  After every call the host arrays hold the chain of
    the ordinary device-resident run."""
    rng = np.random.default_rng(n % 1000 + dim)
    ens = np.ascontiguousarray(rng.uniform(0.9, 1.1, (n, dim)))
    beta = 2.0 ** -np.arange(nb, dtype=np.float64)
    ufs = sb.UpdateFlagSpec.Prob(0.5)
    saved = os.environ.get("SCORUS_HOST_STREAM")
    try:
        _env("SCORUS_HOST_STREAM", "1")
        with sb.Sampler(sb.Rosenbrock(), n, dim, seed=31) as ref, sb.Sampler(sb.Rosenbrock(), n, dim, seed=31) as host:
            ref.upload(ens, None)
            _, lp0 = ref.download()
            h_ens, h_lp = sb.pinned_empty((n, dim)), sb.pinned_empty((n,))
            h_ens[:], h_lp[:] = ens, lp0
            for k in range(3):
                ref.sample_pt(2.0, ufs, beta, 1)
                host.sample_pt_host(h_ens, h_lp, 2.0, ufs, beta)
                e0, l0 = ref.download()
                assert np.array_equal(h_ens, e0), f"positions differ after call {k}: {int(np.any(h_ens != e0, axis=1).sum())} walkers"
                assert np.allclose(h_lp, l0, rtol=1e-12, atol=1e-9), f"log-probs differ after call {k}"
            # two steps per call: the second runs on the device copy the first one wrote
            ref.sample_pt(2.0, ufs, beta, 2)
            host.sample_pt_host(h_ens, h_lp, 2.0, ufs, beta, 2)
            e0, l0 = ref.download()
            assert np.array_equal(h_ens, e0) and np.allclose(h_lp, l0, rtol=1e-12, atol=1e-9)
            e1, l1 = host.download()                        # ... and the device copy is complete
            assert np.array_equal(e1, e0) and np.allclose(l1, l0, rtol=1e-12, atol=1e-9)
            assert ref.stats()["accepted"] == host.stats()["accepted"] > 0
    finally:
        _env("SCORUS_HOST_STREAM", saved)


@pytest.mark.parametrize("make,n,dim,nb,prob", [(lambda sb: sb.Rosenbrock(), 1 << 18, 64, 1, 0.5),
                                                (lambda sb: sb.Rastrigin(), 1 << 18, 10, 4, 0.2)],
                         ids=["rosenbrock64", "rastrigin10_pt"])
def test_one_sided_exchange_step_over_several_tiles_per_warp(sb, make, n, dim, nb, prob):
    rng = np.random.default_rng(11)
    ens = np.ascontiguousarray(rng.uniform(0.4, 1.1, (n, dim)))
    beta = 2.0 ** -np.arange(nb, dtype=np.float64)
    ufs = sb.UpdateFlagSpec.Prob(prob)
    with sb.Sampler(make(sb), n, dim, seed=17) as ref, sb.Sampler(make(sb), n, dim, seed=17) as one:
        ref.upload(ens, None)
        one.upload(ens, None)
        he, hl = one.ipc_export()
        one.ipc_attach(1, 0, [he], [hl])
        for k in range(2):
            if nb > 1:
                ref.swap_walkers(beta)
                one.swap_walkers_peer(beta)
            ref.sample_pt(2.0, ufs, beta, 2)
            one.sample_pt_peer(2.0, ufs, beta, 2)
            e0, l0 = ref.download()
            e1, l1 = one.download()
            assert np.array_equal(e1, e0), f"positions differ after round {k}: {int(np.any(e1 != e0, axis=1).sum())} walkers"
            assert np.allclose(l1, l0, rtol=1e-12, atol=1e-9, equal_nan=True)
        assert ref.stats()["accepted"] == one.stats()["accepted"] > 0
