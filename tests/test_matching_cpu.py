"""Host-side checks of the matchings (no GPU): the pairing bijection's analytic inverse, and the blocked pairings of
the k-schedule as oracle/philox_streams.c restates them."""
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_inverse_bijection_on_the_host(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not found")
    exe = tmp_path / "bij_inverse"
    env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
    subprocess.run([nvcc, "-std=c++17", "-O2", "-o", str(exe), os.path.join(ROOT, "tests", "cpp", "bij_inverse.cu")],
                   check=True, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and "bad=0" in r.stdout, r.stdout


@pytest.mark.parametrize("n,nb,blocks", [(4096, 1, 4), (1 << 15, 2, 8), (48, 3, 2), (1 << 14, 1, 1)])
def test_blocked_pairing_is_a_within_block_involution(oracle, n, nb, blocks):
    """Every walker has exactly one partner, in its own block of its own rung; with one block it is the whole-rung
    matching; the sampled form returns the same partners and streams as the full form."""
    npb = n // nb
    bs = npb // blocks
    for step in (0, 7):
        partner, z, flags, u = oracle.philox_sample_streams_ex(11, step, n, 6, nb, 2.0, oracle.UFS_PROB, 0.3, 0, blocks)
        idx = np.arange(n)
        assert np.array_equal(partner[partner], idx) and np.all(partner != idx)
        assert np.array_equal(partner // bs, idx // bs)
        if blocks == 1:
            p1, z1, f1, u1 = oracle.philox_sample_streams(11, step, n, 6, nb, 2.0, oracle.UFS_PROB, 0.3, 0)
            assert np.array_equal(p1, partner) and np.array_equal(z1, z) and np.array_equal(f1, flags) and np.array_equal(u1, u)
        pick = np.random.default_rng(step).choice(n, size=min(n, 200), replace=False).astype(np.uint32)
        ps, zs, fs, us = oracle.philox_sample_streams_ex(11, step, n, 6, nb, 2.0, oracle.UFS_PROB, 0.3, 0, blocks, pick)
        assert np.array_equal(ps, partner[pick]) and np.array_equal(zs, z[pick]) and np.array_equal(fs, flags[pick]) \
            and np.array_equal(us, u[pick])
        assert np.all(flags.sum(axis=1) >= 1)


def test_flags_without_the_redraw_loop_follow_the_conditional_law(oracle):
    """Prob(p) at a p that is not dyadic: the one-pass construction (first set flag from a threshold table, the rest iid)
    must reproduce D iid Bernoulli(p) flags conditioned on at least one -- checked pattern by pattern -- and tiny p
    (which used to quantise to a threshold of 0 and spin) gives exactly one uniformly placed flag."""
    n, dim, p = 1 << 18, 4, 0.2
    _, _, flags, _ = oracle.philox_sample_streams(7, 3, n, dim, 1, 2.0, oracle.UFS_PROB, p, 0)
    pat = (flags * (1 << np.arange(dim))).sum(axis=1)
    freq = np.bincount(pat, minlength=1 << dim) / n
    want = np.array([0.0 if m == 0 else np.prod([p if (m >> d) & 1 else 1 - p for d in range(dim)]) / (1 - (1 - p) ** dim)
                     for m in range(1 << dim)])
    assert freq[0] == 0.0 and np.abs(freq - want).max() < 5.0 * np.sqrt(want.max() / n) + 1e-4
    _, _, tiny, _ = oracle.philox_sample_streams(7, 4, 1 << 16, 8, 1, 2.0, oracle.UFS_PROB, 1e-12, 0)
    assert np.all(tiny.sum(axis=1) == 1)
    where = tiny.argmax(axis=1)
    assert np.abs(np.bincount(where, minlength=8) / where.size - 0.125).max() < 0.01
