"""T = f32 (SCORUS_CFG_F32; the reference is generic over `T: Float`, src/mcmc/ensemble_sample.rs:95): the f32
instantiation of K1 / K2 / K5 against the f32 instantiation of the oracle (oracle/liboracle_f32.so, gen_f32.py).

Bar: proposals follow the reference's rounding sequence in f32 (x1*z, x2*(1-z), their sum: three roundings), so
positions are BIT-IDENTICAL wherever the accept decision agrees; log-probs within 1e-5 * sum|terms| (the kernel sums
per-lane partials in a tree, the oracle left to right).  Accept decisions compare `u < exp(...)` on values that carry
that 1e-5: a walker whose u lies within the log-prob tolerance of its acceptance probability may legitimately fall
on either side (at f64 the same window is 1e-12 wide and never hit).  Such close calls are identified from the
oracle's own q and allowed; every other decision must be identical.  The device state is re-synchronised with the
oracle's after every step so that one close call does not cascade.
"""
import numpy as np
import pytest

from tests.util import lp_scale

pytestmark = pytest.mark.gpu

M32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10 (the device generator, rng.cuh): counters uint32 arrays, key scalars."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) for c in (c0, c1, c2, c3))
    k0, k1 = np.uint64(k0), np.uint64(k1)
    for _ in range(10):
        p0 = np.uint64(0xD2511F53) * c0
        p1 = np.uint64(0xCD9E8D57) * c2
        c0, c1, c2, c3 = ((p1 >> np.uint64(32)) ^ c1 ^ k0) & M32, p1 & M32, ((p0 >> np.uint64(32)) ^ c3 ^ k1) & M32, p0 & M32
        k0 = (k0 + np.uint64(0x9E3779B9)) & M32
        k1 = (k1 + np.uint64(0xBB67AE85)) & M32
    return c0.astype(np.uint32), c1.astype(np.uint32), c2.astype(np.uint32), c3.astype(np.uint32)


def draw(seed, step, stream, idx, sub):
    idx = np.asarray(idx, dtype=np.uint64)
    sub = np.broadcast_to(np.asarray(sub, dtype=np.uint64), idx.shape)
    step = int(step)
    c2 = np.full(idx.shape, step & 0xFFFFFFFF, dtype=np.uint64)
    c3 = np.full(idx.shape, ((step >> 32) & 0xFFFFFF) | (stream << 24), dtype=np.uint64)
    return philox4x32_10(idx, sub, c2, c3, seed & 0xFFFFFFFF, seed >> 32)


def unit_f32(hi):
    """to_unit<float> (rng.cuh): the top 24 bits of the high word."""
    return (hi >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)


def f32_zu(seed, step, n, a):
    """z and the accept uniform of every walker as the f32 kernel draws them (stream ZU = 1)."""
    _, r1, _, r3 = draw(seed, step, 1, np.arange(n), 0)
    sa = np.sqrt(np.float32(a))
    inv = np.float32(1.0) / sa
    rng_ = np.float32(2.0) * (sa - inv)
    p = unit_f32(r1) * rng_
    y = inv + p / np.float32(2.0)
    return (y * y).astype(np.float32), unit_f32(r3)


def f32_swap_r(seed, swap_call, n, nbeta):
    """swap uniforms (stream SWAP_U = 5): r[s][j2], s = nbeta-1-i, idx = cold slot j2, sub = hot rung i."""
    npb = n // nbeta
    r = np.zeros((nbeta - 1, npb), dtype=np.float32)
    for s in range(nbeta - 1):
        i = nbeta - 1 - s
        _, o1, _, _ = draw(seed, swap_call, 5, np.arange(npb), i)
        r[s] = unit_f32(o1)
    return r


def check_step(kind, params, ens0, lp0, beta, npb, z, u, flags, acc_gpu, ens_gpu, lp_gpu, acc_orc, ens_orc, lp_orc, nlp_orc):
    """One step of the device against the oracle: identical decisions except close calls; bit-identical positions and
    1e-5 * sum|terms| log-probs wherever the decision agrees."""
    n = ens0.shape[0]
    scale = lp_scale(kind, np.asarray(params, dtype=np.float64), ens_orc.astype(np.float64)).astype(np.float64)
    tol = 1e-5 * scale
    b = np.asarray(beta, dtype=np.float64)[np.arange(n) // npb]
    nphi = flags.sum(axis=1).astype(np.float64)
    with np.errstate(all="ignore"):
        logq = (nphi - 1.0) * np.log(z.astype(np.float64)) + (nlp_orc.astype(np.float64) - lp0.astype(np.float64)) * b
        close = np.abs(np.log(np.maximum(u.astype(np.float64), 1e-300)) - logq) <= 4.0 * tol * np.maximum(b, 1.0) + 1e-5
    differ = acc_gpu.astype(bool) != acc_orc.astype(bool)
    assert not np.any(differ & ~close), f"{int((differ & ~close).sum())} accept decisions differ outside the f32 window"
    assert differ.sum() <= max(3, n // 500), f"too many close calls: {int(differ.sum())} of {n}"
    same = ~differ
    assert np.array_equal(ens_gpu[same], ens_orc[same]), "positions differ"
    both_inf = np.isneginf(lp_gpu) & np.isneginf(lp_orc)
    ok = same & ~both_inf
    with np.errstate(invalid="ignore"):
        err = np.abs(lp_gpu.astype(np.float64) - lp_orc.astype(np.float64))
    assert np.array_equal(np.isneginf(lp_gpu[same]), np.isneginf(lp_orc[same]))
    assert np.all(err[ok] <= tol[ok]), f"log-prob error {err[ok].max()} above 1e-5 * sum|terms|"
    return int(differ.sum())


SHAPES = [
    # kind, params, n, dim, nbeta, box
    (0, [0.5], 64, 2, 1, (-1, 1)),
    (0, [1.0], 4096, 64, 1, (-1, 1)),          # one chunk per lane, direct mode
    (2, [], 2048, 64, 2, (0.9, 1.1)),
    (2, [], 70, 33, 1, (0.9, 1.1)),            # odd dim (rows padded to 4 floats), ragged tail tile
    (3, [10.0], 4096, 10, 8, (-0.5, 0.5)),     # c4's shape: two lanes x three chunks
    (7, None, 1024, 32, 1, (-6, 6)),           # c3's shape: tile mode (16 lanes per row)
    (1, [1.5], 512, 5, 2, (-1.6, 1.6)),        # -inf proposals
    (2, [], 96, 200, 4, (0.9, 1.1)),           # four chunks per lane
    (0, [0.5], 4096, 128, 1, (-1, 1)),         # two chunks per lane
    (4, [], 256, 2, 4, None),
]


def _params_for(kind, params, dim):
    if kind == 7:
        return [1.0, 2.0, 5.0] + [0.0] * (dim - 1)
    return list(params)


def _start(rng, kind, n, dim, box):
    if kind == 4:
        e = np.empty((n, dim))
        e[:, 0] = rng.uniform(-10, 10, n)
        e[:, 1] = rng.uniform(0.05, 0.95, n)
        return np.ascontiguousarray(e.astype(np.float32))
    return np.ascontiguousarray(rng.uniform(box[0], box[1], (n, dim)).astype(np.float32))


@pytest.fixture(scope="module")
def o32():
    from oracle import pyoracle_f32
    pyoracle_f32.build()
    return pyoracle_f32


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: f"k{s[0]}_n{s[2]}_d{s[3]}_b{s[4]}")
def test_f32_philox_steps_and_swaps_replayed_by_the_f32_oracle(sb, oracle, o32, shape):
    """The production path at T = f32 -- init_logprob, sample_pt with the device RNG, swap_walkers -- replayed through
    the f32 oracle from the regenerated streams (partner / flags / jvec are type-independent and come from
    oracle/philox_streams.c; z, u, r are rebuilt at 24 bits here)."""
    kind, params, n, dim, nb, box = shape
    params = _params_for(kind, params, dim)
    rng = np.random.default_rng(5)
    ens = _start(rng, kind, n, dim, box)
    beta = 2.0 ** -np.arange(nb, dtype=np.float64)
    npb = n // nb
    seed = 1234
    prob = 0.5 if dim > 2 else 1.0
    with sb.Sampler(sb.LogProb(kind, tuple(params)), n, dim, seed=seed, dtype=np.float32) as s:
        s.upload(ens, None)
        _, lp_dev = s.download()
        lp = o32.init_logprob(kind, params, ens)
        finite = np.isfinite(lp)
        tol0 = 1e-5 * lp_scale(kind, np.asarray(params, dtype=np.float64), ens.astype(np.float64))
        assert np.array_equal(np.isfinite(lp_dev), finite)
        assert np.all(np.abs(lp_dev[finite].astype(np.float64) - lp[finite]) <= tol0[finite]), "init_logprob (K2)"
        close_calls = 0
        for k in range(6):
            s.upload(ens, lp)                                     # the oracle's state: close calls must not cascade
            step, swap_call, onehot = s.counters
            if nb > 1 and k % 2 == 0:
                jvec, _ = oracle.philox_swap_streams(seed, swap_call, n, nb)
                r = f32_swap_r(seed, swap_call, n, nb)
                e2, l2 = ens.copy(), lp.copy()
                rc, sw = o32.swap_walkers_fixed(e2, l2, beta, jvec, r)
                assert rc == 0
                s.swap_walkers(beta)
                ge, gl = s.download()
                # a swap moves rows and log-probs, it computes nothing but the exchange probability: identical unless
                # a uniform sits within f32 rounding of it (not in these sizes)
                assert np.array_equal(ge, e2) and np.array_equal(gl, l2, equal_nan=True), "ladder swap (K5) differs"
                ens, lp = e2, l2
                step, swap_call, onehot = s.counters
            partner, _, flags, _ = oracle.philox_sample_streams(seed, step, n, dim, nb, 2.0, oracle.UFS_PROB, prob, onehot)
            z, u = f32_zu(seed, step, n, 2.0)
            e2, l2 = ens.copy(), lp.copy()
            rc, acc, nlp = o32.sample_pt_fixed(kind, params, e2, l2, beta, partner, z, flags, u)
            assert rc == 0
            s.sample_pt(2.0, sb.UpdateFlagSpec.Prob(prob), beta, 1)
            ge, gl = s.download()
            acc_gpu = np.any(ge != ens, axis=1) | ((gl != lp) & ~(np.isnan(gl) & np.isnan(lp)))
            acc_moved = np.any(e2 != ens, axis=1) | ((l2 != lp) & ~(np.isnan(l2) & np.isnan(lp)))
            close_calls += check_step(kind, params, ens, lp, beta, npb, z, u, flags, acc_gpu, ge, gl, acc_moved, e2, l2, nlp)
            ens, lp = e2, l2
        st = s.stats()
        assert st["proposed"] == 6 * n and 0 < st["accepted"] < 6 * n


def test_f32_fixed_streams_and_flag_kinds(sb, o32):
    """Caller-supplied draws (scorus_sample_pt_fixed_f32 / scorus_swap_walkers_fixed_f32) with short flag vectors, and
    the accept mask returned by the call."""
    rng = np.random.default_rng(2)
    n, dim, nb = 2048, 24, 4
    npb = n // nb
    ens = np.ascontiguousarray(rng.uniform(0.9, 1.1, (n, dim)).astype(np.float32))
    beta = 2.0 ** -np.arange(nb, dtype=np.float64)
    lp = o32.init_logprob(2, [], ens)
    with sb.Sampler(sb.Rosenbrock(), n, dim, seed=3, dtype=np.float32) as s:
        for k in range(4):
            s.upload(ens, lp)
            partner = np.empty(n, dtype=np.uint32)
            for r_ in range(nb):
                perm = rng.permutation(npb) + r_ * npb
                partner[perm[0::2]], partner[perm[1::2]] = perm[1::2], perm[0::2]
            z = ((rng.random(n) * 1.4142 + 0.7071) ** 2).astype(np.float32)
            u = rng.random(n).astype(np.float32)
            flen = dim if k % 2 == 0 else dim - 5
            flags = (rng.random((n, flen)) < 0.4).astype(np.uint8)
            flags[flags.sum(axis=1) == 0, 0] = 1
            if nb > 1 and k == 2:
                jvec = np.stack([rng.permutation(npb) for _ in range(nb - 1)]).astype(np.uint32)
                r = rng.random((nb - 1, npb)).astype(np.float32)
                e2, l2 = ens.copy(), lp.copy()
                rc, sw = o32.swap_walkers_fixed(e2, l2, beta, jvec, r)
                sw_gpu = s.swap_walkers_fixed(beta, jvec, r)
                ge, gl = s.download()
                assert rc == 0 and np.array_equal(sw_gpu, sw) and np.array_equal(ge, e2) and np.array_equal(gl, l2)
                ens, lp = e2, l2
            e2, l2 = ens.copy(), lp.copy()
            rc, acc, nlp = o32.sample_pt_fixed(2, [], e2, l2, beta, partner, z, flags, u)
            assert rc == 0
            acc_gpu = s.sample_pt_fixed(beta, partner, z, flags, u)
            ge, gl = s.download()
            check_step(2, [], ens, lp, beta, npb, z, u, flags, acc_gpu, ge, gl, acc, e2, l2, nlp)
            ens, lp = e2, l2


def test_f32_context_refuses_what_has_no_f32_instantiation(sb):
    """No silent fallback: the entry points without an f32 kernel answer INVALID_ARG on an f32 context, and f64 arrays
    are not accepted where float32 is expected."""
    n, dim = 64, 4
    ens = np.zeros((n, dim), dtype=np.float32)
    with sb.Sampler(sb.Gaussian(0.5), n, dim, seed=1, dtype=np.float32) as s:
        s.upload(ens, None)
        with pytest.raises(TypeError):
            s.upload(ens.astype(np.float64), None)
        for call in (lambda: s.twalk_sample(sb.TWalkParams.new(dim), [1.0], 1), lambda: s.mean(),
                     lambda: s.init_ensemble(0.0, 1.0), lambda: s.device_buffers()):
            with pytest.raises(Exception) as ei:
                call()
            assert "f32" in str(ei.value)
    with pytest.raises(Exception):
        sb.Sampler(sb.CorrGaussian(np.zeros(dim), np.eye(dim)), n, dim, dtype=np.float32)


def test_f32_chain_samples_the_target(sb):
    """Statistical check of the f32 path with its own RNG: 8-D unit Gaussian, 4096 walkers, variance 1 +- 3 %."""
    n, dim = 4096, 8
    rng = np.random.default_rng(0)
    ens = np.ascontiguousarray((rng.normal(size=(n, dim)) * 2.0).astype(np.float32))
    with sb.Sampler(sb.Gaussian(0.5), n, dim, seed=9, dtype=np.float32) as s:
        s.upload(ens, None)
        s.sample(2.0, sb.UpdateFlagSpec.Prob(0.5), 400)
        acc = []
        for _ in range(20):
            s.sample(2.0, sb.UpdateFlagSpec.Prob(0.5), 10)
            e, lp = s.download()
            acc.append(e.astype(np.float64))
        x = np.concatenate(acc)
    assert abs(x.mean()) < 0.03 and abs(x.var() - 1.0) < 0.03
    assert np.allclose(lp, -0.5 * (e.astype(np.float64) ** 2).sum(axis=1), rtol=1e-5, atol=1e-5)
