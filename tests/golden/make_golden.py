"""Generates tests/golden/*.npz: fixed-stream known-answer vectors for the hot path.

The reference ships no golden vectors for this path and cannot be built here (SURVEY.md 0.6,
8c), so these are produced by the C oracle (oracle/scorus_oracle.c) and cross-checked, before
being written, against the independent numpy restatement (oracle/numpy_oracle.py): a case is
only saved if both agree bit for bit.  Re-run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import numpy_oracle as no  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

# name, plugin kind, params, n, dim, nbeta, ufs, prob, steps, init box, seed
CASES = [
    ("c1_test_emcee_gaussian", po.LP_GAUSSIAN, [0.5], 16, 2, 1, po.UFS_PROB, 0.5, 40, (0.9, 1.1), 1),
    ("rosenbrock_d8_pt2", po.LP_ROSENBROCK, [], 24, 8, 2, po.UFS_PROB, 0.5, 12, (0.9, 1.1), 2),
    ("rastrigin_d10_pt4", po.LP_RASTRIGIN, [10.0], 32, 10, 4, po.UFS_PROB, 0.2, 12, (-0.01, 0.01), 4),
    ("high_dim_onehot_pt4", po.LP_GAUSSIAN, [1.0], 80, 10, 4, po.UFS_ONE_HOT_CYCLE, 0.0, 10, (-1.0, 1.0), 5),
    ("bimodal2d_pt8", po.LP_BIMODAL2D, [], 32, 2, 8, po.UFS_PROB, 0.5, 20, None, 6),
    ("bounded_gaussian_all", po.LP_BOUNDED_GAUSSIAN, [1.5], 20, 3, 1, po.UFS_ALL, 0.0, 12, (-1.4, 1.4), 7),
    ("mixture_d5", po.LP_MIXTURE, [1.0, 2.0, 5.0, 0.0, 0.0, 0.0, 0.0], 16, 5, 1, po.UFS_PROB, 0.5, 10, (-6.0, 6.0), 8),
    ("corr_gaussian_d7_odd", po.LP_CORR_GAUSSIAN, None, 12, 7, 1, po.UFS_ALL, 0.0, 8, (-1.0, 1.0), 9),
]

BIMODAL_TABLE = [[0.10, 0.20], [0.20, 0.10], [0.23, 0.21], [0.03, 0.22], [0.10, 0.20], [0.20, 0.24],
                 [0.20, 0.12], [0.23, 0.12]]  # examples/sample_bimodal.rs:68-101 (tiled 4x)


def ar1_factor(d, rho):
    """Upper-triangular L with L^T L = Sigma^-1 for Sigma_ij = rho^|i-j|."""
    cov = rho ** np.abs(np.subtract.outer(np.arange(d), np.arange(d)))
    prec = np.linalg.inv(cov)
    return np.linalg.cholesky(prec).T


def make(case):
    name, kind, params, n, dim, nb, ufs, prob, steps, box, seed = case
    g = po.Rng(seed)
    if kind == po.LP_CORR_GAUSSIAN:
        params = np.concatenate([np.linspace(-0.5, 0.5, dim), ar1_factor(dim, 0.9).ravel()])
    params = np.asarray(params, dtype=np.float64)
    if box is None:
        ens = np.ascontiguousarray(np.tile(np.array(BIMODAL_TABLE), (n // 8, 1)))
    else:
        ens = np.array([[box[0] + g.uniform() * (box[1] - box[0]) for _ in range(dim)] for _ in range(n)])
    beta = 2.0 ** -np.arange(nb, dtype=np.float64)
    lp = po.init_logprob(kind, params, ens)
    assert np.array_equal(lp, no.logprob(kind, params, ens), equal_nan=True), name
    out = dict(kind=kind, params=params, beta=beta, a=2.0, ens0=ens.copy(), lp0=lp.copy())
    partners, zs, flagss, us, accs, enss, lps, jvecs, rs, sws = [], [], [], [], [], [], [], [], [], []
    onehot = 0
    for k in range(steps):
        if nb > 1 and k % 10 == 0:  # examples/sample_rastrigin.rs:51-53: swap before every 10th step
            jvec, r = po.draw_swap_streams(g, n, nb)
            e2, l2, s2 = no.swap_walkers_fixed(ens, lp, beta, jvec, r)
            rc, s1 = po.swap_walkers_fixed(ens, lp, beta, jvec, r)
            assert rc == 0 and np.array_equal(s1, s2) and np.array_equal(ens, e2) and np.array_equal(lp, l2, equal_nan=True), name
            jvecs.append(jvec); rs.append(r); sws.append(s1)
        partner, z, flags, u, onehot = po.draw_sample_streams(g, n, dim, nb, 2.0, ufs, prob, onehot)
        e2, l2, a2 = no.sample_pt_fixed(kind, params, ens, lp, beta, partner, z, flags, u)
        rc, a1 = po.sample_pt_fixed(kind, params, ens, lp, beta, partner, z, flags, u)
        assert rc == 0 and np.array_equal(a1, a2) and np.array_equal(ens, e2) and np.array_equal(lp, l2, equal_nan=True), (name, k)
        partners.append(partner); zs.append(z); flagss.append(flags); us.append(u); accs.append(a1)
        enss.append(ens.copy()); lps.append(lp.copy())
    out.update(partner=np.stack(partners), z=np.stack(zs), flags=np.stack(flagss), u=np.stack(us),
               accepted=np.stack(accs), ens=np.stack(enss), lp=np.stack(lps))
    if jvecs:
        out.update(jvec=np.stack(jvecs), r=np.stack(rs), swapped=np.stack(sws))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"{name}: n={n} dim={dim} nbeta={nb} steps={steps} accept={np.mean(accs):.3f}")


# t-walk (src/mcmc/twalk.rs:586-762): name, plugin kind, params, n, dim, nbeta, pphi (None = TWalkParams::new), steps, box, seed
TWALK_CASES = [
    ("twalk_gaussian_d2", po.LP_GAUSSIAN, [0.5], 16, 2, 1, None, 20, (0.9, 1.1), 11),
    ("twalk_rosenbrock_d8_pt2", po.LP_ROSENBROCK, [], 24, 8, 2, None, 10, (0.9, 1.1), 12),
    ("twalk_rastrigin_d10_pt4", po.LP_RASTRIGIN, [10.0], 32, 10, 4, 0.3, 10, (-0.5, 0.5), 13),
    ("twalk_bimodal2d_pt2", po.LP_BIMODAL2D, [], 16, 2, 2, None, 12, None, 14),
]


def make_twalk(case):
    name, kind, params, n, dim, nb, pphi, steps, box, seed = case
    g = po.Rng(seed)
    params = np.asarray(params, dtype=np.float64)
    if box is None:
        ens = np.ascontiguousarray(np.tile(np.array(BIMODAL_TABLE), (n // 8, 1)))
    else:
        ens = np.array([[box[0] + g.uniform() * (box[1] - box[0]) for _ in range(dim)] for _ in range(n)])
    beta = 2.0 ** -np.arange(nb, dtype=np.float64)
    lp = po.init_logprob(kind, params, ens)
    tw = po.twalk_params(dim, pphi=pphi)
    out = dict(kind=kind, params=params, beta=beta, tw=np.array([tw.aw, tw.at, tw.pphi, tw.fw[0], tw.fw[1]]),
               ens0=ens.copy(), lp0=lp.copy())
    keys = ["order", "kernel", "b", "flags", "walk_u", "u", "accepted", "ens", "lp"]
    acc = {k: [] for k in keys}
    for k in range(steps):
        order, kernel, b, flags, wu, u = po.twalk_draw_streams(g, n, dim, nb, tw)
        accepted, enss, lps = [], [], []
        for half in (0, 1):
            e2, l2, a2 = no.twalk_half_fixed(kind, params, ens, lp, beta, tw.aw, order, half, kernel[half], b[half],
                                             flags[half], wu[half], u[half])
            a1 = po.twalk_half_fixed(kind, params, ens, lp, beta, tw, order, half, kernel[half], b[half], flags[half],
                                     wu[half], u[half])
            assert np.array_equal(a1, a2) and np.array_equal(ens, e2) and np.array_equal(lp, l2, equal_nan=True), (name, k, half)
            accepted.append(a1); enss.append(ens.copy()); lps.append(lp.copy())
        for key, v in zip(keys, [order, kernel, b, flags, wu, u, np.stack(accepted), np.stack(enss), np.stack(lps)]):
            acc[key].append(v)
    out.update({k: np.stack(v) for k, v in acc.items()})
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"{name}: n={n} dim={dim} nbeta={nb} steps={steps} accept={np.mean(out['accepted']):.3f} "
          f"traverse={np.mean(out['kernel']):.2f}")


# ensemble-PT HMC (src/mcmc/hmc/naive.rs:122-292): name, kind, params, n, dim, nbeta, l, epsilon, steps, box, seed
HMC_CASES = [
    ("hmc_gaussian_d3_pt2", po.LP_GAUSSIAN, [0.5], 14, 3, 2, 5, 1.6, 12, (-1.0, 1.0), 21),
    ("hmc_rosenbrock_d2", po.LP_ROSENBROCK, [], 9, 2, 1, 2, 0.06, 12, (0.8, 1.2), 22),   # examples/test_hmc.rs shape
    ("hmc_rosenbrock_d8_pt3", po.LP_ROSENBROCK, [], 24, 8, 3, 3, 0.03, 10, (0.9, 1.1), 23),
    ("hmc_bounded_gaussian_d3", po.LP_BOUNDED_GAUSSIAN, [1.5], 16, 3, 1, 6, 0.4, 10, (-1.4, 1.4), 24),
]


def make_hmc(case):
    name, kind, params, n, dim, nb, l, eps0, steps, box, seed = case
    g = po.Rng(seed)
    params = np.asarray(params, dtype=np.float64)
    q = np.array([[box[0] + g.uniform() * (box[1] - box[0]) for _ in range(dim)] for _ in range(n)])
    beta = 2.0 ** -np.arange(nb, dtype=np.float64)
    lp = po.init_logprob(kind, params, q)
    eps = np.full(nb, eps0)
    target, adj = 0.7, 0.01
    out = dict(kind=kind, params=params, beta=beta, l=l, target=target, adj=adj, q0=q.copy(), lp0=lp.copy(), eps0=eps.copy())
    keys = ["p0", "u1", "u2", "accepted", "cnt", "q", "lp", "eps"]
    acc = {k: [] for k in keys}
    nrm = np.random.default_rng(seed)      # the momenta are inputs: any N(0, 1) draws do
    for k in range(steps):
        p0 = nrm.standard_normal((n, dim))
        u1 = np.array([g.uniform() for _ in range(n)])
        u2 = np.array([g.uniform() for _ in range(n)])
        q2, l2, e2, a2, c2 = no.hmc_ensemble_pt_fixed(kind, params, q, lp, eps, beta, l, target, adj, p0, u1, u2)
        a1, c1 = po.hmc_ensemble_pt_fixed(kind, params, q, lp, eps, beta, l, target, adj, p0, u1, u2)
        assert np.array_equal(a1, a2) and np.array_equal(c1, c2) and np.array_equal(q, q2), (name, k)
        assert np.array_equal(lp, l2, equal_nan=True) and np.array_equal(eps, e2), (name, k)
        for key, v in zip(keys, [p0, u1, u2, a1, c1, q.copy(), lp.copy(), eps.copy()]):
            acc[key].append(v)
    out.update({k: np.stack(v) for k, v in acc.items()})
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"{name}: n={n} dim={dim} nbeta={nb} l={l} steps={steps} accept={np.mean(out['accepted']):.3f} eps={eps}")


if __name__ == "__main__":
    po.build()
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "stretch"):
        for c in CASES:
            make(c)
    if which in ("all", "twalk"):
        for c in TWALK_CASES:
            make_twalk(c)
    if which in ("all", "hmc"):
        for c in HMC_CASES:
            make_hmc(c)
