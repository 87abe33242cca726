"""Second, independent restatement of the reference hot path (numpy + math).

TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (the reference has no golden vectors and cannot
be built here): two independent restatements agreeing bit for bit -- this
file and oracle/scorus_oracle.c -- is the substitute.  Written from the Rust source, not
from the C file.  Vectorised over walkers, strictly sequential over coordinates so that
every per-walker sum has the reference's left-to-right order; exp / log / cos go through
`math` (glibc libm, what Rust's f64::exp/ln/cos call on Linux), never numpy's SIMD routines.

Reference: src/mcmc/ensemble_sample.rs:57-75,107-190; src/mcmc/utils.rs:47-70,72-112;
log-densities from examples/*.rs.
"""
from __future__ import annotations

import math

import numpy as np

INF = float("inf")


def _vmath(fn, arr):
    return np.array([fn(float(v)) for v in np.asarray(arr).ravel()], dtype=np.float64).reshape(np.shape(arr))


def logprob(kind: int, params, x: np.ndarray) -> np.ndarray:
    """x is [n, d]; returns [n]."""
    x = np.asarray(x, dtype=np.float64)
    n, d = x.shape
    p = np.asarray(params if params is not None else [], dtype=np.float64).ravel()
    if kind == 0:  # examples/test_emcee.rs:18-24 (c=.5), examples/sample_high_dim.rs:22-28 (c=1)
        c = p[0] if p.size else 0.5
        res = np.zeros(n)
        for k in range(d):
            res = res - (x[:, k] * x[:, k]) * c
        return res
    if kind == 1:  # examples/sample_bimodal.rs:35-44
        limit = p[0] if p.size else 1.5
        res = np.zeros(n)
        dead = np.zeros(n, dtype=bool)
        for k in range(d):
            res = res - x[:, k] * x[:, k]
            dead |= np.abs(x[:, k]) > limit
        res[dead] = -INF
        return res
    if kind == 2:  # examples/test_emcee.rs:26-32
        res = np.zeros(n)
        for k in range(d - 1):
            t = x[:, k + 1] - x[:, k] * x[:, k]
            s = 1.0 - x[:, k]
            res = res + (100.0 * (t * t) + s * s)
        return -res
    if kind == 3:  # examples/sample_rastrigin.rs:17-25
        a = p[0] if p.size else 10.0
        two_pi = 2.0 * math.pi
        res = np.full(n, -(float(d) * a))
        for k in range(d):
            c = _vmath(math.cos, two_pi * x[:, k])
            res = res + (a * c - x[:, k] * x[:, k])
        return res
    if kind == 4:  # examples/sample_bimodal.rs:46-54
        lo0, hi0, lo1, hi1, split, mu_a, sg_a, mu_b, sg_b = (p if p.size >= 9 else
                                                             [-15.0, 15.0, 0.0, 1.0, 0.5, -5.0, 0.1, 5.0, 1.0])
        out = np.empty(n)
        for i in range(n):
            x0, x1 = float(x[i, 0]), float(x[i, 1])
            if x0 < lo0 or x0 > hi0 or x1 < lo1 or x1 > hi1:
                out[i] = -INF
                continue
            mu, sg = (mu_a, sg_a) if x1 < split else (mu_b, sg_b)
            out[i] = -(x0 - mu) * (x0 - mu) / (2.0 * sg * sg) - math.log(sg)
        return out
    if kind == 5:  # examples/sample_bimodal.rs:56-65
        out = np.empty(n)
        for i in range(n):
            x0, x1 = float(x[i, 0]), float(x[i, 1])
            if x0 < 0.0 or x0 > 1.0 or x1 < 0.0 or x1 > 1.0:
                out[i] = -INF
            else:
                out[i] = math.log(0.1) if x0 > 0.5 else math.log(0.9)
        return out
    if kind == 6:  # builder-defined correlated Gaussian: -0.5 |L (x - mu)|^2
        mu, L = p[:d], p[d:d + d * d].reshape(d, d)
        s = np.zeros(n)
        for i in range(d):
            w = np.zeros(n)
            for j in range(d):
                w = w + L[i, j] * (x[:, j] - mu[j])
            s = s + w * w
        return -0.5 * s
    if kind == 7:  # builder-defined two-component mixture
        s1, s2, m = p[0], p[1], p[2:2 + d]
        q1 = np.zeros(n)
        q2 = np.zeros(n)
        for k in range(d):
            a_ = x[:, k] - m[k]
            b_ = x[:, k] + m[k]
            q1 = q1 + a_ * a_
            q2 = q2 + b_ * b_
        t1 = (-q1) / ((2.0 * s1) * s1) - float(d) * math.log(s1)
        t2 = (-q2) / ((2.0 * s2) * s2) - float(d) * math.log(s2)
        out = np.empty(n)
        for i in range(n):
            hi, lo = (t1[i], t2[i]) if t1[i] > t2[i] else (t2[i], t1[i])
            out[i] = -INF if hi == -INF else hi + math.log1p(math.exp(lo - hi))
        return out
    raise ValueError(kind)


def propose(ens, partner, z, flags):
    """src/mcmc/utils.rs:55 then src/mcmc/ensemble_sample.rs:67-73, for all walkers at once."""
    x1 = ens
    x2 = ens[partner]
    zc = z[:, None]
    y = x1 * zc + x2 * (1.0 - zc)          # two rounded products, one rounded sum
    fl = flags.shape[1]
    keep = np.zeros(ens.shape, dtype=bool)
    keep[:, :fl] = flags == 0               # only the first flags.len() entries can be held back
    return np.where(keep, x1, y)


def sample_pt_fixed(kind, params, ens, lp, beta, partner, z, flags, u):
    """Returns (new_ens, new_lp, accepted); inputs untouched.  Raises on the reference's asserts."""
    ens = np.asarray(ens, dtype=np.float64)
    lp = np.asarray(lp, dtype=np.float64)
    n = ens.shape[0]
    nb = len(beta)
    npb = n // nb
    assert npb * nb == n and n > 0 and npb % 2 == 0 and lp.shape[0] == n  # ensemble_sample.rs:130-133
    partner = np.asarray(partner, dtype=np.int64)
    y = propose(ens, partner, np.asarray(z), np.asarray(flags))
    ylp = logprob(kind, params, y)
    nphi = np.asarray(flags).astype(bool).sum(axis=1)
    out_e, out_lp, acc = ens.copy(), lp.copy(), np.zeros(n, dtype=np.uint8)
    for i in range(n):
        b = float(beta[i // npb])
        delta = float(ylp[i]) - float(lp[i])
        e = float(nphi[i] - 1) * math.log(float(z[i])) + delta * b
        try:
            q = math.exp(e)
        except OverflowError:
            q = INF
        if float(u[i]) < q:  # NaN compares false
            out_e[i] = y[i]
            out_lp[i] = ylp[i]
            acc[i] = 1
    return out_e, out_lp, acc


def exchange_prob(lp1, lp2, beta1, beta2):
    e = (beta2 - beta1) * (-lp2 + lp1)
    if e != e:
        return e
    try:
        x = math.exp(e)
    except OverflowError:
        x = INF
    return 1.0 if x > 1.0 else x


def swap_walkers_fixed(ens, lp, beta, jvec, r):
    """src/mcmc/utils.rs:72-112; jvec/r are [(nbeta-1), npb] in sweep order (hottest first)."""
    ens = np.array(ens, dtype=np.float64)
    lp = np.array(lp, dtype=np.float64)
    n = ens.shape[0]
    nb = len(beta)
    npb = n // nb
    assert nb * npb == n
    swapped = np.zeros(np.shape(jvec), dtype=np.uint8)
    if n != lp.shape[0]:
        return ens, lp, swapped
    for s, i in enumerate(range(nb - 1, 0, -1)):
        b1, b2 = float(beta[i]), float(beta[i - 1])
        if b1 > b2:
            raise ValueError("beta list must be in decreasing order")
        for j2 in range(npb):
            j1 = int(jvec[s][j2])
            hot, cold = i * npb + j1, (i - 1) * npb + j2
            ep = exchange_prob(float(lp[hot]), float(lp[cold]), b1, b2)
            if float(r[s][j2]) < ep:
                ens[[hot, cold]] = ens[[cold, hot]]
                lp[[hot, cold]] = lp[[cold, hot]]
                swapped[s][j2] = 1
    return ens, lp, swapped


# ------------------------------------------------------------------------------------------------
# t-walk, src/mcmc/twalk.rs (second restatement of oracle/twalk.c, written from the Rust source)
# ------------------------------------------------------------------------------------------------
def twalk_b(x: float, v: float, at: float) -> float:
    """sim_b, twalk.rs:306-322, from its two uniforms."""
    if x == 0.0:
        return x
    if v < (at - 1.0) / (2.0 * at):
        return math.pow(x, 1.0 / (at + 1.0))
    return math.pow(x, 1.0 / (1.0 - at))


def twalk_half_fixed(kind, params, ens, lp, beta, aw, order, half, kernel, b, flags, walk_u, u):
    """sample1 (twalk.rs:650-762) with supplied draws; pair k moves order[2k+half] against order[2k+1-half],
    both read inside their own rung.  Returns (new_ens, new_lp, accepted[n/2])."""
    ens = np.array(ens, dtype=np.float64)
    lp = np.array(lp, dtype=np.float64)
    n, d = ens.shape
    npb = n // len(beta)
    order = np.asarray(order).reshape(-1, 2)
    mover, partner = order[:, half], order[:, 1 - half]
    x, xp = ens[mover], ens[partner]
    f = np.asarray(flags, dtype=bool)
    kern = np.asarray(kernel)
    uu = np.asarray(walk_u, dtype=np.float64)
    # sim_walk :287-301 / sim_traverse :342-350, all proposals from the ensemble as it stands
    z = np.where(f, (aw / (1.0 + aw)) * (aw * (uu * uu) + 2.0 * uu - 1.0), 0.0)
    y_walk = x + (x - xp) * z
    bb = np.asarray(b, dtype=np.float64)[:, None]
    y_trav = np.where(f, xp + bb * (xp - x), x)
    y = np.where((kern == 0)[:, None], y_walk, y_trav)
    ylp = logprob(kind, params, y)
    acc = np.zeros(len(mover), dtype=np.uint8)
    for k in range(len(mover)):   # accept loop :736-760; a pair only touches its own mover, so order is immaterial
        nphi = int(f[k].sum())
        if nphi == 0 or np.any(y[k] == xp[k]):   # calc_a :482-483 with all_different :226-240
            a = 0.0
        else:
            t = (ylp[k] - lp[mover[k]]) * beta[mover[k] // npb]
            if kern[k] == 1:
                with np.errstate(divide="ignore", invalid="ignore"):
                    t = t + float(nphi - 2) * (math.log(b[k]) if b[k] > 0.0 else (-INF if b[k] == 0.0 else float("nan")))
            try:
                a = math.exp(t)
            except OverflowError:
                a = INF
        if u[k] < a:
            ens[mover[k]] = y[k]
            lp[mover[k]] = ylp[k]
            acc[k] = 1
    return ens, lp, acc


# ------------------------------------------------------------------------------------------------
# ensemble / parallel-tempering HMC, src/mcmc/hmc/naive.rs + utils.rs (second restatement of oracle/hmc.c)
# ------------------------------------------------------------------------------------------------
def grad_logprob(kind: int, params, x: np.ndarray) -> np.ndarray:
    """Gradient of the plugin's log-density at the rows of x[n, d] (formulas of oracle/hmc.c, the Rosenbrock one
    from the reference's example closure, examples/test_hmc.rs:27-37)."""
    x = np.asarray(x, dtype=np.float64)
    n, d = x.shape
    p = np.asarray(params if params is not None else [], dtype=np.float64).ravel()
    if kind == 0:
        c = p[0] if p.size else 0.5
        return (-2.0 * c) * x
    if kind == 1:
        return -2.0 * x
    if kind == 2:
        g = np.zeros_like(x)
        for j in range(d):
            r = np.zeros(n)
            if j >= 1:
                r = r - (200.0 * (x[:, j] - x[:, j - 1] * x[:, j - 1])) * 1.0
            if j + 1 < d:
                r = r - ((200.0 * (x[:, j + 1] - x[:, j] * x[:, j])) * (0.0 - 2.0 * x[:, j]) + 2.0 * (x[:, j] - 1.0))
            g[:, j] = r
        return g
    if kind == 3:
        a = p[0] if p.size else 10.0
        two_pi = 2.0 * 3.14159265358979323846
        k = two_pi * a
        s = k * _vmath(math.sin, two_pi * x)
        return (0.0 - s) - 2.0 * x
    if kind == 6:   # -1/2 |L (x - mu)|^2: g = 0 - L^T (L (x - mu)), both products summed left to right
        mu, L = p[:d], p[d:d + d * d].reshape(d, d)
        v = np.zeros((n, d))
        for i in range(d):
            w = np.zeros(n)
            for j in range(d):
                w = w + L[i, j] * (x[:, j] - mu[j])
            v[:, i] = w
        g = np.zeros_like(x)
        for k in range(d):
            s = np.zeros(n)
            for i in range(d):
                s = s + L[i, k] * v[:, i]
            g[:, k] = 0.0 - s
        return g
    if kind == 7:
        s1, s2, m = p[0], p[1], p[2:2 + d]
        q1, q2 = np.zeros(n), np.zeros(n)
        for i in range(d):
            a, b = x[:, i] - m[i], x[:, i] + m[i]
            q1 = q1 + a * a
            q2 = q2 + b * b
        t1 = (-q1) / ((2.0 * s1) * s1) - float(d) * math.log(s1)
        t2 = (-q2) / ((2.0 * s2) * s2) - float(d) * math.log(s2)
        lse = logprob(kind, params, x)
        r1 = _vmath(math.exp, t1 - lse) / (s1 * s1)
        r2 = _vmath(math.exp, t2 - lse) / (s2 * s2)
        return (0.0 - r1[:, None] * (x - m)) - r2[:, None] * (x + m)
    raise ValueError("plugin without a gradient")


def _kinetic(p):
    s = np.zeros(p.shape[0])
    for j in range(p.shape[1]):     # p.dot(p): left to right
        s = s + p[:, j] * p[:, j]
    return s / 2.0 * 1.0


def hmc_ensemble_pt_fixed(kind, params, q0, lp, epsilon, beta, l, target, adj, p0, u1, u2):
    """sample_ensemble_pt (naive.rs:122-292) with supplied draws.  Returns (q, lp, epsilon, accepted[n], cnt[nbeta])."""
    q0 = np.array(q0, dtype=np.float64)
    lp = np.array(lp, dtype=np.float64)
    epsilon = np.array(epsilon, dtype=np.float64)
    beta = np.asarray(beta, dtype=np.float64)
    n, d = q0.shape
    npb = n // len(beta)
    b = beta[np.arange(n) // npb][:, None]
    e = epsilon[np.arange(n) // npb][:, None]
    q, p = q0.copy(), np.array(p0, dtype=np.float64)
    k0 = _kinetic(p)
    lhq = grad_logprob(kind, params, q) * (-1.0 * b)               # :201-205 (gradient recomputed at entry, :139)
    he = e * 0.5
    for _ in range(l):                                              # leapfrog, utils.rs:17-29
        p = p - lhq * he
        q = q + (p * 1.0) * e
        lhq = grad_logprob(kind, params, q) * (0.0 - b)
        p = p - lhq * he
    pu = -logprob(kind, params, q)
    k1 = _kinetic(p)
    with np.errstate(invalid="ignore", over="ignore"):
        dh = b[:, 0] * (-lp - pu) + k0 - k1                         # :250
    acc = np.zeros(n, dtype=np.uint8)
    cnt = np.zeros(len(beta), dtype=np.uint64)
    for i in range(n):                                              # accept loop :255-290
        r = i // npb
        ok = False
        if math.isfinite(dh[i]):
            try:
                ok = u1[i] < math.exp(dh[i])
            except OverflowError:
                ok = True
        if ok:
            q0[i] = q[i]
            lp[i] = -pu[i]
            acc[i] = 1
            cnt[r] += 1
            if u2[i] < 1.0 - target:
                epsilon[r] = epsilon[r] * (1.0 + adj)
        elif u2[i] < target:
            epsilon[r] = epsilon[r] / (1.0 + adj)
    return q0, lp, epsilon, acc, cnt
