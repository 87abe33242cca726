"""The exp-free decisions of the kernels, restated in numpy and checked against the literal comparisons they replace.

* ladder swap (scorus_b200/csrc/swap_kernel.cuh): the reference decides `r < min(1, exp(x))`,
  x = (beta_cold - beta_hot) (lp_carry - lp_cold) (src/mcmc/utils.rs:63-69, 105).  swap_paths_kernel turns that into a
  threshold on the carried log-prob, mid = lp_cold + ln r / db with a band w; swap_walk_kernel answers from
  `lp_carry > mid + w` / `lp_carry < mid - w` and evaluates the literal expression only inside the band.  The claim:
  whenever the shortcut answers, it answers what the literal expression answers -- with ln r taken in f32 and 1/db
  instead of a division, each perturbed here by a few ulps more than the device functions can be off.
* stretch move / t-walk accept (step_reg_kernel.cuh, twalk_direct_kernel.cuh): `u < exp(t)`
  (src/mcmc/ensemble_sample.rs:184-185) answered from `ln u < t -+ d`, d = 1e-5 (1 + |t|), outside the band.
* Rastrigin's cos(2 pi x) (logprob.cuh: cos2pi): the polynomial against mpmath."""
import math

import numpy as np
import pytest

GUARD = 1e-5          # swap_guard<double>()
ROUND_GUARD = 1e-12   # swap_round_guard<double>()


def _paths(r, lp_cold, db, lr_ulps=0, inv_ulps=0):
    """mid, w as swap_paths_kernel computes them (float64 arithmetic, ln r through f32), with the f32 logarithm moved by
    lr_ulps and 1/db by inv_ulps units in the last place."""
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        lr32 = np.log(r.astype(np.float32)).astype(np.float32)
        for _ in range(abs(lr_ulps)):   # (an infinite logarithm -- r below f32's range -- stays infinite)
            lr32 = np.where(np.isfinite(lr32), np.nextafter(lr32, np.float32(np.inf if lr_ulps > 0 else -np.inf)), lr32)
        L = np.where((r > 0.0) & (r < 1.0), lr32.astype(np.float64), np.nan)
        inv_db = np.where(db > 0.0, 1.0 / db, np.nan)
        for _ in range(abs(inv_ulps)):
            inv_db = np.nextafter(inv_db, np.inf if inv_ulps > 0 else -np.inf)
        q = L * inv_db
        w = (GUARD * (1.0 + np.abs(L))) * inv_db + ROUND_GUARD * (np.abs(lp_cold) + np.abs(q))
        mid = lp_cold + q
        w32 = (w.astype(np.float32) * np.float32(1.001)).astype(np.float32)
    return mid, w32.astype(np.float64)


def _literal_swap(r, lp_cold, lp_carry, db):
    with np.errstate(over="ignore", invalid="ignore"):
        x = db * (-lp_cold + lp_carry)
        ex = np.exp(x)
        ep = np.where(ex > 1.0, 1.0, ex)
        return r < ep          # NaN -> False, as in Rust


@pytest.mark.parametrize("lr_ulps,inv_ulps", [(0, 0), (3, 2), (-3, -2), (3, -2), (-3, 2)])
def test_swap_threshold_answers_what_the_literal_expression_answers(lr_ulps, inv_ulps):
    g = np.random.default_rng(5)
    n = 400_000
    r = g.random(n)
    r[:1000] = np.nextafter(0.0, 1.0) * g.integers(1, 1 << 20, 1000)       # tiny uniforms
    r[1000:2000] = 1.0 - g.random(1000) * 1e-9                             # just below one
    db = np.concatenate([g.uniform(1e-6, 1.0, n // 2), 10.0 ** g.uniform(-9, 2, n - n // 2)])
    lp_cold = np.concatenate([g.normal(0, 30, n // 2), -10.0 ** g.uniform(-3, 14, n - n // 2)])
    # carried log-probs all over, and adversarially close to the threshold: x = ln r (1 + delta)
    with np.errstate(divide="ignore"):
        L = np.log(r)
    delta = np.concatenate([np.zeros(n // 4), 10.0 ** g.uniform(-9, -1, n - n // 4) * g.choice([-1.0, 1.0], n - n // 4)])
    lp_carry = np.where(g.random(n) < 0.5, lp_cold + L * (1.0 + delta) / db, lp_cold + g.normal(0, 3, n) / db)
    mid, w = _paths(r, lp_cold, db, lr_ulps, inv_ulps)
    with np.errstate(invalid="ignore"):
        up, dn = lp_carry > mid + w, lp_carry < mid - w
    lit = _literal_swap(r, lp_cold, lp_carry, db)
    assert not np.any(up & dn)
    assert np.array_equal(lit[up], np.ones(int(up.sum()), dtype=bool)), "shortcut said swap, the literal expression does not"
    assert not np.any(lit[dn]), "shortcut said no swap, the literal expression swaps"
    sure = up | dn
    assert sure.mean() > 0.6          # the shortcut does answer (the adversarial half of the cases sits in the band)
    # away from the threshold and at ordinary magnitudes (the rounding part of the band grows with |lp_cold|) it always does
    far = (np.abs(delta) > 1e-3) & (np.abs(lp_cold) < 1e3) & np.isfinite(mid)
    assert sure[far].mean() > 0.99


def test_swap_threshold_sends_everything_odd_to_the_literal_expression():
    r = np.array([0.0, 1.0, 1.5, -0.1, np.nan, 0.5, 0.5, 0.5, 0.5, 0.5, 0.5])
    lp_cold = np.array([0.0, 0.0, 0.0, 0.0, 0.0, np.nan, -np.inf, np.inf, 0.0, 0.0, 1e300])
    db = np.array([0.1, 0.1, 0.1, 0.1, 0.1, 0.1, 0.1, 0.1, 0.0, -0.1, 1e300])
    mid, w = _paths(r, lp_cold, db)
    for lp_carry in (-np.inf, -1e308, -1.0, 0.0, 1.0, 1e308, np.inf, np.nan):
        with np.errstate(invalid="ignore"):
            up, dn = lp_carry > mid + w, lp_carry < mid - w
        lit = _literal_swap(r, lp_cold, np.full_like(r, lp_carry), db)
        assert np.array_equal(lit[up], np.ones(int(up.sum()), dtype=bool)) and not np.any(lit[dn])
    # r outside (0, 1), db <= 0 and a NaN log-prob never take the shortcut
    assert np.all(np.isnan(mid[[0, 1, 2, 3, 4, 5, 8, 9]]))


def test_accept_by_logarithms_answers_what_u_less_than_exp_t_answers():
    g = np.random.default_rng(6)
    n = 500_000
    u = g.random(n)
    u[:500] = np.nextafter(0.0, 1.0) * g.integers(1, 1 << 30, 500)
    with np.errstate(divide="ignore"):
        lu = np.where(u > 0.0, np.log(u), np.nan)
    delta = 10.0 ** g.uniform(-12, -1, n) * g.choice([-1.0, 1.0], n)
    t = np.where(g.random(n) < 0.6, lu * (1.0 + delta), g.normal(0, 20, n))
    t[-6:] = [np.inf, -np.inf, np.nan, 800.0, -800.0, 0.0]
    for ulps in (0, 2, -2):         # the device logarithm may differ from this one by an ulp or two
        l2 = lu.copy()
        for _ in range(abs(ulps)):
            l2 = np.nextafter(l2, np.inf if ulps > 0 else -np.inf)
        with np.errstate(invalid="ignore", over="ignore"):
            d = GUARD * (1.0 + np.abs(t))
            yes, no = l2 < t - d, l2 > t + d
            lit = u < np.exp(t)
        assert not np.any(yes & no)
        assert np.all(lit[yes]) and not np.any(lit[no])
        assert (yes | no).mean() > 0.5
    # u = 0 and non-finite t always go to the literal comparison
    with np.errstate(invalid="ignore"):
        d0 = GUARD * (1.0 + np.abs(t[-6:-3]))
        assert not np.any((lu[-6:-3] < t[-6:-3] - d0) | (lu[-6:-3] > t[-6:-3] + d0))


def _cos2pi(y):
    """logprob.cuh: LpRastrigin::cos2pi, operation by operation (math.fma where Python has it, else the two roundings:
    the bound below holds for both)."""
    c = [float.fromhex(h) for h in (
        "-0x1.3bd3cc9be45dep+4", "0x1.03c1f081b5ac4p+6", "-0x1.55d3c7e3cbffap+6", "0x1.e1f506891babbp+5", "-0x1.a6d1f2a204a8cp+4",
        "0x1.f9d38a3763cc3p+2", "-0x1.b6e24f44b128fp+0", "0x1.20c62c2f2d7f5p-2", "-0x1.2a0c591af8314p-5", "0x1.ef6e308d6d1c4p-9",
        "-0x1.52ae4120fde27p-12")]
    fma = getattr(math, "fma", lambda a, b, cc: a * b + cc)
    r = y - float(np.rint(y))
    ar = abs(r)
    flip = ar > 0.25
    b = 0.5 - ar if flip else ar
    s = b * b
    p = c[10]
    for k in range(9, -1, -1):
        p = fma(p, s, c[k])
    p = fma(p, s, 1.0)
    return -p if flip else p


def test_cos2pi_polynomial_against_mpmath():
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 40
    g = np.random.default_rng(7)
    ys = np.concatenate([g.uniform(-6.0, 6.0, 4000), g.uniform(-1e-3, 1e-3, 200), np.array([0.0, 0.25, 0.5, 0.75, 1.0, -0.25, 5.12, -5.12, 1e6 + 0.125])])
    worst = 0.0
    for y in ys:
        ref = mp.cos(2 * mp.pi * mp.mpf(float(y)))
        worst = max(worst, abs(float(mp.mpf(_cos2pi(float(y))) - ref)))
    assert worst < 5e-16, worst
    # and the reference's own expression, cos(fl(2 pi y)), is further from the true value than that (argument rounding)
    worst_ref = max(abs(float(mp.mpf(math.cos(2.0 * math.pi * float(y))) - mp.cos(2 * mp.pi * mp.mpf(float(y))))) for y in ys[:2000])
    assert worst_ref > worst
    for y in (float("inf"), float("nan")):
        assert math.isnan(_cos2pi(y))
