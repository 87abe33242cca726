"""Shared helpers for the parity tests."""
import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# plugins whose log-density uses no transcendental: bit-exact against the oracle
EXACT_KINDS = {0, 1, 2, 6}


def golden_files():
    """stretch-move vectors (the t-walk's are twalk_*.npz)"""
    return sorted(f for f in glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))
                  if not os.path.basename(f).startswith(("twalk_", "hmc_")))


def hmc_golden_files():
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, "hmc_*.npz")))


def twalk_golden_files():
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, "twalk_*.npz")))


def lp_scale(kind, params, ens):
    """Sum of |terms| of the log-density: the H5 tolerance scale (1e-12 relative to this)."""
    ens = np.asarray(ens)
    d = ens.shape[1]
    sq = (ens * ens).sum(axis=1)
    if kind == 3:
        a = params[0] if len(params) else 10.0
        return 2.0 * d * a + sq + 1.0
    if kind == 7:
        return 4.0 * (sq + (np.asarray(params[2:2 + d]) ** 2).sum()) / min(params[0], params[1]) ** 2 + 10.0 * d + 1.0
    if kind == 2:  # 100 (x[i+1] - x[i]^2)^2 + (1 - x[i])^2
        x = ens
        t = 100.0 * (np.abs(x[:, 1:]) + x[:, :-1] ** 2) ** 2 + (1.0 + np.abs(x[:, :-1])) ** 2
        return t.sum(axis=1) + 1.0
    if kind == 6:
        d = ens.shape[1]
        mu, Lm = np.asarray(params[:d]), np.abs(np.asarray(params[d:d + d * d]).reshape(d, d))
        w = (np.abs(ens) + np.abs(mu)) @ Lm.T
        return 0.5 * (w * w).sum(axis=1) * d + 1.0
    return np.abs(sq) * 2.0 + 1.0


def assert_lp_close(kind, params, ens, got, want, rtol=1e-12, exact=True):
    """exact=True (sequential-sum kernels): transcendental-free targets must be bit-identical.
    Otherwise |got - want| <= rtol * sum|terms| (a relative bound on a cancelling sum
    is only meaningful against the magnitude of its terms)."""
    got, want = np.asarray(got), np.asarray(want)
    both_inf = np.isneginf(got) & np.isneginf(want)
    both_nan = np.isnan(got) & np.isnan(want)
    skip = both_inf | both_nan
    assert np.array_equal(np.isneginf(got), np.isneginf(want)), "-inf pattern differs"
    assert np.array_equal(np.isnan(got), np.isnan(want)), "NaN pattern differs"
    if exact and kind in EXACT_KINDS:
        assert np.array_equal(got[~skip], want[~skip]), f"log-probs not bit-identical: max diff {np.max(np.abs(got[~skip]-want[~skip]))}"
    else:
        tol = rtol * lp_scale(kind, params, ens)
        with np.errstate(invalid="ignore"):
            err = np.abs(got - want)
        assert np.all(err[~skip] <= tol[~skip]), f"log-prob error {err[~skip].max()} above tolerance"
