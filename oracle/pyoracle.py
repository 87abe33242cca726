"""ctypes binding of the C oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY -- see oracle/scorus_oracle.h.  Importable from tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs; never from
scorus_b200/ (the product).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

# plugin ids / ufs kinds / status codes (same values as include/scorus_b200.h)
LP_GAUSSIAN, LP_BOUNDED_GAUSSIAN, LP_ROSENBROCK, LP_RASTRIGIN = 0, 1, 2, 3
LP_BIMODAL2D, LP_STEP2D, LP_CORR_GAUSSIAN, LP_MIXTURE = 4, 5, 6, 7
UFS_PROB, UFS_ALL, UFS_ONE_HOT_CYCLE = 0, 1, 2
OK = 0
ERR_NWALKERS_IS_ZERO, ERR_NWALKERS_IS_NOT_EVEN = 1, 2
ERR_NWALKERS_MISMATCHES_NBETA, ERR_BETA_NOT_IN_DECR_ORD = 3, 4
ERR_LENGTH_MISMATCH, ERR_INVALID_ARG, ERR_NPHI_ZERO = 5, 6, 7


def build(force: bool = False) -> str:
    """Compile oracle/liboracle.so with the committed Makefile (gcc, no FMA contraction)."""
    srcs = ["scorus_oracle.c", "philox_streams.c", "ref_structure_bench.c", "chain_stats.c", "twalk.c", "hmc.c", "scorus_oracle.h", "Makefile"]
    stale = force or not os.path.exists(_LIB_PATH) or any(
        os.path.getmtime(os.path.join(_HERE, s)) > os.path.getmtime(_LIB_PATH) for s in srcs)
    if stale:
        env = dict(os.environ)
        env.pop("CC", None)
        subprocess.run(["make", "-C", _HERE, "-B", "liboracle.so"], check=True, env=env,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return _LIB_PATH


class TwalkParams(C.Structure):
    """TWalkParams, src/mcmc/twalk.rs:75-83 (fw cumulative)."""
    _fields_ = [("aw", C.c_double), ("at", C.c_double), ("pphi", C.c_double), ("fw", C.c_double * 2)]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        dp, u32p, u8p, u64p = (C.POINTER(C.c_double), C.POINTER(C.c_uint32), C.POINTER(C.c_uint8),
                               C.POINTER(C.c_uint64))
        sz, u64, dbl, ci = C.c_size_t, C.c_uint64, C.c_double, C.c_int
        L.orc_logprob.restype = dbl
        L.orc_logprob.argtypes = [ci, dp, sz, dp, sz]
        L.orc_z_from_p.restype = dbl
        L.orc_z_from_p.argtypes = [dbl, dbl]
        L.orc_z_range.restype = dbl
        L.orc_z_range.argtypes = [dbl]
        L.orc_exchange_prob.restype = dbl
        L.orc_exchange_prob.argtypes = [dbl] * 4
        L.orc_propose_move.restype = None
        L.orc_propose_move.argtypes = [dp, dp, dbl, u8p, sz, sz, dp]
        L.orc_init_logprob.restype = None
        L.orc_init_logprob.argtypes = [ci, dp, sz, dp, sz, sz, dp]
        L.orc_sample_pt_fixed.restype = ci
        L.orc_sample_pt_fixed.argtypes = [ci, dp, sz, dp, dp, sz, sz, sz, dp, sz, u32p, dp, u8p, sz,
                                          dp, u8p, dp, dp]
        L.orc_swap_walkers_fixed.restype = ci
        L.orc_swap_walkers_fixed.argtypes = [dp, dp, sz, sz, sz, dp, sz, u32p, dp, u8p]
        L.orc_rng_seed.restype = None
        L.orc_rng_seed.argtypes = [C.c_void_p, u64]
        L.orc_rng_uniform.restype = dbl
        L.orc_rng_uniform.argtypes = [C.c_void_p]
        L.orc_draw_sample_streams.restype = ci
        L.orc_draw_sample_streams.argtypes = [C.c_void_p, sz, sz, sz, dbl, ci, dbl, u64p, u32p, dp,
                                              u8p, dp]
        L.orc_draw_swap_streams.restype = ci
        L.orc_draw_swap_streams.argtypes = [C.c_void_p, sz, sz, u32p, dp]
        L.orc_sample_pt.restype = ci
        L.orc_sample_pt.argtypes = [ci, dp, sz, dp, dp, sz, sz, C.c_void_p, dbl, ci, dbl, u64p, dp, sz]
        L.orc_swap_walkers.restype = ci
        L.orc_swap_walkers.argtypes = [dp, dp, sz, sz, C.c_void_p, dp, sz]
        L.orc_mean.restype = None
        L.orc_mean.argtypes = [dp, sz, sz, dp]
        L.orc_cov.restype = None
        L.orc_cov.argtypes = [dp, sz, sz, dp]
        L.orc_init_realization.restype = None
        L.orc_init_realization.argtypes = [C.c_void_p, dp, dp, sz, dp]
        L.orc_philox4x32_10.restype = None
        L.orc_philox4x32_10.argtypes = [u32p, u32p, u32p]
        L.orc_philox_sample_streams.restype = ci
        L.orc_philox_sample_streams.argtypes = [u64, u64, sz, sz, sz, dbl, ci, dbl, u64, u32p, dp, u8p, dp]
        L.orc_philox_sample_streams_ex.restype = ci
        L.orc_philox_sample_streams_ex.argtypes = [u64, u64, sz, sz, sz, dbl, ci, dbl, u64, sz, u32p, sz, u32p, dp, u8p, dp]
        L.orc_philox_swap_streams.restype = ci
        L.orc_philox_swap_streams.argtypes = [u64, u64, sz, sz, u32p, dp]
        L.orc_philox_init_ensemble.restype = None
        L.orc_philox_init_ensemble.argtypes = [u64, sz, sz, u64, dp, dp, dp]
        L.orc_ref_structure_run.restype = dbl
        L.orc_ref_structure_run.argtypes = [ci, dp, sz, dp, dp, sz, sz, dp, sz, dbl, ci, dbl, u64, sz,
                                            ci, dp]
        L.orc_flat_threaded_run.restype = dbl
        L.orc_flat_threaded_run.argtypes = [ci, dp, sz, dp, dp, sz, sz, dp, sz, dbl, ci, dbl, u64, u64,
                                            sz, ci]
        L.orc_max_threads.restype = ci
        L.orc_max_threads.argtypes = []
        L.orc_twalk_params_default.restype = None
        L.orc_twalk_params_default.argtypes = [sz, C.POINTER(TwalkParams)]
        L.orc_twalk_b_from_uniforms.restype = dbl
        L.orc_twalk_b_from_uniforms.argtypes = [dbl, dbl, dbl]
        L.orc_twalk_half_fixed.argtypes = [ci, dp, sz, dp, dp, sz, sz, dp, sz, C.POINTER(TwalkParams), u32p, ci, u8p, dp,
                                           u8p, dp, dp, u8p, dp, dp]
        L.orc_twalk_half_fixed_literal.argtypes = L.orc_twalk_half_fixed.argtypes
        L.orc_twalk_draw_streams.argtypes = [C.c_void_p, sz, sz, sz, C.POINTER(TwalkParams), u32p, u8p, dp, u8p, dp, dp]
        L.orc_twalk_sample.argtypes = [ci, dp, sz, dp, dp, sz, sz, C.c_void_p, C.POINTER(TwalkParams), dp, sz]
        L.orc_philox_twalk_streams.argtypes = [u64, u64, u64, sz, sz, sz, C.POINTER(TwalkParams), ci, u32p, u8p, dp,
                                               u8p, dp, dp]
        L.orc_grad_logprob.argtypes = [ci, dp, sz, dp, sz, dp]
        L.orc_hmc_ensemble_pt_fixed.argtypes = [ci, dp, sz, dp, dp, sz, sz, dp, dp, sz, sz, dbl, dbl, dp, dp, dp, u8p, u64p]
        L.orc_hmc_ensemble_pt.argtypes = [ci, dp, sz, dp, dp, sz, sz, C.c_void_p, dp, dp, sz, sz, dbl, dbl, u64p]
        L.orc_trace_iat.restype = None
        L.orc_trace_iat.argtypes = [dp, sz, sz, dbl, sz, dp, u32p, dp, dp, dp]
        L.orc_running_moments.restype = None
        L.orc_running_moments.argtypes = [dp, sz, sz, sz, dp, dp]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _u32p(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint32)) if a is not None else None


def _u8p(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8)) if a is not None else None


def _params(params):
    p = np.ascontiguousarray(np.asarray(params if params is not None else [], dtype=np.float64).ravel())
    return p, _dp(p) if p.size else None, p.size


class Rng:
    """The oracle's own seeded generator (xoshiro256**)."""

    def __init__(self, seed: int):
        self._state = (C.c_uint64 * 4)()
        lib().orc_rng_seed(self._state, seed)

    @property
    def ptr(self):
        return C.cast(self._state, C.c_void_p)

    def uniform(self) -> float:
        return lib().orc_rng_uniform(self.ptr)


def logprob(kind, params, x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    p, pp, npar = _params(params)
    if x.ndim == 1:
        return lib().orc_logprob(kind, pp, npar, _dp(x), x.shape[0])
    out = np.empty(x.shape[0])
    lib().orc_init_logprob(kind, pp, npar, _dp(x), x.shape[0], x.shape[1], _dp(out))
    return out


def init_logprob(kind, params, ens):
    return logprob(kind, params, ens)


def z_from_p(p, a):
    return lib().orc_z_from_p(p, a)


def z_range(a):
    return lib().orc_z_range(a)


def propose_move(p1, p2, z, flags=None):
    p1 = np.ascontiguousarray(p1, dtype=np.float64)
    p2 = np.ascontiguousarray(p2, dtype=np.float64)
    out = np.empty_like(p1)
    f = np.ascontiguousarray(flags, dtype=np.uint8) if flags is not None else None
    lib().orc_propose_move(_dp(p1), _dp(p2), z, _u8p(f), f.size if f is not None else 0, p1.size, _dp(out))
    return out


def exchange_prob(lp1, lp2, beta1, beta2):
    return lib().orc_exchange_prob(lp1, lp2, beta1, beta2)


def sample_pt_fixed(kind, params, ens, lp, beta, partner, z, flags, u, want_proposals=False):
    """In place on ens/lp (float64 C-contiguous).  Returns (rc, accepted[, proposed, new_lp])."""
    n, dim = ens.shape
    assert ens.flags.c_contiguous and ens.dtype == np.float64 and lp.dtype == np.float64
    beta = np.ascontiguousarray(beta, dtype=np.float64)
    partner = np.ascontiguousarray(partner, dtype=np.uint32)
    z = np.ascontiguousarray(z, dtype=np.float64)
    u = np.ascontiguousarray(u, dtype=np.float64)
    flags = np.ascontiguousarray(flags, dtype=np.uint8)
    flag_len = flags.shape[1] if flags.ndim == 2 else dim
    p, pp, npar = _params(params)
    acc = np.zeros(n, dtype=np.uint8)
    prop = np.empty((n, dim)) if want_proposals else None
    nlp = np.empty(n) if want_proposals else None
    rc = lib().orc_sample_pt_fixed(kind, pp, npar, _dp(ens), _dp(lp), n, lp.shape[0], dim, _dp(beta),
                                   beta.shape[0], _u32p(partner), _dp(z), _u8p(flags), flag_len,
                                   _dp(u), _u8p(acc), _dp(prop), _dp(nlp))
    if want_proposals:
        return rc, acc, prop, nlp
    return rc, acc


def swap_walkers_fixed(ens, lp, beta, jvec, r):
    n, dim = ens.shape
    beta = np.ascontiguousarray(beta, dtype=np.float64)
    jvec = np.ascontiguousarray(jvec, dtype=np.uint32)
    r = np.ascontiguousarray(r, dtype=np.float64)
    sw = np.zeros(max(jvec.size, 1), dtype=np.uint8)
    rc = lib().orc_swap_walkers_fixed(_dp(ens), _dp(lp), n, lp.shape[0], dim, _dp(beta), beta.shape[0],
                                      _u32p(jvec), _dp(r), _u8p(sw))
    return rc, sw[:jvec.size].reshape(jvec.shape)


def draw_sample_streams(rng: Rng, n, dim, nbeta, a, ufs_kind, prob=0.5, onehot_counter=0):
    partner = np.zeros(n, dtype=np.uint32)
    z = np.zeros(n)
    u = np.zeros(n)
    flags = np.zeros((n, dim), dtype=np.uint8)
    oc = C.c_uint64(onehot_counter)
    rc = lib().orc_draw_sample_streams(rng.ptr, n, dim, nbeta, a, ufs_kind, prob, C.byref(oc),
                                       _u32p(partner), _dp(z), _u8p(flags), _dp(u))
    if rc:
        raise ValueError(f"orc_draw_sample_streams rc={rc}")
    return partner, z, flags, u, oc.value


def draw_swap_streams(rng: Rng, n, nbeta):
    npb = n // nbeta
    jvec = np.zeros((max(nbeta - 1, 0), npb), dtype=np.uint32)
    r = np.zeros((max(nbeta - 1, 0), npb))
    rc = lib().orc_draw_swap_streams(rng.ptr, n, nbeta, _u32p(jvec), _dp(r))
    if rc:
        raise ValueError(f"orc_draw_swap_streams rc={rc}")
    return jvec, r


def sample_pt(kind, params, ens, lp, rng: Rng, a, ufs_kind, prob, beta, onehot_counter=0):
    n, dim = ens.shape
    beta = np.ascontiguousarray(beta, dtype=np.float64)
    p, pp, npar = _params(params)
    oc = C.c_uint64(onehot_counter)
    rc = lib().orc_sample_pt(kind, pp, npar, _dp(ens), _dp(lp), n, dim, rng.ptr, a, ufs_kind, prob,
                             C.byref(oc), _dp(beta), beta.shape[0])
    return rc, oc.value


def swap_walkers(ens, lp, rng: Rng, beta):
    n, dim = ens.shape
    beta = np.ascontiguousarray(beta, dtype=np.float64)
    return lib().orc_swap_walkers(_dp(ens), _dp(lp), n, dim, rng.ptr, _dp(beta), beta.shape[0])


def mean(ens):
    n, dim = ens.shape
    out = np.empty(dim)
    lib().orc_mean(_dp(np.ascontiguousarray(ens)), n, dim, _dp(out))
    return out


def cov(ens):
    n, dim = ens.shape
    out = np.empty((dim, dim))
    lib().orc_cov(_dp(np.ascontiguousarray(ens)), n, dim, _dp(out))
    return out


def twalk_params(dim, pphi=None, fw=None, aw=None, at=None) -> TwalkParams:
    """TWalkParams::new(dim) (twalk.rs:89-110), with_pphi (:126-128), with_fw (:112-124: plain weights in)."""
    t = TwalkParams()
    lib().orc_twalk_params_default(dim, C.byref(t))
    if pphi is not None:
        t.pphi = pphi
    if fw is not None:
        t.fw[0], t.fw[1] = fw[0], fw[0] + fw[1]
    if aw is not None:
        t.aw = aw
    if at is not None:
        t.at = at
    return t


def twalk_half_fixed(kind, params, ens, lp, beta, tw, order, half, kernel, b, flags, walk_u, u, want_proposals=False,
                     literal_index=False):
    """One half-pass of twalk::sample1 with supplied draws, in place; returns accepted[n/2] (+ proposals, new_lp).
    literal_index: the reference's index arithmetic as written (reads rung 0's walkers for every rung)."""
    n, dim = ens.shape
    beta = np.ascontiguousarray(beta, dtype=np.float64)
    p, pp, npar = _params(params)
    order = np.ascontiguousarray(order, dtype=np.uint32)
    kernel = np.ascontiguousarray(kernel, dtype=np.uint8)
    b = np.ascontiguousarray(b, dtype=np.float64)
    flags = np.ascontiguousarray(flags, dtype=np.uint8)
    walk_u = np.ascontiguousarray(walk_u, dtype=np.float64)
    u = np.ascontiguousarray(u, dtype=np.float64)
    acc = np.zeros(n // 2, dtype=np.uint8)
    prop = np.zeros((n // 2, dim)) if want_proposals else None
    nlp = np.zeros(n // 2) if want_proposals else None
    fn = lib().orc_twalk_half_fixed_literal if literal_index else lib().orc_twalk_half_fixed
    rc = fn(kind, pp, npar, _dp(ens), _dp(lp), n, dim, _dp(beta), beta.shape[0], C.byref(tw),
                                    _u32p(order), half, _u8p(kernel), _dp(b), _u8p(flags), _dp(walk_u), _dp(u),
                                    _u8p(acc), _dp(prop), _dp(nlp))
    if rc:
        raise ValueError(f"orc_twalk_half_fixed rc={rc}")
    return (acc, prop, nlp) if want_proposals else acc


def twalk_draw_streams(rng: Rng, n, dim, nbeta, tw):
    """Draws of one twalk::sample call in the reference's order: order[n], then kernel / b / flags / walk_u / u with a
    leading axis of 2 (the half-passes)."""
    order = np.zeros(n, dtype=np.uint32)
    kernel = np.zeros((2, n // 2), dtype=np.uint8)
    b, u = np.zeros((2, n // 2)), np.zeros((2, n // 2))
    flags = np.zeros((2, n // 2, dim), dtype=np.uint8)
    walk_u = np.zeros((2, n // 2, dim))
    rc = lib().orc_twalk_draw_streams(rng.ptr, n, dim, nbeta, C.byref(tw), _u32p(order), _u8p(kernel), _dp(b), _u8p(flags),
                                      _dp(walk_u), _dp(u))
    if rc:
        raise ValueError(f"orc_twalk_draw_streams rc={rc}")
    return order, kernel, b, flags, walk_u, u


def twalk_sample(kind, params, ens, lp, rng: Rng, tw, beta):
    """twalk::sample (twalk.rs:586-648) driven by the oracle's generator in the reference's draw order."""
    n, dim = ens.shape
    beta = np.ascontiguousarray(beta, dtype=np.float64)
    p, pp, npar = _params(params)
    return lib().orc_twalk_sample(kind, pp, npar, _dp(ens), _dp(lp), n, dim, rng.ptr, C.byref(tw), _dp(beta), beta.shape[0])


def philox_twalk_streams(seed, step, order_step, n, dim, nbeta, tw, half):
    """The device streams of one t-walk half-pass: (order, kernel, b, flags, walk_u, u), per pair."""
    order = np.zeros(n, dtype=np.uint32)
    kernel = np.zeros(n // 2, dtype=np.uint8)
    b, u = np.zeros(n // 2), np.zeros(n // 2)
    flags = np.zeros((n // 2, dim), dtype=np.uint8)
    walk_u = np.zeros((n // 2, dim))
    rc = lib().orc_philox_twalk_streams(seed, step, order_step, n, dim, nbeta, C.byref(tw), half, _u32p(order),
                                        _u8p(kernel), _dp(b), _u8p(flags), _dp(walk_u), _dp(u))
    if rc:
        raise ValueError(f"orc_philox_twalk_streams rc={rc}")
    return order, kernel, b, flags, walk_u, u


def grad_logprob(kind, params, x):
    """Gradient of the plugin's log-density at every row of x[n][dim]."""
    x = np.ascontiguousarray(np.atleast_2d(x), dtype=np.float64)
    p, pp, npar = _params(params)
    out = np.empty_like(x)
    for i in range(x.shape[0]):
        row = np.ascontiguousarray(x[i])
        g = np.empty(x.shape[1])
        rc = lib().orc_grad_logprob(kind, pp, npar, _dp(row), x.shape[1], _dp(g))
        if rc:
            raise ValueError(f"orc_grad_logprob rc={rc}")
        out[i] = g
    return out


def hmc_ensemble_pt_fixed(kind, params, q0, lp, epsilon, beta, l, target, adj, p0, u1, u2):
    """sample_ensemble_pt (naive.rs:122-292) with supplied momenta / uniforms, in place (q0, lp, epsilon);
    returns (accepted[n], accepted_cnt[nbeta])."""
    n, dim = q0.shape
    beta = np.ascontiguousarray(beta, dtype=np.float64)
    p, pp, npar = _params(params)
    p0 = np.ascontiguousarray(p0, dtype=np.float64)
    u1 = np.ascontiguousarray(u1, dtype=np.float64)
    u2 = np.ascontiguousarray(u2, dtype=np.float64)
    acc = np.zeros(n, dtype=np.uint8)
    cnt = np.zeros(beta.shape[0], dtype=np.uint64)
    rc = lib().orc_hmc_ensemble_pt_fixed(kind, pp, npar, _dp(q0), _dp(lp), n, dim, _dp(epsilon), _dp(beta), beta.shape[0], l,
                                         target, adj, _dp(p0), _dp(u1), _dp(u2), _u8p(acc),
                                         cnt.ctypes.data_as(C.POINTER(C.c_uint64)))
    if rc:
        raise ValueError(f"orc_hmc_ensemble_pt_fixed rc={rc}")
    return acc, cnt


def hmc_ensemble_pt(kind, params, q0, lp, rng: Rng, epsilon, beta, l, target, adj):
    n, dim = q0.shape
    beta = np.ascontiguousarray(beta, dtype=np.float64)
    p, pp, npar = _params(params)
    cnt = np.zeros(beta.shape[0], dtype=np.uint64)
    rc = lib().orc_hmc_ensemble_pt(kind, pp, npar, _dp(q0), _dp(lp), n, dim, rng.ptr, _dp(epsilon), _dp(beta), beta.shape[0],
                                   l, target, adj, cnt.ctypes.data_as(C.POINTER(C.c_uint64)))
    if rc:
        raise ValueError(f"orc_hmc_ensemble_pt rc={rc}")
    return cnt


def trace_iat(trace, c_window=5.0, max_lag=0):
    """(tau, window, mean, var) of every column of trace[T][ncap]; bit-identical to scorus_trace_iat."""
    trace = np.ascontiguousarray(trace, dtype=np.float64)
    T, ncap = trace.shape
    tau, mean_, var = np.empty(ncap), np.empty(ncap), np.empty(ncap)
    window = np.empty(ncap, dtype=np.uint32)
    scratch = np.empty(T)
    lib().orc_trace_iat(_dp(trace), T, ncap, float(c_window), int(max_lag), _dp(tau), _u32p(window), _dp(mean_),
                        _dp(var), _dp(scratch))
    return tau, window, mean_, var


def running_moments(samples, want_cov=True):
    """mean / biased cov of samples[nsnap][npb][dim] over all nsnap * npb rows (one rung)."""
    samples = np.ascontiguousarray(samples, dtype=np.float64)
    nsnap, npb, dim = samples.shape
    mean_ = np.empty(dim)
    cov_ = np.empty((dim, dim)) if want_cov else None
    lib().orc_running_moments(_dp(samples), nsnap, npb, dim, _dp(mean_), _dp(cov_) if want_cov else None)
    return mean_, cov_


def philox4x32_10(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().orc_philox4x32_10(c, k, o)
    return list(o)


def philox_sample_streams(seed, step, n, dim, nbeta, a, ufs_kind, prob=0.5, onehot_counter=0):
    partner = np.zeros(n, dtype=np.uint32)
    z = np.zeros(n)
    u = np.zeros(n)
    flags = np.zeros((n, dim), dtype=np.uint8)
    rc = lib().orc_philox_sample_streams(seed, step, n, dim, nbeta, a, ufs_kind, prob, onehot_counter,
                                         _u32p(partner), _dp(z), _u8p(flags), _dp(u))
    if rc:
        raise ValueError(f"orc_philox_sample_streams rc={rc}")
    return partner, z, flags, u


def philox_sample_streams_ex(seed, step, n, dim, nbeta, a, ufs_kind, prob=0.5, onehot_counter=0, nblocks=1, walkers=None):
    """philox_sample_streams with `nblocks` separately matched blocks per rung and, if `walkers` is given, for that
    sample of walkers only (outputs in the order of `walkers`; partner holds global ids)."""
    w = None if walkers is None else np.ascontiguousarray(walkers, dtype=np.uint32)
    m = n if w is None else w.size
    partner = np.zeros(m, dtype=np.uint32)
    z = np.zeros(m)
    u = np.zeros(m)
    flags = np.zeros((m, dim), dtype=np.uint8)
    rc = lib().orc_philox_sample_streams_ex(seed, step, n, dim, nbeta, a, ufs_kind, prob, onehot_counter, nblocks,
                                            _u32p(w), m, _u32p(partner), _dp(z), _u8p(flags), _dp(u))
    if rc:
        raise ValueError(f"orc_philox_sample_streams_ex rc={rc}")
    return partner, z, flags, u


def philox_swap_streams(seed, swap_call, n, nbeta):
    npb = n // nbeta
    jvec = np.zeros((max(nbeta - 1, 0), npb), dtype=np.uint32)
    r = np.zeros((max(nbeta - 1, 0), npb))
    rc = lib().orc_philox_swap_streams(seed, swap_call, n, nbeta, _u32p(jvec), _dp(r))
    if rc:
        raise ValueError(f"orc_philox_swap_streams rc={rc}")
    return jvec, r


def philox_init_ensemble(seed, n, dim, lo, hi, walker_offset=0):
    lo = np.ascontiguousarray(lo, dtype=np.float64)
    hi = np.ascontiguousarray(hi, dtype=np.float64)
    out = np.empty((n, dim))
    lib().orc_philox_init_ensemble(seed, n, dim, walker_offset, _dp(lo), _dp(hi), _dp(out))
    return out


def ref_structure_run(kind, params, ens, lp, beta, a, ufs_kind, prob, seed, nsteps, nthreads=0):
    """CPU baseline with the reference's execution structure. Returns (seconds, stage_seconds[6])."""
    n, dim = ens.shape
    beta = np.ascontiguousarray(beta, dtype=np.float64)
    p, pp, npar = _params(params)
    stage = np.zeros(6)
    t = lib().orc_ref_structure_run(kind, pp, npar, _dp(ens), _dp(lp), n, dim, _dp(beta), beta.shape[0],
                                    a, ufs_kind, prob, seed, nsteps, nthreads, _dp(stage))
    return t, stage


def flat_threaded_run(kind, params, ens, lp, beta, a, ufs_kind, prob, seed, first_step, nsteps, nthreads=0):
    n, dim = ens.shape
    beta = np.ascontiguousarray(beta, dtype=np.float64)
    p, pp, npar = _params(params)
    return lib().orc_flat_threaded_run(kind, pp, npar, _dp(ens), _dp(lp), n, dim, _dp(beta),
                                       beta.shape[0], a, ufs_kind, prob, seed, first_step, nsteps, nthreads)


def max_threads() -> int:
    return lib().orc_max_threads()
