"""Host-side mirror of scorus::mcmc::ensemble_sample (src/mcmc/ensemble_sample.rs) over the C ABI.

Same names and argument order as the reference:

    init_logprob(flogprob, ensemble, logprob)                                   # :77
    sample(flogprob, ensemble, cached_logprob, rng, a, ufs)                      # :87
    sample_pt(flogprob, ensemble, cached_logprob, rng, a, ufs, beta_list)        # :107
    propose_move(p1, p2, z, update_flags)                                        # :57
    UpdateFlagSpec.Prob(p) / .All() / .Func(f)                                   # :25

with two substitutions the device forces:
  * `flogprob` is a registered device plugin (`LogProb` instances below) instead of a closure;
  * `rng` is a `DeviceRng(seed)`: the key of the counter-based device streams.  It also caches
    the device context, so repeated calls on same-shaped ensembles reuse their device buffers.

`ensemble` / `cached_logprob` are numpy float64 arrays [n, dim] / [n], updated in place like the
reference's `&mut` slices.  Where the reference panics, these raise `McmcErr`
(src/mcmc/mcmc_errors.rs) or ValueError.  Long runs should keep the state on the device with
`Sampler` and download when needed; the free functions move the ensemble both ways per call.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Callable, Optional, Sequence

import numpy as np

from . import _lib as L


class McmcErr(RuntimeError):
    """src/mcmc/mcmc_errors.rs:1-7; `.code` is the C ABI status."""

    NAMES = {1: "NWalkersIsZero", 2: "NWalkersIsNotEven", 3: "NWalkersMismatchesNBeta", 4: "BetaNotInDecrOrd"}

    def __init__(self, code: int, msg: str = ""):
        self.code = code
        self.variant = self.NAMES.get(code)
        super().__init__(f"{self.variant or 'status ' + str(code)}: {msg}")


def _raise(code: int, ctx=None):
    if code == L.OK:
        return
    lib = L.load()
    msg = lib.scorus_status_string(code).decode()
    if ctx is not None:
        detail = lib.scorus_last_error(ctx).decode()
        if detail:
            msg = f"{msg} ({detail})"
    if 1 <= code <= 4:
        raise McmcErr(code, msg)
    if code == L.ERR_LENGTH_MISMATCH:
        raise AssertionError(msg)  # assert!(ensemble.len() == cached_logprob.len()), :133
    if code in (L.ERR_CUDA, L.ERR_NO_DEVICE):
        raise RuntimeError(msg)
    raise ValueError(msg)


# ------------------------------------------------------------------------------------------------
# log-density plugins (replace the `flogprob: &F` closure)
# ------------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class LogProb:
    kind: int
    params: tuple = ()

    def param_array(self) -> np.ndarray:
        return np.ascontiguousarray(np.asarray(self.params, dtype=np.float64).ravel())


def Gaussian(c: float = 0.5) -> LogProb:
    """-sum c x^2: examples/test_emcee.rs:18-24 (c=1/2), examples/sample_high_dim.rs:22-28 (c=1)."""
    return LogProb(L.LP_GAUSSIAN, (float(c),))


def BoundedGaussian(limit: float = 1.5) -> LogProb:
    """examples/sample_bimodal.rs:35-44"""
    return LogProb(L.LP_BOUNDED_GAUSSIAN, (float(limit),))


def Rosenbrock() -> LogProb:
    """examples/test_emcee.rs:26-32"""
    return LogProb(L.LP_ROSENBROCK)


def Rastrigin(a: float = 10.0) -> LogProb:
    """examples/sample_rastrigin.rs:17-25"""
    return LogProb(L.LP_RASTRIGIN, (float(a),))


def Bimodal2d(*params: float) -> LogProb:
    """examples/sample_bimodal.rs:46-54 (optionally the 9 constants)."""
    return LogProb(L.LP_BIMODAL2D, tuple(float(p) for p in params))


def Step2d() -> LogProb:
    """examples/sample_bimodal.rs:56-65"""
    return LogProb(L.LP_STEP2D)


def CorrGaussian(mu, Lfac) -> LogProb:
    """-0.5 |L (x - mu)|^2 with L^T L the precision matrix (BASELINE config 2; not in the reference)."""
    mu = np.asarray(mu, dtype=np.float64).ravel()
    Lfac = np.asarray(Lfac, dtype=np.float64)
    assert Lfac.shape == (mu.size, mu.size)
    return LogProb(L.LP_CORR_GAUSSIAN, tuple(np.concatenate([mu, Lfac.ravel()]).tolist()))


def Mixture(sigma1: float, sigma2: float, m) -> LogProb:
    """logsumexp of N(+m, sigma1^2 I) and N(-m, sigma2^2 I) up to constants (BASELINE config 3)."""
    m = np.asarray(m, dtype=np.float64).ravel()
    return LogProb(L.LP_MIXTURE, (float(sigma1), float(sigma2)) + tuple(m.tolist()))


# ------------------------------------------------------------------------------------------------
# UpdateFlagSpec (:25-55)
# ------------------------------------------------------------------------------------------------
class UpdateFlagSpec:
    """enum UpdateFlagSpec { Prob(T), All, Func(&mut dyn FnMut() -> Vec<bool>) }"""

    def __init__(self, kind: int, prob: float = 0.0, func: Optional[Callable[[], Sequence[bool]]] = None, mask=None):
        self.kind, self.prob, self.func, self.mask = kind, prob, func, mask

    @classmethod
    def Prob(cls, p: float) -> "UpdateFlagSpec":
        return cls(L.UFS_PROB, prob=float(p))

    @classmethod
    def All(cls) -> "UpdateFlagSpec":
        return cls(L.UFS_ALL)

    @classmethod
    def Func(cls, f: Callable[[], Sequence[bool]]) -> "UpdateFlagSpec":
        """General closure: called once per walker per step on the host (:150-153), the flags are
        uploaded as a byte matrix.  For the cycling one-hot closure of
        examples/sample_high_dim.rs:59-65 use OneHotCycle(), which stays on the device."""
        return cls(L.UFS_MASK, func=f)

    @classmethod
    def Mask(cls, flags) -> "UpdateFlagSpec":
        """The flags of ONE step as already evaluated, [n_walkers, flag_len] (what Func's closure returned, walker by
        walker): how a sharded run hands each shard -- or the whole-ensemble step -- its rows of one evaluation."""
        m = np.ascontiguousarray(np.asarray(flags).astype(np.uint8))
        if m.ndim != 2:
            raise ValueError("flags must be [n_walkers, flag_len]")
        return cls(L.UFS_MASK, mask=m)

    def evaluate(self, n: int) -> np.ndarray:
        """Func: the closure called once per walker, in order (ensemble_sample.rs:150-153) -> [n, flag_len] uint8."""
        rows = [np.asarray(self.func(), dtype=bool) for _ in range(n)]
        lens = {r.size for r in rows}
        if len(lens) != 1:
            raise ValueError("Func must return flag vectors of one length per step")
        return np.ascontiguousarray(np.stack(rows).astype(np.uint8))

    @classmethod
    def OneHotCycle(cls) -> "UpdateFlagSpec":
        return cls(L.UFS_ONE_HOT_CYCLE)

    def _c(self, n: int, dim: int):
        """-> (Ufs struct, keepalive)"""
        u = L.Ufs()
        u.kind = self.kind
        u.prob = self.prob
        keep = None
        if self.kind == L.UFS_MASK:
            keep = self.mask if self.mask is not None else self.evaluate(n)
            if keep.shape[0] != n:
                raise ValueError(f"mask has {keep.shape[0]} rows for {n} walkers")
            flen = keep.shape[1]
            if flen > dim:
                raise IndexError("update_flags longer than the walker (index panic, :70)")
            u.mask = keep.ctypes.data_as(C.POINTER(C.c_uint8))
            u.mask_len = flen
        return u, keep


# ------------------------------------------------------------------------------------------------
# device context
# ------------------------------------------------------------------------------------------------
def _dp(a: Optional[np.ndarray]):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _check_arrays(ensemble: np.ndarray, logprob: Optional[np.ndarray], dtype=np.float64):
    name = np.dtype(dtype).name
    if not (isinstance(ensemble, np.ndarray) and ensemble.dtype == dtype and ensemble.ndim == 2
            and ensemble.flags.c_contiguous):
        raise TypeError(f"ensemble must be a C-contiguous {name} array [n_walkers, dim]")
    if logprob is not None and not (isinstance(logprob, np.ndarray) and logprob.dtype == dtype
                                    and logprob.ndim == 1 and logprob.flags.c_contiguous):
        raise TypeError(f"logprob must be a C-contiguous {name} array [n_walkers]")


def _fp(a: Optional[np.ndarray]):
    return a.ctypes.data_as(C.POINTER(C.c_float)) if a is not None else None


class Sampler:
    """RAII handle over scorus_ctx: the device-resident (ensemble, logprob) plus RNG counters."""

    def __init__(self, flogprob: LogProb, n_walkers: int, dim: int, seed: int = 0, device: int = -1,
                 sequential_sum: bool = False, dtype=np.float64):
        """sequential_sum=True selects SCORUS_CFG_SEQUENTIAL_SUM: one lane per walker, log-density
        summed left to right like the reference closures (bit-identical log-probs, ~2x slower).
        dtype=np.float32 selects SCORUS_CFG_F32 (`T = f32`, ensemble_sample.rs:95): upload / download /
        sample_pt_fixed / swap_walkers_fixed then take float32 arrays; sample, sample_pt, swap_walkers and
        init_logprob are the same calls."""
        self._lib = L.load()
        self._ctx = C.c_void_p()
        self.dtype = np.dtype(dtype)
        if self.dtype not in (np.dtype(np.float64), np.dtype(np.float32)):
            raise TypeError("dtype must be float64 or float32")
        self.f32 = self.dtype == np.dtype(np.float32)
        cfg = L.Cfg(n_walkers=n_walkers, dim=dim, device=device, seed=seed,
                    flags=(L.CFG_SEQUENTIAL_SUM if sequential_sum else 0) | (L.CFG_F32 if self.f32 else 0), reserved=0)
        _raise(self._lib.scorus_ctx_create(C.byref(self._ctx), C.byref(cfg)))
        self.n, self.dim, self.seed = n_walkers, dim, seed
        self.n_global = n_walkers  # (set_shard: this context holds a shard of a larger ensemble)
        self.flogprob = None
        self.set_logprob(flogprob)

    # -- lifetime
    def close(self):
        if getattr(self, "_ctx", None) and self._ctx.value:
            self._lib.scorus_ctx_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ok(self, code: int):
        _raise(code, self._ctx)

    # -- configuration
    def set_logprob(self, flogprob: LogProb):
        if flogprob is self.flogprob:
            return
        p = flogprob.param_array()
        self._ok(self._lib.scorus_set_logprob(self._ctx, flogprob.kind, _dp(p) if p.size else None, p.size))
        self.flogprob = flogprob

    def set_stream(self, cuda_stream: int):
        """Launch on a caller-owned cudaStream_t handle (0 = CUDA's legacy default stream)."""
        self._ok(self._lib.scorus_set_stream(self._ctx, C.c_void_p(cuda_stream)))

    def reset_stream(self):
        self._ok(self._lib.scorus_reset_stream(self._ctx))

    # -- data movement
    def upload(self, ensemble: np.ndarray, logprob: Optional[np.ndarray] = None):
        _check_arrays(ensemble, logprob, self.dtype)
        if ensemble.shape != (self.n, self.dim):
            raise ValueError("ensemble shape does not match the sampler")
        if logprob is not None and logprob.shape[0] != self.n:
            raise AssertionError("ensemble and logprob lengths differ")
        if self.f32:
            self._ok(self._lib.scorus_upload_f32(self._ctx, _fp(ensemble), _fp(logprob)))
        else:
            self._ok(self._lib.scorus_upload(self._ctx, _dp(ensemble), _dp(logprob)))
        self._ok(self._lib.scorus_sync(self._ctx))

    def download(self, ensemble: Optional[np.ndarray] = None, logprob: Optional[np.ndarray] = None):
        if ensemble is None:
            ensemble = np.empty((self.n, self.dim), dtype=self.dtype)
        if logprob is None:
            logprob = np.empty(self.n, dtype=self.dtype)
        _check_arrays(ensemble, logprob, self.dtype)
        if self.f32:
            self._ok(self._lib.scorus_download_f32(self._ctx, _fp(ensemble), _fp(logprob)))
        else:
            self._ok(self._lib.scorus_download(self._ctx, _dp(ensemble), _dp(logprob)))
        return ensemble, logprob

    def download_logprob(self) -> np.ndarray:
        """Only the cached log-probs (one scalar per walker): scorus_download with a NULL ensemble."""
        lp = np.empty(self.n, dtype=self.dtype)
        if self.f32:
            self._ok(self._lib.scorus_download_f32(self._ctx, None, _fp(lp)))
        else:
            self._ok(self._lib.scorus_download(self._ctx, None, _dp(lp)))
        return lp

    # -- the hot path
    def init_logprob(self):
        self._ok(self._lib.scorus_init_logprob(self._ctx))

    def sample_pt(self, a: float, ufs: UpdateFlagSpec, beta_list, nsteps: int = 1):
        beta = np.ascontiguousarray(beta_list, dtype=np.float64)
        if ufs.kind == L.UFS_MASK:
            for _ in range(nsteps):
                u, keep = ufs._c(self.n, self.dim)
                self._ok(self._lib.scorus_sample_pt(self._ctx, a, C.byref(u), _dp(beta), beta.size, 1))
            return
        u, keep = ufs._c(self.n, self.dim)
        self._ok(self._lib.scorus_sample_pt(self._ctx, a, C.byref(u), _dp(beta), beta.size, nsteps))

    def sample(self, a: float, ufs: UpdateFlagSpec, nsteps: int = 1):
        self.sample_pt(a, ufs, [1.0], nsteps)

    def sample_pt_fixed(self, beta_list, partner, z, flags, u):
        beta = np.ascontiguousarray(beta_list, dtype=np.float64)
        partner = np.ascontiguousarray(partner, dtype=np.uint32)
        z = np.ascontiguousarray(z, dtype=self.dtype)
        u = np.ascontiguousarray(u, dtype=self.dtype)
        flags = np.ascontiguousarray(flags, dtype=np.uint8)
        if partner.shape != (self.n,) or z.shape != (self.n,) or u.shape != (self.n,) or flags.shape[0] != self.n:
            raise ValueError("stream arrays must have one entry per walker")
        flag_len = flags.shape[1] if flags.ndim == 2 else self.dim
        acc = np.zeros(self.n, dtype=np.uint8)
        if self.f32:
            self._ok(self._lib.scorus_sample_pt_fixed_f32(
                self._ctx, _dp(beta), beta.size, partner.ctypes.data_as(C.POINTER(C.c_uint32)), _fp(z),
                flags.ctypes.data_as(C.POINTER(C.c_uint8)), flag_len, _fp(u),
                acc.ctypes.data_as(C.POINTER(C.c_uint8))))
            return acc
        self._ok(self._lib.scorus_sample_pt_fixed(
            self._ctx, _dp(beta), beta.size, partner.ctypes.data_as(C.POINTER(C.c_uint32)), _dp(z),
            flags.ctypes.data_as(C.POINTER(C.c_uint8)), flag_len, _dp(u),
            acc.ctypes.data_as(C.POINTER(C.c_uint8))))
        return acc

    def draw_z(self, a: float, count: int = 1) -> np.ndarray:
        out = np.empty(count)
        self._ok(self._lib.scorus_draw_z(self._ctx, a, count, _dp(out)))
        return out

    def swap_walkers(self, beta_list):
        beta = np.ascontiguousarray(beta_list, dtype=np.float64)
        self._ok(self._lib.scorus_swap_walkers(self._ctx, _dp(beta), beta.size))

    def swap_walkers_fixed(self, beta_list, jvec, r):
        beta = np.ascontiguousarray(beta_list, dtype=np.float64)
        jvec = np.ascontiguousarray(jvec, dtype=np.uint32)
        r = np.ascontiguousarray(r, dtype=self.dtype)
        sw = np.zeros(max(jvec.size, 1), dtype=np.uint8)
        if self.f32:
            self._ok(self._lib.scorus_swap_walkers_fixed_f32(
                self._ctx, _dp(beta), beta.size, jvec.ctypes.data_as(C.POINTER(C.c_uint32)), _fp(r),
                sw.ctypes.data_as(C.POINTER(C.c_uint8))))
            return sw[:jvec.size].reshape(jvec.shape)
        self._ok(self._lib.scorus_swap_walkers_fixed(
            self._ctx, _dp(beta), beta.size, jvec.ctypes.data_as(C.POINTER(C.c_uint32)), _dp(r),
            sw.ctypes.data_as(C.POINTER(C.c_uint8))))
        return sw[:jvec.size].reshape(jvec.shape)

    def sample_pt_host(self, ensemble, logprob, a, ufs: UpdateFlagSpec, beta_list, nsteps: int = 1):
        """One reference-shaped call: host buffers in, one step (or nsteps, for drivers that record
        every k-th iteration), host buffers out."""
        _check_arrays(ensemble, logprob)
        beta = np.ascontiguousarray(beta_list, dtype=np.float64)
        u, keep = ufs._c(self.n, self.dim)
        self._ok(self._lib.scorus_sample_pt_host_n(self._ctx, _dp(ensemble), _dp(logprob), logprob.shape[0], a,
                                                   C.byref(u), _dp(beta), beta.size, nsteps))

    def sample_pt_peer_host(self, ensemble, logprob, a, ufs: UpdateFlagSpec, beta_list, nsteps: int = 1):
        """The reference-shaped call for one shard of a multi-GPU ensemble (collective: every rank calls it with its
        own shard's host buffers): upload, one-sided whole-ensemble steps, write-back of the accepted walkers."""
        _check_arrays(ensemble, logprob)
        beta = np.ascontiguousarray(beta_list, dtype=np.float64)
        u, keep = ufs._c(self.n, self.dim)
        self._ok(self._lib.scorus_sample_pt_peer_host(self._ctx, _dp(ensemble), _dp(logprob), logprob.shape[0], a,
                                                      C.byref(u), _dp(beta), beta.size, nsteps))

    def swap_walkers_host(self, ensemble, logprob, beta_list):
        _check_arrays(ensemble, logprob)
        beta = np.ascontiguousarray(beta_list, dtype=np.float64)
        self._ok(self._lib.scorus_swap_walkers_host(self._ctx, _dp(ensemble), _dp(logprob), logprob.shape[0],
                                                    _dp(beta), beta.size))

    # -- periphery: box init, ensemble statistics, chain trace
    def init_ensemble(self, lo, hi):
        """get_one_init_realization (src/mcmc/init_ensemble.rs:19-32) for every walker, on the device."""
        lo = np.ascontiguousarray(np.broadcast_to(np.asarray(lo, dtype=np.float64), (self.dim,)))
        hi = np.ascontiguousarray(np.broadcast_to(np.asarray(hi, dtype=np.float64), (self.dim,)))
        self._ok(self._lib.scorus_init_ensemble(self._ctx, _dp(lo), _dp(hi)))

    def mean(self, first: int = 0, count: Optional[int] = None) -> np.ndarray:
        """linear_space::utils::mean (src/linear_space/utils.rs:4-15) of walkers [first, first+count)."""
        out = np.empty(self.dim)
        self._ok(self._lib.scorus_mean(self._ctx, first, self.n - first if count is None else count, _dp(out)))
        return out

    def cov(self, first: int = 0, count: Optional[int] = None) -> np.ndarray:
        """linear_space::utils::cov (src/linear_space/utils.rs:17-41): biased (/N) covariance."""
        out = np.empty((self.dim, self.dim))
        self._ok(self._lib.scorus_cov(self._ctx, first, self.n - first if count is None else count, _dp(out)))
        return out

    def trace_enable(self, walkers, coords, capacity_steps: int):
        """Capture ens[walkers[c]][coords[c]] after every step (the examples' recording loops)."""
        w = np.ascontiguousarray(walkers, dtype=np.uint32)
        c = np.ascontiguousarray(coords, dtype=np.uint32)
        self._trace_ncap = w.size
        self._ok(self._lib.scorus_trace_enable(self._ctx, w.ctypes.data_as(C.POINTER(C.c_uint32)),
                                               c.ctypes.data_as(C.POINTER(C.c_uint32)), w.size, capacity_steps))

    def trace_download(self, max_steps: int) -> np.ndarray:
        out = np.empty((max_steps, self._trace_ncap))
        k = C.c_size_t()
        self._ok(self._lib.scorus_trace_download(self._ctx, _dp(out), max_steps, C.byref(k)))
        return out[:min(k.value, max_steps)]

    def trace_stride(self, every: int):
        """Capture after every `every`-th step only (examples/test_emcee.rs:82-84 keeps every 10th)."""
        self._ok(self._lib.scorus_trace_stride(self._ctx, every))

    # -- chain statistics on the device
    def rung_stats(self, nrungs: int):
        """Accepted / proposed moves per rung and accepted / proposed exchanges per ladder stage (entry k =
        rung k <-> rung k+1) since the sampler was created."""
        acc, prop = np.zeros(nrungs, dtype=np.uint64), np.zeros(nrungs, dtype=np.uint64)
        sw, swp = np.zeros(max(nrungs - 1, 1), dtype=np.uint64), np.zeros(max(nrungs - 1, 1), dtype=np.uint64)
        u64p = C.POINTER(C.c_uint64)
        self._ok(self._lib.scorus_get_rung_stats(self._ctx, nrungs, acc.ctypes.data_as(u64p), prop.ctypes.data_as(u64p),
                                                 sw.ctypes.data_as(u64p), swp.ctypes.data_as(u64p)))
        return dict(accepted=acc, proposed=prop, swapped=sw[:nrungs - 1], swap_proposed=swp[:nrungs - 1])

    def moments_enable(self, nrungs: int, every: int = 1, with_cov: bool = True):
        """Accumulate per-rung mean / covariance (src/linear_space/utils.rs:4-41, over steps x walkers) on the
        device after every `every`-th step from now on; nrungs = 0 stops."""
        self._ok(self._lib.scorus_moments_enable(self._ctx, nrungs, every, 1 if with_cov else 0))
        self._moments_cov = bool(with_cov)

    def moments(self, rung: int = 0):
        """(mean, cov or None, nsamples) of one rung from the running sums."""
        mean = np.empty(self.dim)
        cov = np.empty((self.dim, self.dim)) if getattr(self, "_moments_cov", False) else None
        m = C.c_uint64()
        self._ok(self._lib.scorus_moments_download(self._ctx, rung, _dp(mean), _dp(cov) if cov is not None else None,
                                                   C.byref(m)))
        return mean, cov, m.value

    def trace_iat(self, c_window: float = 5.0, max_lag: int = 0):
        """Integrated autocorrelation time of every captured series (Sokal window M >= c_window * tau):
        dict(tau, window, mean, var), one entry per captured (walker, coordinate)."""
        k = self._trace_ncap
        tau, mean, var = np.empty(k), np.empty(k), np.empty(k)
        window = np.empty(k, dtype=np.uint32)
        self._ok(self._lib.scorus_trace_iat(self._ctx, c_window, max_lag, _dp(tau),
                                            window.ctypes.data_as(C.POINTER(C.c_uint32)), _dp(mean), _dp(var)))
        return dict(tau=tau, window=window, mean=mean, var=var)

    # -- t-walk, src/mcmc/twalk.rs (scorus_b200/twalk.py mirrors the module)
    def twalk_sample(self, param, beta_list, nsteps: int = 1):
        """twalk::sample (twalk.rs:586-648) on the resident ensemble, nsteps times."""
        beta = np.ascontiguousarray(beta_list, dtype=np.float64)
        self._ok(self._lib.scorus_twalk_sample(self._ctx, C.byref(param._c()), _dp(beta), beta.size, nsteps))

    def twalk_half_fixed(self, param, beta_list, order, half, kernel, b, flags, walk_u, u):
        """One half-pass (sample1, twalk.rs:650-762) with supplied draws, per pair; returns accepted[n/2]."""
        beta = np.ascontiguousarray(beta_list, dtype=np.float64)
        order = np.ascontiguousarray(order, dtype=np.uint32)
        kernel = np.ascontiguousarray(kernel, dtype=np.uint8)
        b = np.ascontiguousarray(b, dtype=np.float64)
        flags = np.ascontiguousarray(flags, dtype=np.uint8)
        walk_u = np.ascontiguousarray(walk_u, dtype=np.float64)
        u = np.ascontiguousarray(u, dtype=np.float64)
        npair = self.n // 2
        if (order.shape != (self.n,) or kernel.shape != (npair,) or b.shape != (npair,) or u.shape != (npair,)
                or flags.shape != (npair, self.dim) or walk_u.shape != (npair, self.dim)):
            raise ValueError("t-walk stream arrays must have one entry per pair (order: one per walker)")
        acc = np.zeros(max(npair, 1), dtype=np.uint8)
        u8p, u32p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint32)
        self._ok(self._lib.scorus_twalk_half_fixed(
            self._ctx, C.byref(param._c()), _dp(beta), beta.size, order.ctypes.data_as(u32p), int(half),
            kernel.ctypes.data_as(u8p), _dp(b), flags.ctypes.data_as(u8p), _dp(walk_u), _dp(u), acc.ctypes.data_as(u8p)))
        return acc[:npair]

    def twalk_sample_host(self, ensemble, logprob, param, beta_list, nsteps: int = 1):
        _check_arrays(ensemble, logprob)
        beta = np.ascontiguousarray(beta_list, dtype=np.float64)
        self._ok(self._lib.scorus_twalk_sample_host(self._ctx, _dp(ensemble), _dp(logprob), logprob.shape[0],
                                                    C.byref(param._c()), _dp(beta), beta.size, nsteps))

    def twalk_draw_b(self, at: float, step: int, count: Optional[int] = None) -> np.ndarray:
        """sim_b (twalk.rs:306-322) of walkers [0, count) from the device stream at step counter `step`."""
        count = self.n if count is None else count
        out = np.empty(count)
        self._ok(self._lib.scorus_twalk_draw_b(self._ctx, at, step, count, _dp(out)))
        return out

    # -- ensemble / parallel-tempering HMC, src/mcmc/hmc/naive.rs (scorus_b200/hmc.py mirrors the module)
    def hmc_sample_ensemble_pt(self, epsilon, beta_list, l: int, param, nsteps: int = 1) -> np.ndarray:
        """sample_ensemble_pt (naive.rs:122-292) on the resident ensemble; `epsilon` (float64 array, one per rung)
        is updated in place; returns accepted_cnt per rung."""
        beta = np.ascontiguousarray(beta_list, dtype=np.float64)
        _check_epsilon(epsilon, beta.size)
        cnt = np.zeros(beta.size, dtype=np.uint64)
        self._ok(self._lib.scorus_hmc_sample_ensemble_pt(self._ctx, _dp(epsilon), _dp(beta), beta.size, l, C.byref(param._c()),
                                                         cnt.ctypes.data_as(C.POINTER(C.c_uint64)), nsteps))
        return cnt

    def hmc_sample_ensemble_pt_fixed(self, epsilon, beta_list, l: int, param, p0, u1, u2):
        """One step with supplied momenta p0[n][dim], accept uniforms u1[n] and adaptation uniforms u2[n];
        returns (accepted[n], accepted_cnt[nbeta])."""
        beta = np.ascontiguousarray(beta_list, dtype=np.float64)
        _check_epsilon(epsilon, beta.size)
        p0 = np.ascontiguousarray(p0, dtype=np.float64)
        u1 = np.ascontiguousarray(u1, dtype=np.float64)
        u2 = np.ascontiguousarray(u2, dtype=np.float64)
        if p0.shape != (self.n, self.dim) or u1.shape != (self.n,) or u2.shape != (self.n,):
            raise ValueError("p0 must be [n][dim], u1 and u2 [n]")
        acc = np.zeros(self.n, dtype=np.uint8)
        cnt = np.zeros(beta.size, dtype=np.uint64)
        self._ok(self._lib.scorus_hmc_sample_ensemble_pt_fixed(
            self._ctx, _dp(epsilon), _dp(beta), beta.size, l, C.byref(param._c()), _dp(p0), _dp(u1), _dp(u2),
            acc.ctypes.data_as(C.POINTER(C.c_uint8)), cnt.ctypes.data_as(C.POINTER(C.c_uint64))))
        return acc, cnt

    def hmc_sample_ensemble_pt_host(self, q0, lp, epsilon, beta_list, l: int, param) -> np.ndarray:
        _check_arrays(q0, lp)
        beta = np.ascontiguousarray(beta_list, dtype=np.float64)
        _check_epsilon(epsilon, beta.size)
        cnt = np.zeros(beta.size, dtype=np.uint64)
        self._ok(self._lib.scorus_hmc_sample_ensemble_pt_host(self._ctx, _dp(q0), _dp(lp), lp.shape[0], _dp(epsilon), _dp(beta),
                                                              beta.size, l, C.byref(param._c()),
                                                              cnt.ctypes.data_as(C.POINTER(C.c_uint64))))
        return cnt

    def hmc_adapt_net(self) -> int:
        """#grow - #shrink of the adaptation codes of the last HMC step (scorus_hmc_adapt_net)."""
        v = C.c_int64(0)
        self._ok(self._lib.scorus_hmc_adapt_net(self._ctx, C.byref(v)))
        return int(v.value)

    def hmc_draw_momenta(self, step: int, count: Optional[int] = None) -> np.ndarray:
        """Momenta of walkers [0, count) from the device stream at step counter `step`, [count][dim]."""
        count = self.n if count is None else count
        out = np.empty((count, self.dim))
        self._ok(self._lib.scorus_hmc_draw_momenta(self._ctx, step, count, _dp(out)))
        return out

    # -- sharding (one process per GPU; scorus_b200/distributed.py drives these)
    def set_shard(self, walker_offset: int, rung_offset: int, n_walkers_global: int):
        self._ok(self._lib.scorus_set_shard(self._ctx, walker_offset, rung_offset, n_walkers_global))
        self.n_global = int(n_walkers_global)

    def sample_pt_global(self, a: float, ufs: UpdateFlagSpec, beta_list, ens_all_ptr: int, lp_all_ptr: int):
        beta = np.ascontiguousarray(beta_list, dtype=np.float64)
        u, keep = ufs._c(self.n_global if ufs.kind == L.UFS_MASK else self.n, self.dim)  # a mask covers the whole ensemble
        self._ok(self._lib.scorus_sample_pt_global(self._ctx, a, C.byref(u), _dp(beta), beta.size,
                                                   C.c_void_p(ens_all_ptr), C.c_void_p(lp_all_ptr)))

    def swap_plan_device(self, beta_list, lp_all_ptr: int, src_all_ptr: int):
        beta = np.ascontiguousarray(beta_list, dtype=np.float64)
        self._ok(self._lib.scorus_swap_plan_device(self._ctx, _dp(beta), beta.size, C.c_void_p(lp_all_ptr),
                                                   C.c_void_p(src_all_ptr)))

    def ipc_export(self):
        he, hl = C.create_string_buffer(64), C.create_string_buffer(64)
        self._ok(self._lib.scorus_ipc_export(self._ctx, he, hl))
        return he.raw, hl.raw

    def ipc_attach(self, world: int, rank: int, handles_ens, handles_lp):
        be, bl = b"".join(handles_ens), b"".join(handles_lp)
        self._ok(self._lib.scorus_ipc_attach(self._ctx, world, rank, be, bl))

    def sample_pt_peer(self, a: float, ufs: UpdateFlagSpec, beta_list, nsteps: int = 1):
        """Whole-ensemble matching over walker shards, one-sided (scorus_sample_pt_peer): the owner of a pair reads the
        partner's row from its home GPU over NVLink and writes both walkers' results back."""
        beta = np.ascontiguousarray(beta_list, dtype=np.float64)
        u, keep = ufs._c(self.n_global if ufs.kind == L.UFS_MASK else self.n, self.dim)  # a mask covers the whole ensemble
        self._ok(self._lib.scorus_sample_pt_peer(self._ctx, a, C.byref(u), _dp(beta), beta.size, nsteps))

    def peer_read_bandwidth(self, peer: int, nbytes: int = 0, iters: int = 5):
        """(copy-engine, SM sequential loads, SM loads of whole rows in a scrambled order) GB/s of this GPU reading rank
        `peer`'s ensemble over NVLink (measurement only)."""
        ce, sm, rr = C.c_double(0.0), C.c_double(0.0), C.c_double(0.0)
        self._ok(self._lib.scorus_peer_read_bandwidth(self._ctx, peer, nbytes, iters, C.byref(ce), C.byref(sm), C.byref(rr)))
        return ce.value, sm.value, rr.value

    def swap_walkers_peer(self, beta_list):
        """swap_walkers over rung shards, one-sided (scorus_swap_walkers_peer); beta_list is the whole ladder."""
        beta = np.ascontiguousarray(beta_list, dtype=np.float64)
        self._ok(self._lib.scorus_swap_walkers_peer(self._ctx, _dp(beta), beta.size))

    def peer_barrier(self):
        self._ok(self._lib.scorus_peer_barrier(self._ctx))

    def set_matching_blocks(self, nblocks: int = 1, block_offset: int = 0, global_every: int = 0):
        """Blocked matchings (scorus_set_matching_blocks): `nblocks` separately matched blocks per rung, whole-rung
        matching on the steps whose index is a multiple of `global_every` (0: never)."""
        self._ok(self._lib.scorus_set_matching_blocks(self._ctx, nblocks, block_offset, global_every))

    def swap_apply_p2p(self, src_all_ptr: int, lp_all_ptr: int, phase: int):
        self._ok(self._lib.scorus_swap_apply_p2p(self._ctx, C.c_void_p(src_all_ptr), C.c_void_p(lp_all_ptr), phase))

    # -- plumbing
    def sync(self):
        self._ok(self._lib.scorus_sync(self._ctx))

    @property
    def counters(self):
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self._ok(self._lib.scorus_get_counters(self._ctx, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    @counters.setter
    def counters(self, v):
        self._ok(self._lib.scorus_set_counters(self._ctx, int(v[0]), int(v[1]), int(v[2])))

    def stats(self):
        v = [C.c_uint64() for _ in range(4)]
        self._ok(self._lib.scorus_get_stats(self._ctx, *[C.byref(x) for x in v]))
        return dict(accepted=v[0].value, proposed=v[1].value, swapped=v[2].value, swap_proposed=v[3].value)

    @property
    def launch_count(self) -> int:
        return self._lib.scorus_launch_count(self._ctx)

    def device_buffers(self):
        e, lp, pitch = C.POINTER(C.c_double)(), C.POINTER(C.c_double)(), C.c_size_t()
        self._ok(self._lib.scorus_device_buffers(self._ctx, C.byref(e), C.byref(pitch), C.byref(lp)))
        return C.cast(e, C.c_void_p).value, pitch.value, C.cast(lp, C.c_void_p).value


@dataclass
class DeviceRng:
    """Stands in for the reference's `rng: &mut U`: the seed of the device streams.  Caches one
    Sampler per (n, dim, plugin) so the free functions below do not re-create device state."""
    seed: int = 0
    device: int = -1
    sequential_sum: bool = False
    _samplers: dict = field(default_factory=dict, repr=False)

    def sampler(self, flogprob: LogProb, n: int, dim: int) -> Sampler:
        key = (n, dim)
        s = self._samplers.get(key)
        if s is None:
            s = Sampler(flogprob, n, dim, seed=self.seed, device=self.device, sequential_sum=self.sequential_sum)
            self._samplers[key] = s
        else:
            s.set_logprob(flogprob)
        return s


def _check_epsilon(epsilon, nbeta: int):
    if not (isinstance(epsilon, np.ndarray) and epsilon.dtype == np.float64 and epsilon.flags.c_contiguous
            and epsilon.shape == (nbeta,)):
        raise AssertionError("assert_eq!(epsilon.len(), nbeta): epsilon must be a float64 array with one entry per rung")


def pinned_empty(shape, dtype=np.float64) -> np.ndarray:
    """numpy array over pinned host memory (scorus_host_alloc) for full-speed host calls."""
    lib = L.load()
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = lib.scorus_host_alloc(n)
    if not p:
        raise MemoryError("scorus_host_alloc failed")
    buf = (C.c_char * n).from_address(p)
    arr = np.frombuffer(buf, dtype=dtype).reshape(shape)
    _PINNED[id(buf)] = (buf, p)
    return arr


_PINNED: dict = {}


# ------------------------------------------------------------------------------------------------
# the reference's free functions
# ------------------------------------------------------------------------------------------------
def init_logprob(flogprob: LogProb, ensemble: np.ndarray, logprob: np.ndarray, rng: Optional[DeviceRng] = None):
    """src/mcmc/ensemble_sample.rs:77-85"""
    _check_arrays(ensemble, logprob)
    if logprob.shape[0] != ensemble.shape[0]:
        raise ValueError("copy_from_slice: source and destination lengths differ")  # :84
    rng = rng or DeviceRng()
    s = rng.sampler(flogprob, *ensemble.shape)
    s.upload(ensemble, None)
    s.download(None, logprob)


def sample_pt(flogprob: LogProb, ensemble: np.ndarray, cached_logprob: np.ndarray, rng: DeviceRng, a: float,
              ufs: UpdateFlagSpec, beta_list):
    """src/mcmc/ensemble_sample.rs:107-190, in place on the host arrays."""
    _check_arrays(ensemble, cached_logprob)
    n = ensemble.shape[0]
    nb = len(beta_list)
    _raise(L.load().scorus_check_sample_args(n, cached_logprob.shape[0], nb))
    s = rng.sampler(flogprob, *ensemble.shape)
    s.sample_pt_host(ensemble, cached_logprob, a, ufs, beta_list)


def sample(flogprob: LogProb, ensemble: np.ndarray, cached_logprob: np.ndarray, rng: DeviceRng, a: float,
           ufs: UpdateFlagSpec):
    """src/mcmc/ensemble_sample.rs:87-105"""
    sample_pt(flogprob, ensemble, cached_logprob, rng, a, ufs, [1.0])


def propose_move(p1, p2, z: float, update_flags=None) -> np.ndarray:
    """src/mcmc/ensemble_sample.rs:57-75, evaluated by the fused step kernel on a two-walker
    ensemble with the accept forced (u = 0 < q for any finite log-density)."""
    p1 = np.ascontiguousarray(p1, dtype=np.float64)
    p2 = np.ascontiguousarray(p2, dtype=np.float64)
    d = p1.size
    ens = np.ascontiguousarray(np.stack([p1, p2]))
    if update_flags is None:
        fl = np.ones((2, d), dtype=np.uint8)
    else:
        f = np.asarray(update_flags, dtype=np.uint8)
        fl = np.ascontiguousarray(np.stack([f, f]))
    with Sampler(LogProb(L.LP_GAUSSIAN, (0.0,)), 2, d) as s:
        s.upload(ens, np.zeros(2))
        s.sample_pt_fixed([1.0], [1, 0], [z, z], fl if fl.any() else np.ones((2, d), dtype=np.uint8),
                          [0.0, 0.0])
        out, _ = s.download()
    if update_flags is not None and not fl.any():
        return p1.copy()
    return out[0]
