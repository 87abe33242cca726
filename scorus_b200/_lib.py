"""ctypes loader for libscorus_b200.so (the C ABI of include/scorus_b200.h).

The product has no CPU path: if the shared library is missing this raises, it does not fall
back to anything.  Build it with `python -c "import __graft_entry__ as g; g.build()"` or
`make -C scorus_b200/csrc`.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SCORUS_B200_LIB") or os.path.join(_HERE, "lib", "libscorus_b200.so")


class Cfg(C.Structure):
    _fields_ = [("n_walkers", C.c_uint64), ("dim", C.c_uint32), ("device", C.c_int32),
                ("seed", C.c_uint64), ("flags", C.c_uint32), ("reserved", C.c_uint32)]


class Ufs(C.Structure):
    _fields_ = [("kind", C.c_int32), ("reserved", C.c_int32), ("prob", C.c_double),
                ("mask", C.POINTER(C.c_uint8)), ("mask_len", C.c_size_t)]


# status codes (include/scorus_b200.h)
OK = 0
ERR_NWALKERS_IS_ZERO, ERR_NWALKERS_IS_NOT_EVEN, ERR_NWALKERS_MISMATCHES_NBETA, ERR_BETA_NOT_IN_DECR_ORD = 1, 2, 3, 4
ERR_LENGTH_MISMATCH, ERR_INVALID_ARG, ERR_NPHI_ZERO, ERR_NO_LOGPROB, ERR_NOT_UPLOADED, ERR_DIM_TOO_LARGE = 5, 6, 7, 8, 9, 10
ERR_CUDA, ERR_NO_DEVICE = 100, 101

LP_GAUSSIAN, LP_BOUNDED_GAUSSIAN, LP_ROSENBROCK, LP_RASTRIGIN = 0, 1, 2, 3
LP_BIMODAL2D, LP_STEP2D, LP_CORR_GAUSSIAN, LP_MIXTURE = 4, 5, 6, 7
UFS_PROB, UFS_ALL, UFS_ONE_HOT_CYCLE, UFS_MASK = 0, 1, 2, 3
CFG_SEQUENTIAL_SUM = 1
CFG_F32 = 2

class TwalkParams(C.Structure):
    """scorus_twalk_params (include/scorus_b200.h): TWalkParams of src/mcmc/twalk.rs:75-83, fw cumulative."""
    _fields_ = [("aw", C.c_double), ("at", C.c_double), ("pphi", C.c_double), ("fw", C.c_double * 2)]


class HmcParam(C.Structure):
    """scorus_hmc_param (include/scorus_b200.h): HmcParam of src/mcmc/hmc/naive.rs:17-23."""
    _fields_ = [("target_accept_ratio", C.c_double), ("adj_factor", C.c_double)]


_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built "
            "(run __graft_entry__.build() or `make -C scorus_b200/csrc`); there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp, sz, u64, dbl, ci = C.c_void_p, C.c_size_t, C.c_uint64, C.c_double, C.c_int
    dp, u32p, u8p, u64p = (C.POINTER(C.c_double), C.POINTER(C.c_uint32), C.POINTER(C.c_uint8),
                           C.POINTER(C.c_uint64))
    fp = C.POINTER(C.c_float)
    sig = {
        "scorus_abi_version": (ci, []),
        "scorus_status_string": (C.c_char_p, [ci]),
        "scorus_device_count": (ci, []),
        "scorus_check_sample_args": (ci, [sz, sz, sz]),
        "scorus_check_beta_order": (ci, [dp, sz]),
        "scorus_ctx_create": (ci, [C.POINTER(vp), C.POINTER(Cfg)]),
        "scorus_ctx_destroy": (None, [vp]),
        "scorus_last_error": (C.c_char_p, [vp]),
        "scorus_set_logprob": (ci, [vp, ci, dp, sz]),
        "scorus_upload": (ci, [vp, dp, dp]),
        "scorus_download": (ci, [vp, dp, dp]),
        "scorus_init_logprob": (ci, [vp]),
        "scorus_sample": (ci, [vp, dbl, C.POINTER(Ufs), u64]),
        "scorus_sample_pt": (ci, [vp, dbl, C.POINTER(Ufs), dp, sz, u64]),
        "scorus_sample_pt_fixed": (ci, [vp, dp, sz, u32p, dp, u8p, sz, dp, u8p]),
        "scorus_draw_z": (ci, [vp, dbl, sz, dp]),
        "scorus_swap_walkers": (ci, [vp, dp, sz]),
        "scorus_swap_walkers_fixed": (ci, [vp, dp, sz, u32p, dp, u8p]),
        "scorus_sample_pt_host": (ci, [vp, dp, dp, sz, dbl, C.POINTER(Ufs), dp, sz]),
        "scorus_sample_pt_host_n": (ci, [vp, dp, dp, sz, dbl, C.POINTER(Ufs), dp, sz, u64]),
        "scorus_swap_walkers_host": (ci, [vp, dp, dp, sz, dp, sz]),
        "scorus_init_ensemble": (ci, [vp, dp, dp]),
        "scorus_mean": (ci, [vp, sz, sz, dp]),
        "scorus_cov": (ci, [vp, sz, sz, dp]),
        "scorus_trace_enable": (ci, [vp, u32p, u32p, sz, sz]),
        "scorus_trace_download": (ci, [vp, dp, sz, C.POINTER(sz)]),
        "scorus_trace_stride": (ci, [vp, u64]),
        "scorus_twalk_params_default": (ci, [sz, C.POINTER(TwalkParams)]),
        "scorus_twalk_sample": (ci, [vp, C.POINTER(TwalkParams), dp, sz, u64]),
        "scorus_twalk_half_fixed": (ci, [vp, C.POINTER(TwalkParams), dp, sz, u32p, ci, u8p, dp, u8p, dp, dp, u8p]),
        "scorus_twalk_sample_host": (ci, [vp, dp, dp, sz, C.POINTER(TwalkParams), dp, sz, u64]),
        "scorus_twalk_draw_b": (ci, [vp, dbl, u64, sz, dp]),
        "scorus_hmc_sample_ensemble_pt": (ci, [vp, dp, dp, sz, sz, C.POINTER(HmcParam), u64p, u64]),
        "scorus_hmc_sample_ensemble_pt_fixed": (ci, [vp, dp, dp, sz, sz, C.POINTER(HmcParam), dp, dp, dp, u8p, u64p]),
        "scorus_hmc_sample_ensemble_pt_host": (ci, [vp, dp, dp, sz, dp, dp, sz, sz, C.POINTER(HmcParam), u64p]),
        "scorus_hmc_draw_momenta": (ci, [vp, u64, sz, dp]),
        "scorus_get_rung_stats": (ci, [vp, sz, u64p, u64p, u64p, u64p]),
        "scorus_moments_enable": (ci, [vp, sz, u64, ci]),
        "scorus_moments_download": (ci, [vp, sz, dp, dp, u64p]),
        "scorus_trace_iat": (ci, [vp, dbl, sz, dp, u32p, dp, dp]),
        "scorus_set_shard": (ci, [vp, u64, C.c_uint32, u64]),
        "scorus_sample_pt_global": (ci, [vp, dbl, C.POINTER(Ufs), dp, sz, vp, vp]),
        "scorus_swap_plan_device": (ci, [vp, dp, sz, vp, vp]),
        "scorus_ipc_export": (ci, [vp, vp, vp]),
        "scorus_ipc_attach": (ci, [vp, ci, ci, vp, vp]),
        "scorus_sample_pt_peer": (ci, [vp, dbl, C.POINTER(Ufs), dp, sz, u64]),
        "scorus_upload_f32": (ci, [vp, fp, fp]),
        "scorus_download_f32": (ci, [vp, fp, fp]),
        "scorus_sample_pt_fixed_f32": (ci, [vp, dp, sz, u32p, fp, u8p, sz, fp, u8p]),
        "scorus_swap_walkers_fixed_f32": (ci, [vp, dp, sz, u32p, fp, u8p]),
        "scorus_sample_pt_peer_host": (ci, [vp, dp, dp, sz, dbl, C.POINTER(Ufs), dp, sz, u64]),
        "scorus_hmc_adapt_net": (ci, [vp, C.POINTER(C.c_int64)]),
        "scorus_peer_barrier": (ci, [vp]),
        "scorus_swap_walkers_peer": (ci, [vp, dp, sz]),
        "scorus_peer_read_bandwidth": (ci, [vp, ci, sz, ci, dp, dp, dp]),
        "scorus_set_matching_blocks": (ci, [vp, C.c_uint32, C.c_uint32, u64]),
        "scorus_swap_apply_p2p": (ci, [vp, vp, vp, ci]),
        "scorus_sync": (ci, [vp]),
        "scorus_set_stream": (ci, [vp, vp]),
        "scorus_reset_stream": (ci, [vp]),
        "scorus_get_stream": (vp, [vp]),
        "scorus_device_buffers": (ci, [vp, C.POINTER(dp), C.POINTER(sz), C.POINTER(dp)]),
        "scorus_get_counters": (ci, [vp, u64p, u64p, u64p]),
        "scorus_set_counters": (ci, [vp, u64, u64, u64]),
        "scorus_get_stats": (ci, [vp, u64p, u64p, u64p, u64p]),
        "scorus_launch_count": (u64, [vp]),
        "scorus_host_alloc": (vp, [sz]),
        "scorus_host_free": (None, [vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


EXPORTS = [
    "scorus_abi_version", "scorus_status_string", "scorus_device_count", "scorus_check_sample_args",
    "scorus_check_beta_order", "scorus_ctx_create", "scorus_ctx_destroy", "scorus_last_error",
    "scorus_set_logprob", "scorus_upload", "scorus_download", "scorus_init_logprob", "scorus_sample",
    "scorus_sample_pt", "scorus_sample_pt_fixed", "scorus_draw_z", "scorus_swap_walkers", "scorus_swap_walkers_fixed",
    "scorus_sample_pt_host", "scorus_sample_pt_host_n", "scorus_swap_walkers_host", "scorus_init_ensemble", "scorus_mean", "scorus_cov",
    "scorus_trace_enable", "scorus_trace_download", "scorus_trace_stride", "scorus_get_rung_stats",
    "scorus_moments_enable", "scorus_moments_download", "scorus_trace_iat", "scorus_twalk_params_default",
    "scorus_twalk_sample", "scorus_twalk_half_fixed", "scorus_twalk_sample_host", "scorus_twalk_draw_b", "scorus_hmc_sample_ensemble_pt",
    "scorus_hmc_sample_ensemble_pt_fixed", "scorus_hmc_sample_ensemble_pt_host", "scorus_hmc_draw_momenta", "scorus_set_shard", "scorus_sample_pt_global",
    "scorus_swap_plan_device", "scorus_ipc_export", "scorus_ipc_attach",
    "scorus_sample_pt_peer", "scorus_upload_f32", "scorus_download_f32", "scorus_sample_pt_fixed_f32", "scorus_swap_walkers_fixed_f32",
    "scorus_sample_pt_peer_host", "scorus_hmc_adapt_net", "scorus_peer_barrier", "scorus_swap_walkers_peer", "scorus_peer_read_bandwidth", "scorus_set_matching_blocks", "scorus_swap_apply_p2p", "scorus_sync", "scorus_set_stream", "scorus_reset_stream",
    "scorus_get_stream", "scorus_device_buffers", "scorus_get_counters", "scorus_set_counters",
    "scorus_get_stats", "scorus_launch_count", "scorus_host_alloc", "scorus_host_free",
]
