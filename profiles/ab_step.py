#!/usr/bin/env python
"""A/B timing of the single-GPU step through the C ABI of a GIVEN library file (raw ctypes: works with any build of
libscorus_b200.so that has the round-1 entry points), so that two builds can be compared on the same box:
  python profiles/ab_step.py path/to/lib.so [c5|c2|c3|c4]"""
import ctypes as C
import sys
import time

import numpy as np


class Cfg(C.Structure):
    _fields_ = [("n_walkers", C.c_uint64), ("dim", C.c_uint32), ("device", C.c_int32), ("seed", C.c_uint64),
                ("flags", C.c_uint32), ("reserved", C.c_uint32)]


class Ufs(C.Structure):
    _fields_ = [("kind", C.c_int32), ("reserved", C.c_int32), ("prob", C.c_double), ("mask", C.POINTER(C.c_uint8)),
                ("mask_len", C.c_size_t)]


W = {"c5": (2, 1 << 22, 64, 1, 0, 0.5, (0.9, 1.1)), "c3": (7, 1 << 20, 32, 1, 0, 0.5, (-6.0, 6.0)),
     "c4": (3, 1 << 20, 10, 64, 0, 0.2, (-0.01, 0.01)), "c2": (6, 65536, 100, 1, 1, 0.0, (-1.0, 1.0))}
lib = C.CDLL(sys.argv[1])
dp = C.POINTER(C.c_double)
for wl in sys.argv[2:] or ["c5"]:
    kind, n, dim, nb, ufs_kind, prob, box = W[wl]
    if kind == 6:
        rho, L = 0.9, np.zeros((dim, dim))
        sc = 1.0 / np.sqrt(1.0 - rho * rho)
        for i in range(dim - 1):
            L[i, i], L[i, i + 1] = sc, -rho * sc
        L[dim - 1, dim - 1] = 1.0
        params = np.concatenate([np.zeros(dim), L.ravel()])
    elif kind == 7:
        params = np.array([1.0, 2.0, 5.0] + [0.0] * (dim - 1))
    elif kind == 3:
        params = np.array([10.0])
    else:
        params = np.zeros(0)
    ctx = C.c_void_p()
    cfg = Cfg(n, dim, -1, 2024, 0, 0)
    assert lib.scorus_ctx_create(C.byref(ctx), C.byref(cfg)) == 0
    assert lib.scorus_set_logprob(ctx, kind, params.ctypes.data_as(dp) if params.size else None, C.c_size_t(params.size)) == 0
    lo, hi = np.full(dim, box[0]), np.full(dim, box[1])
    assert lib.scorus_init_ensemble(ctx, lo.ctypes.data_as(dp), hi.ctypes.data_as(dp)) == 0
    beta = (2.0 ** -(np.arange(nb) / max(1, nb // 8))).astype(np.float64)
    u = Ufs(ufs_kind, 0, prob, None, 0)
    call = lambda k: lib.scorus_sample_pt(ctx, C.c_double(2.0), C.byref(u), beta.ctypes.data_as(dp), C.c_size_t(nb), C.c_uint64(k))
    assert call(300) == 0
    lib.scorus_sync(ctx)
    best = 1e9
    for rep in range(5):
        t0 = time.perf_counter()
        call(300)
        lib.scorus_sync(ctx)
        best = min(best, (time.perf_counter() - t0) / 300)
    print(f"{sys.argv[1]} {wl}: {1e3 * best:.4f} ms/step (best of 5 x 300 steps, wall clock around a synchronise)")
    lib.scorus_ctx_destroy(ctx)
