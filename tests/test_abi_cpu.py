"""No-GPU checks of the boundary: the library loads, exports every symbol the header declares,
the precondition checks mirror the reference's asserts, the host mirrors keep the reference's
names / argument order, and the product never reaches into oracle/."""
import ctypes as C
import inspect
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "scorus_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(scorus_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(sb):
    from scorus_b200 import _lib
    lib = _lib.load()
    declared = _header_functions()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in include/scorus_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == declared, "scorus_b200/_lib.py EXPORTS out of sync with the header"
    assert lib.scorus_abi_version() == 1


def test_rust_sys_and_cpp_mirror_cover_the_header():
    """The Rust -sys crate (source only: no toolchain here) declares every function of the header and nothing else;
    the C++ mirror includes the header itself, so it only has to parse."""
    rs = open(os.path.join(ROOT, "rust", "scorus-b200-sys", "src", "lib.rs")).read()
    declared_rs = sorted(set(re.findall(r"pub fn (scorus_[a-z0-9_]+)", rs)))
    assert declared_rs == _header_functions()
    for struct in ("scorus_cfg", "scorus_ufs", "scorus_twalk_params", "scorus_hmc_param"):
        assert f"pub struct {struct}" in rs
    import subprocess
    env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
    subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", "driver_like_reference.cpp")], check=True, env=env)
    subprocess.run(["gcc", "-std=c99", "-fsyntax-only", "-x", "c", os.path.join(ROOT, "include", "scorus_b200.h")],
                   check=True, env=env)


def test_status_strings_and_mcmc_err_codes(sb):
    from scorus_b200 import _lib
    lib = _lib.load()
    # src/mcmc/mcmc_errors.rs:1-7, declaration order
    assert [lib.scorus_status_string(i).decode() for i in (1, 2, 3, 4)] == [
        "NWalkersIsZero", "NWalkersIsNotEven", "NWalkersMismatchesNBeta", "BetaNotInDecrOrd"]
    assert lib.scorus_status_string(0) == b"ok"


def test_precondition_checks_follow_assert_order(sb):
    from scorus_b200 import _lib
    lib = _lib.load()
    chk = lib.scorus_check_sample_args
    assert chk(16, 16, 1) == 0
    assert chk(12, 12, 5) == _lib.ERR_NWALKERS_MISMATCHES_NBETA  # ensemble_sample.rs:130
    assert chk(0, 0, 1) == _lib.ERR_NWALKERS_IS_ZERO              # :131
    assert chk(12, 12, 4) == _lib.ERR_NWALKERS_IS_NOT_EVEN        # :132
    assert chk(12, 11, 2) == _lib.ERR_LENGTH_MISMATCH             # :133
    assert chk(13, 12, 5) == _lib.ERR_NWALKERS_MISMATCHES_NBETA   # earlier assert wins
    assert chk(12, 12, 0) == _lib.ERR_INVALID_ARG

    def order(b):
        a = np.asarray(b, dtype=np.float64)
        return lib.scorus_check_beta_order(a.ctypes.data_as(C.POINTER(C.c_double)), a.size)
    assert order([1.0, 0.5, 0.25]) == 0
    assert order([1.0, 1.0, 0.5]) == 0                            # equal betas pass (utils.rs:93 is `>`)
    assert order([1.0, 0.25, 0.5]) == _lib.ERR_BETA_NOT_IN_DECR_ORD


def test_no_cpu_fallback_without_a_device(sb):
    from scorus_b200 import _lib
    lib = _lib.load()
    if lib.scorus_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(RuntimeError, match="no CUDA device"):
        sb.Sampler(sb.Gaussian(), 16, 2)
    ctx = C.c_void_p()
    cfg = _lib.Cfg(n_walkers=16, dim=2, device=-1, seed=0, flags=0, reserved=0)
    assert lib.scorus_ctx_create(C.byref(ctx), C.byref(cfg)) == _lib.ERR_NO_DEVICE and not ctx.value
    cfg = _lib.Cfg(n_walkers=0, dim=2, device=-1, seed=0, flags=0, reserved=0)
    assert lib.scorus_ctx_create(C.byref(ctx), C.byref(cfg)) == _lib.ERR_NWALKERS_IS_ZERO


def test_host_mirror_keeps_reference_signatures(sb):
    es, ut = sb.ensemble_sample, sb.utils
    # src/mcmc/ensemble_sample.rs:87-94, 107-115, 77, 57; src/mcmc/utils.rs:72, 33, 47
    assert list(inspect.signature(es.sample).parameters) == ["flogprob", "ensemble", "cached_logprob", "rng", "a", "ufs"]
    assert list(inspect.signature(es.sample_pt).parameters) == ["flogprob", "ensemble", "cached_logprob", "rng", "a",
                                                                "ufs", "beta_list"]
    assert list(inspect.signature(es.init_logprob).parameters)[:3] == ["flogprob", "ensemble", "logprob"]
    assert list(inspect.signature(es.propose_move).parameters) == ["p1", "p2", "z", "update_flags"]
    assert list(inspect.signature(ut.swap_walkers).parameters)[:4] == ["ensemble", "logprob", "rng", "beta_list"]
    assert list(inspect.signature(ut.draw_z).parameters) == ["rng", "a"]
    assert list(inspect.signature(ut.scale_vec).parameters) == ["x1", "x2", "z"]
    for variant in ("Prob", "All", "Func"):  # enum UpdateFlagSpec, ensemble_sample.rs:25-32
        assert hasattr(es.UpdateFlagSpec, variant)


def test_update_flag_func_is_called_once_per_walker(sb):
    from scorus_b200 import _lib
    calls = []

    def f():  # the cycling closure of examples/sample_high_dim.rs:59-65
        calls.append(1)
        i = len(calls) % 4
        return [j == i for j in range(4)]
    u, keep = sb.UpdateFlagSpec.Func(f)._c(6, 4)
    assert len(calls) == 6 and u.kind == _lib.UFS_MASK and u.mask_len == 4
    assert keep.tolist() == [[0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1], [1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0]]
    with pytest.raises(IndexError):
        sb.UpdateFlagSpec.Func(lambda: [True] * 5)._c(2, 4)  # more flags than coordinates panics (:70)
    p = sb.UpdateFlagSpec.Prob(0.2)._c(6, 4)[0]
    assert p.kind == _lib.UFS_PROB and p.prob == 0.2


def test_reference_shaped_errors_raise_before_touching_the_device(sb):
    ens, lp = np.zeros((12, 2)), np.zeros(12)
    dev = sb.DeviceRng(seed=1)
    with pytest.raises(sb.McmcErr) as e:
        sb.sample_pt(sb.Gaussian(), ens, lp, dev, 2.0, sb.UpdateFlagSpec.All(), [1.0] * 5)
    assert e.value.variant == "NWalkersMismatchesNBeta" and e.value.code == 3
    with pytest.raises(AssertionError):
        sb.sample_pt(sb.Gaussian(), ens, lp[:-1].copy(), dev, 2.0, sb.UpdateFlagSpec.All(), [1.0])
    with pytest.raises(sb.McmcErr):
        sb.swap_walkers(ens, lp, dev, [1.0] * 5)
    sb.swap_walkers(ens, lp[:-1].copy(), dev, [1.0, 0.5])  # silent no-op (utils.rs:88), no device needed
    with pytest.raises(TypeError):
        sb.sample_pt(sb.Gaussian(), ens.astype(np.float32), lp, dev, 2.0, sb.UpdateFlagSpec.All(), [1.0])


def test_product_never_imports_the_oracle():
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|liboracle|scorus_oracle\.h|pyoracle", re.M)
    offenders = []
    for base, _, files in os.walk(os.path.join(ROOT, "scorus_b200")):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                text = open(os.path.join(base, fn), errors="ignore").read()
                if pat.search(text):
                    offenders.append(os.path.join(base, fn))
    for fn in ("scorus_b200.h", "scorus_b200.hpp"):
        if pat.search(open(os.path.join(ROOT, "include", fn)).read()):
            offenders.append(fn)
    assert not offenders, f"product code references the oracle: {offenders}"


def test_sources_target_sm_100a_only():
    mk = open(os.path.join(ROOT, "scorus_b200", "csrc", "Makefile")).read()
    assert "arch=compute_100a,code=sm_100a" in mk and "-lineinfo" in mk and "-fmad=false" in mk
    assert "triton" not in mk.lower()
