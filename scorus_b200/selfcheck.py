"""Sharded runs against the single-GPU chain, case by case (one process per GPU, torch.distributed initialised).

Every rank also runs the WHOLE ensemble on its own GPU through the single-GPU path and checks that its shard of the
sharded run is identical: positions bit for bit, log-probs to 1e-12.  Sharding must not change the chain:
streams are keyed by global walker / rung / block ids, so this holds for every mode -- rung shards, walker shards with
the whole-ensemble matching (one-sided or all-gather), and the k-schedule / shard-local matchings, whose single-GPU
counterpart is the blocked matching of scorus_set_matching_blocks.

Used by tests/multi_gpu_check.py (all cases) and by `bench.py --gpus N` (the quick subset, before timing; the result
goes into the JSON line as "parity_check").  Compares the product with itself on two layouts; no oracle involved.
"""
from __future__ import annotations

import numpy as np

from .ensemble_sample import Gaussian, Mixture, Rastrigin, Rosenbrock, Sampler, UpdateFlagSpec
from .hmc import HmcParam
from .twalk import TWalkParams


def _all_ok(torch, dist, ok: bool) -> bool:
    flag = torch.tensor([int(ok)], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    return bool(flag.item())


def _cycling_flags(dim):
    """A stateful Func closure (in the spirit of examples/sample_high_dim.rs:59-65): two flags, advancing per call."""
    state = {"i": 0}

    def f():
        v = np.zeros(dim, dtype=bool)
        v[state["i"] % dim] = True
        v[(state["i"] * 3 + 1) % dim] = True
        state["i"] += 1
        return v
    return f


def run_case(name, flogprob, n, dim, nbeta, ufs, steps, swap_every, box, matching="global", p2p_swap=True, twalk=False,
             global_every=1, steps_per_call=1, log=print):
    """One sharded configuration against the single-GPU chain of the whole ensemble.  -> identical (bool)."""
    import torch
    import torch.distributed as dist
    from .distributed import ShardedSampler
    rank, world = dist.get_rank(), dist.get_world_size()
    rng = np.random.default_rng(1234)
    ens = np.ascontiguousarray(rng.uniform(box[0], box[1], (n, dim)))
    beta = 2.0 ** -(np.arange(nbeta) / max(1, nbeta // 8))
    seed = 4242
    ref = Sampler(flogprob, n, dim, seed=seed, device=torch.cuda.current_device())
    ref.upload(ens, None)
    sh = ShardedSampler(flogprob, n, dim, nbeta=nbeta, seed=seed, matching=matching, p2p_swap=p2p_swap,
                        global_every=global_every)
    if sh.mode == "walkers" and (matching == "local" or global_every > 1):
        # the single-GPU statement of "pair inside the shards except on every k-th step"
        ref.set_matching_blocks(world, 0, 0 if matching == "local" else global_every)
    lo, hi = sh.w_off, sh.w_off + sh.n_local
    sh.upload(np.ascontiguousarray(ens[lo:hi]), None)
    ufs_ref = ufs_sh = ufs
    if isinstance(ufs, str) and ufs == "func":   # two copies of one stateful closure: the reference's and the sharded run's
        ufs_ref, ufs_sh = UpdateFlagSpec.Func(_cycling_flags(dim)), UpdateFlagSpec.Func(_cycling_flags(dim))
    k = 0
    while k < steps:
        if swap_every and k % swap_every == 0:
            ref.swap_walkers(beta)
            sh.swap_walkers(beta)
        run = min(steps_per_call, steps - k)
        if swap_every:
            run = min(run, swap_every - k % swap_every)
        if twalk:   # the t-walk move family over the same shards (ufs unused)
            ref.twalk_sample(TWalkParams.new(dim), beta, run)
            sh.twalk_sample(TWalkParams.new(dim), beta, run)
        else:
            ref.sample_pt(2.0, ufs_ref, beta, run)
            sh.sample_pt(2.0, ufs_sh, beta, run)
        k += run
    want_e, want_lp = ref.download()
    got_e, got_lp = sh.download()
    ok_pos = np.array_equal(got_e, want_e[lo:hi])
    ok_lp = np.allclose(got_lp, want_lp[lo:hi], rtol=1e-12, atol=1e-9, equal_nan=True)
    moved = not np.array_equal(want_e, ens)
    a_ref, a_sh = ref.stats()["accepted"], torch.tensor([sh.stats()["accepted"]], device="cuda", dtype=torch.int64)
    dist.all_reduce(a_sh)
    ok_cnt = twalk or int(a_sh.item()) == a_ref      # accept counters add up over the ranks
    ok = _all_ok(torch, dist, ok_pos and ok_lp and moved and ok_cnt)
    if rank == 0:
        every = f" every {global_every}" if global_every > 1 else ""
        log(f"{name}: mode={sh.mode}/{matching}{every}{'/one-sided' if sh.p2p else ''} world={world} n={n} dim={dim} "
            f"nbeta={nbeta} steps={steps} {'OK' if ok else 'MISMATCH'} (pos={ok_pos} lp={ok_lp} moved={moved} counts={ok_cnt})")
    ref.close()
    sh.close()
    return ok


def run_host_call(world, log=print):
    """The reference-shaped call on pinned host buffers, per shard (ShardedSampler.sample_pt_host), on the schedule
    'whole-ensemble matching on every 2nd step': the host buffers must hold the single-GPU chain after every call --
    i.e. every accepted walker of the shard travelled back, whichever rank computed it."""
    import torch
    import torch.distributed as dist
    from .distributed import ShardedSampler
    from .ensemble_sample import pinned_empty
    rank = dist.get_rank()
    n, dim, seed = 4096 * world, 64, 77
    rng = np.random.default_rng(9)
    ens = np.ascontiguousarray(rng.uniform(0.9, 1.1, (n, dim)))
    ref = Sampler(Rosenbrock(), n, dim, seed=seed, device=torch.cuda.current_device())
    ref.set_matching_blocks(world, 0, 2)
    ref.upload(ens, None)
    _, lp0 = ref.download()
    sh = ShardedSampler(Rosenbrock(), n, dim, nbeta=1, seed=seed, matching="p2p", global_every=2)
    lo, hi = sh.w_off, sh.w_off + sh.n_local
    h_ens, h_lp = pinned_empty((sh.n_local, dim)), pinned_empty((sh.n_local,))
    h_ens[:] = ens[lo:hi]
    h_lp[:] = lp0[lo:hi]
    ok = True
    for k in range(6):
        ref.sample_pt(2.0, UpdateFlagSpec.Prob(0.5), [1.0], 1)
        sh.sample_pt_host(h_ens, h_lp, 2.0, UpdateFlagSpec.Prob(0.5), [1.0])
        want_e, want_lp = ref.download()
        ok = ok and np.array_equal(h_ens, want_e[lo:hi]) and np.allclose(h_lp, want_lp[lo:hi], rtol=1e-12, atol=1e-9)
    ok = _all_ok(torch, dist, ok and not np.array_equal(want_e, ens))
    if rank == 0:
        log(f"host_call_one_sided_every2: world={world} n={n} dim={dim} {'OK' if ok else 'MISMATCH'}")
    ref.close()
    sh.close()
    return ok


def run_hmc(world, log=print):
    """Ensemble-PT HMC over rung shards: identical to the single-GPU chain, adapted step sizes included."""
    import torch
    import torch.distributed as dist
    from .distributed import ShardedSampler
    rank = dist.get_rank()
    nbeta, npb, dim, l = 2 * world, 512, 8, 4
    n = nbeta * npb
    rng = np.random.default_rng(55)
    q = np.ascontiguousarray(rng.uniform(0.9, 1.1, (n, dim)))
    beta = 2.0 ** -(np.arange(nbeta) / 2.0)
    param = HmcParam.quick_adj(0.7)
    ref = Sampler(Rosenbrock(), n, dim, seed=31, device=torch.cuda.current_device())
    ref.upload(q, None)
    sh = ShardedSampler(Rosenbrock(), n, dim, nbeta=nbeta, seed=31)
    lo, hi = sh.w_off, sh.w_off + sh.n_local
    sh.upload(np.ascontiguousarray(q[lo:hi]), None)
    eps_ref, eps_sh = np.full(nbeta, 0.003), np.full(nbeta, 0.003)
    for k in range(8):
        if k % 4 == 0:
            ref.swap_walkers(beta)
            sh.swap_walkers(beta)
        ref.hmc_sample_ensemble_pt(eps_ref, beta, l, param, 1)
        sh.hmc_sample_ensemble_pt(eps_sh, beta, l, param, 1)
    want_q, want_lp = ref.download()
    got_q, got_lp = sh.download()
    r0, r1 = sh.rung_off, sh.rung_off + sh.nb_local
    ok = (np.array_equal(got_q, want_q[lo:hi]) and np.allclose(got_lp, want_lp[lo:hi], rtol=1e-12, atol=1e-9)
          and np.array_equal(eps_sh[r0:r1], eps_ref[r0:r1]) and not np.array_equal(want_q, q))
    ok = _all_ok(torch, dist, ok)
    if rank == 0:
        log(f"hmc_rosenbrock_pt: mode={sh.mode} world={world} n={n} dim={dim} nbeta={nbeta} {'OK' if ok else 'MISMATCH'}")
    ref.close()
    sh.close()
    return ok


def run_hmc_walker_shards(world, log=print):
    """Adaptive-step-size HMC over walker shards of ONE rung: positions, log-probs and the adapted epsilon identical to the
    single-GPU chain (whose rung, above 4096 walkers, adapts by the same net count)."""
    import torch
    import torch.distributed as dist
    from .distributed import ShardedSampler
    rank = dist.get_rank()
    n, dim, l = 4096 * world, 8, 3
    rng = np.random.default_rng(56)
    q = np.ascontiguousarray(rng.uniform(0.9, 1.1, (n, dim)))
    param = HmcParam.quick_adj(0.7)
    ref = Sampler(Rosenbrock(), n, dim, seed=33, device=torch.cuda.current_device())
    ref.upload(q, None)
    sh = ShardedSampler(Rosenbrock(), n, dim, nbeta=1, seed=33, matching="local")
    lo, hi = sh.w_off, sh.w_off + sh.n_local
    sh.upload(np.ascontiguousarray(q[lo:hi]), None)
    eps_ref, eps_sh = np.full(1, 0.003), np.full(1, 0.003)
    cnt_ref = ref.hmc_sample_ensemble_pt(eps_ref, [1.0], l, param, 5)
    cnt_sh = torch.tensor([int(sh.hmc_sample_ensemble_pt(eps_sh, [1.0], l, param, 5)[0])], device="cuda", dtype=torch.int64)
    dist.all_reduce(cnt_sh)
    want_q, want_lp = ref.download()
    got_q, got_lp = sh.download()
    ok = (np.array_equal(got_q, want_q[lo:hi]) and np.allclose(got_lp, want_lp[lo:hi], rtol=1e-12, atol=1e-9)
          and eps_sh[0] == eps_ref[0] and eps_ref[0] != 0.003 and int(cnt_sh.item()) == int(cnt_ref[0]))
    ok = _all_ok(torch, dist, ok)
    if rank == 0:
        log(f"hmc_rosenbrock_walker_shards_adaptive: world={world} n={n} dim={dim} eps {eps_ref[0]:.6g} / {eps_sh[0]:.6g} "
            f"{'OK' if ok else 'MISMATCH'}")
    ref.close()
    sh.close()
    return ok


def run_mixed_statistics(world, log=print):
    """Whole-ensemble matching every 5th step, shard-local in between, also checked for what it samples: unit
    Gaussian in, unit Gaussian out, across all shards."""
    import torch
    import torch.distributed as dist
    from .distributed import ShardedSampler
    n, dim = 8192 * world, 8
    rng = np.random.default_rng(7)
    ens = np.ascontiguousarray(rng.normal(size=(n, dim)) * 3.0)       # over-dispersed start
    sh = ShardedSampler(Gaussian(0.5), n, dim, nbeta=1, seed=99, matching="p2p", global_every=5)
    sh.upload(np.ascontiguousarray(ens[sh.w_off:sh.w_off + sh.n_local]), None)
    sh.sample_pt(2.0, UpdateFlagSpec.Prob(0.5), [1.0], 300)
    e, lp = sh.download()
    stat = torch.tensor([e.sum(), (e ** 2).sum(), float(e.size)], device="cuda", dtype=torch.float64)
    dist.all_reduce(stat)
    mean, var = (stat[0] / stat[2]).item(), (stat[1] / stat[2]).item()
    good = abs(mean) < 0.02 and abs(var - 1.0) < 0.03 and np.allclose(lp, -0.5 * (e ** 2).sum(axis=1), rtol=1e-12)
    good = _all_ok(torch, dist, good)
    if dist.get_rank() == 0:
        log(f"mixed_one_sided_every5_statistics: world={world} mean={mean:.4f} var={var:.4f} {'OK' if good else 'MISMATCH'}")
    sh.close()
    return good


def sharded_parity_cases(quick: bool = False, log=print):
    """Runs the cases; -> (number of cases, all identical).  quick: the subset bench.py runs before timing (the modes
    its lines are measured in, a few seconds)."""
    import torch.distributed as dist
    world = dist.get_world_size()
    P, A = UpdateFlagSpec.Prob, UpdateFlagSpec.All
    res = []
    # single temperature, walker shards, the reference's whole-rung matching, one-sided over NVLink
    res.append(run_case("rosenbrock_one_sided", Rosenbrock(), 8192 * world, 64, 1, P(0.5), 10, 0, (0.9, 1.1), matching="p2p",
                        log=log))
    res.append(run_case("rosenbrock_one_sided_multi_step_call", Rosenbrock(), 4096 * world, 64, 1, P(0.5), 12, 0, (0.9, 1.1),
                        matching="p2p", steps_per_call=5, log=log))
    # the k-schedule: whole-ensemble matching on every 4th step, shard-local (blocked) matching in between
    res.append(run_case("rosenbrock_one_sided_every4", Rosenbrock(), 8192 * world, 64, 1, P(0.5), 13, 0, (0.9, 1.1),
                        matching="p2p", global_every=4, steps_per_call=13, log=log))
    # tempered, whole rungs per rank: no per-step communication, cross-rank swaps
    res.append(run_case("rastrigin_pt_large_rungs", Rastrigin(), 2048 * 4 * world, 10, 4 * world, P(0.2), 12, 4, (-0.5, 0.5),
                        log=log))
    if quick:
        return len(res), all(res)
    res.append(run_case("mixture32_one_sided", Mixture(1.0, 2.0, [5.0] + [0.0] * 31), 4096 * world, 32, 1, P(0.5), 8, 0,
                        (-6.0, 6.0), matching="p2p", log=log))
    res.append(run_case("rastrigin_one_sided_small_odd", Rastrigin(), 70 * world, 7, 1, P(0.2), 10, 0, (-0.5, 0.5),
                        matching="p2p", log=log))
    res.append(run_case("gauss_one_sided_every3_small_blocks", Gaussian(0.5), 64 * world, 6, 1, P(0.5), 10, 0, (-1.0, 1.0),
                        matching="p2p", global_every=3, log=log))
    res.append(run_case("rosenbrock_shard_local", Rosenbrock(), 8192 * world, 64, 1, P(0.5), 6, 0, (0.9, 1.1),
                        matching="local", steps_per_call=3, log=log))
    # a tempered ensemble whose three rungs do not divide over the ranks: walker shards cut through rungs; every step and
    # every ladder swap is one-sided
    res.append(run_case("rastrigin_pt_walker_shards", Rastrigin(), 3 * 2048 * world, 10, 3, P(0.2), 12, 4, (-0.5, 0.5),
                        matching="p2p", log=log))
    # the same matching over an NCCL all-gather of the ensemble (BASELINE.json's wording), also on the k-schedule
    res.append(run_case("rosenbrock_allgather", Rosenbrock(), 8192 * world, 64, 1, P(0.5), 10, 0, (0.9, 1.1), log=log))
    res.append(run_case("mixture_allgather_small", Mixture(1.0, 2.0, [5.0] + [0.0] * 5), 64 * world, 6, 1, A(), 10, 0,
                        (-6.0, 6.0), log=log))
    res.append(run_case("rosenbrock_allgather_every3", Rosenbrock(), 4096 * world, 64, 1, P(0.5), 8, 0, (0.9, 1.1),
                        global_every=3, log=log))
    # tempered: small rungs (exact shuffles), one-hot flags, the NCCL all-to-all row exchange
    res.append(run_case("rastrigin_pt_small_rungs", Rastrigin(), 16 * 8 * world, 10, 8 * world, P(0.2), 25, 5, (-0.5, 0.5),
                        log=log))
    res.append(run_case("gauss_onehot_pt", Gaussian(1.0), 40 * 2 * world, 10, 2 * world, UpdateFlagSpec.OneHotCycle(), 12, 3,
                        (-1.0, 1.0), log=log))
    res.append(run_case("rastrigin_pt_nccl_exchange", Rastrigin(), 2048 * 4 * world, 10, 4 * world, P(0.2), 12, 4,
                        (-0.5, 0.5), p2p_swap=False, log=log))
    # t-walk over rung shards: no per-step communication, cross-rank ladder swaps
    res.append(run_case("twalk_rastrigin_pt", Rastrigin(), 2048 * 4 * world, 10, 4 * world, None, 12, 4, (-0.5, 0.5),
                        twalk=True, log=log))
    res.append(run_case("twalk_rosenbrock_pt_small_rungs", Rosenbrock(), 64 * 2 * world, 20, 2 * world, None, 12, 3,
                        (0.9, 1.1), twalk=True, log=log))
    # UpdateFlagSpec::Func over shards: every rank makes the closure's n_global calls; one-sided, k-schedule, all-gather
    res.append(run_case("rosenbrock_func_flags_one_sided_every2", Rosenbrock(), 512 * world, 16, 1, "func", 6, 0, (0.9, 1.1),
                        matching="p2p", global_every=2, log=log))
    res.append(run_case("gauss_func_flags_allgather", Gaussian(0.5), 64 * world, 6, 1, "func", 4, 0, (-1.0, 1.0), log=log))
    # t-walk with the whole-rung matching across walker shards (replicated compute)
    res.append(run_case("twalk_rosenbrock_walker_shards_whole_matching", Rosenbrock(), 2048 * world, 20, 1, None, 6, 0,
                        (0.9, 1.1), matching="p2p", twalk=True, steps_per_call=2, log=log))
    res.append(run_hmc(world, log=log))
    res.append(run_hmc_walker_shards(world, log=log))
    res.append(run_host_call(world, log=log))
    res.append(run_mixed_statistics(world, log=log))
    return len(res), all(res)
