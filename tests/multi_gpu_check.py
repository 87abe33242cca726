"""Multi-GPU parity check, run under torchrun (one rank per GPU):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
      tests/multi_gpu_check.py

Every rank also runs the whole ensemble on its own GPU through the single-GPU path and checks that
its shard of the sharded run is identical (positions bit for bit, log-probs to 1e-12): sharding must
not change the chain."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scorus_b200 as sb  # noqa: E402
from scorus_b200.distributed import ShardedSampler  # noqa: E402


def run_case(name, flogprob, n, dim, nbeta, ufs, steps, swap_every, box, matching="global", p2p_swap=True, twalk=False):
    rank, world = dist.get_rank(), dist.get_world_size()
    rng = np.random.default_rng(1234)
    ens = np.ascontiguousarray(rng.uniform(box[0], box[1], (n, dim)))
    beta = 2.0 ** -(np.arange(nbeta) / max(1, nbeta // 8))
    seed = 4242
    # single-GPU reference of the WHOLE ensemble
    ref = sb.Sampler(flogprob, n, dim, seed=seed, device=torch.cuda.current_device())
    ref.upload(ens, None)
    sh = ShardedSampler(flogprob, n, dim, nbeta=nbeta, seed=seed, matching=matching, p2p_swap=p2p_swap)
    lo, hi = sh.w_off, sh.w_off + sh.n_local
    sh.upload(np.ascontiguousarray(ens[lo:hi]), None)
    for k in range(steps):
        if swap_every and k % swap_every == 0:
            ref.swap_walkers(beta)
            sh.swap_walkers(beta)
        if twalk:   # the t-walk move family over the same shards (ufs unused)
            ref.twalk_sample(sb.TWalkParams.new(dim), beta, 1)
            sh.twalk_sample(sb.TWalkParams.new(dim), beta, 1)
        else:
            ref.sample_pt(2.0, ufs, beta, 1)
            sh.sample_pt(2.0, ufs, beta, 1)
    want_e, want_lp = ref.download()
    got_e, got_lp = sh.download()
    ok_pos = np.array_equal(got_e, want_e[lo:hi])
    ok_lp = np.allclose(got_lp, want_lp[lo:hi], rtol=1e-12, atol=1e-9, equal_nan=True)
    moved = not np.array_equal(want_e, ens)
    flag = torch.tensor([int(ok_pos and ok_lp and moved)], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"{name}: mode={sh.mode}/{matching}{'/p2p' if sh.p2p else ''} world={world} n={n} dim={dim} nbeta={nbeta} "
              f"{'OK' if flag.item() else 'MISMATCH'} (pos={ok_pos} lp={ok_lp} moved={moved})", flush=True)
    ref.close()
    return bool(flag.item())


def run_hmc(world):
    """Ensemble-PT HMC over rung shards: identical to the single-GPU chain, adapted step sizes included."""
    rank = dist.get_rank()
    nbeta, npb, dim, l = 2 * world, 512, 8, 4
    n = nbeta * npb
    rng = np.random.default_rng(55)
    q = np.ascontiguousarray(rng.uniform(0.9, 1.1, (n, dim)))
    beta = 2.0 ** -(np.arange(nbeta) / 2.0)
    param = sb.HmcParam.quick_adj(0.7)
    ref = sb.Sampler(sb.Rosenbrock(), n, dim, seed=31, device=torch.cuda.current_device())
    ref.upload(q, None)
    sh = ShardedSampler(sb.Rosenbrock(), n, dim, nbeta=nbeta, seed=31)
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
    flag = torch.tensor([int(ok)], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"hmc_rosenbrock_pt: mode={sh.mode} world={world} n={n} dim={dim} nbeta={nbeta} {'OK' if flag.item() else 'MISMATCH'}", flush=True)
    ref.close()
    return bool(flag.item())


def run_mixed(world):
    """Whole-ensemble matching every 5th step, shard-local in between: a different (still valid) kernel,
    so it is checked statistically: unit Gaussian in, unit Gaussian out, across all shards."""
    n, dim = 8192 * world, 8
    rng = np.random.default_rng(7)
    ens = np.ascontiguousarray(rng.normal(size=(n, dim)) * 3.0)       # over-dispersed start
    sh = ShardedSampler(sb.Gaussian(0.5), n, dim, nbeta=1, seed=99, matching="p2p", global_every=5)
    sh.upload(np.ascontiguousarray(ens[sh.w_off:sh.w_off + sh.n_local]), None)
    sh.sample_pt(2.0, sb.UpdateFlagSpec.Prob(0.5), [1.0], 300)
    e, lp = sh.download()
    stat = torch.tensor([e.sum(), (e ** 2).sum(), float(e.size)], device="cuda", dtype=torch.float64)
    dist.all_reduce(stat)
    mean, var = (stat[0] / stat[2]).item(), (stat[1] / stat[2]).item()
    good = abs(mean) < 0.02 and abs(var - 1.0) < 0.03 and np.allclose(lp, -0.5 * (e ** 2).sum(axis=1), rtol=1e-12)
    if dist.get_rank() == 0:
        print(f"mixed_p2p_every5: world={world} mean={mean:.4f} var={var:.4f} {'OK' if good else 'MISMATCH'}", flush=True)
    return good


def main():
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    world = dist.get_world_size()
    ok = True
    # tempered, whole rungs per rank: no per-step communication, cross-rank swaps (small + large rungs)
    ok &= run_case("rastrigin_pt_small_rungs", sb.Rastrigin(), 16 * 8 * world, 10, 8 * world,
                   sb.UpdateFlagSpec.Prob(0.2), 25, 5, (-0.5, 0.5))
    ok &= run_case("rastrigin_pt_large_rungs", sb.Rastrigin(), 2048 * 4 * world, 10, 4 * world,
                   sb.UpdateFlagSpec.Prob(0.2), 12, 4, (-0.5, 0.5))
    ok &= run_case("gauss_onehot_pt", sb.Gaussian(1.0), 40 * 2 * world, 10, 2 * world,
                   sb.UpdateFlagSpec.OneHotCycle(), 12, 3, (-1.0, 1.0))
    # the same swaps with the NCCL all-to-all row exchange instead of one-sided reads
    ok &= run_case("rastrigin_pt_nccl_exchange", sb.Rastrigin(), 2048 * 4 * world, 10, 4 * world,
                   sb.UpdateFlagSpec.Prob(0.2), 12, 4, (-0.5, 0.5), p2p_swap=False)
    # single temperature, walker shards, the reference's whole-rung matching over an all-gather
    ok &= run_case("rosenbrock_global_matching", sb.Rosenbrock(), 8192 * world, 64, 1,
                   sb.UpdateFlagSpec.Prob(0.5), 10, 0, (0.9, 1.1))
    ok &= run_case("mixture_global_matching_small", sb.Mixture(1.0, 2.0, [5.0] + [0.0] * 5), 64 * world, 6, 1,
                   sb.UpdateFlagSpec.All(), 10, 0, (-6.0, 6.0))
    # the same matching with one-sided NVLink reads of the partners instead of the all-gather
    ok &= run_case("rosenbrock_p2p_matching", sb.Rosenbrock(), 8192 * world, 64, 1,
                   sb.UpdateFlagSpec.Prob(0.5), 10, 0, (0.9, 1.1), matching="p2p")
    ok &= run_case("rastrigin_p2p_matching_small_odd", sb.Rastrigin(), 70 * world, 7, 1,
                   sb.UpdateFlagSpec.Prob(0.2), 10, 0, (-0.5, 0.5), matching="p2p")
    # t-walk over rung shards: no per-step communication, cross-rank ladder swaps
    ok &= run_case("twalk_rastrigin_pt", sb.Rastrigin(), 2048 * 4 * world, 10, 4 * world, None, 12, 4, (-0.5, 0.5), twalk=True)
    ok &= run_case("twalk_rosenbrock_pt_small_rungs", sb.Rosenbrock(), 64 * 2 * world, 20, 2 * world, None, 12, 3, (0.9, 1.1),
                   twalk=True)
    ok &= run_hmc(world)
    ok &= run_mixed(world)
    dist.barrier()
    dist.destroy_process_group()
    if not ok:
        sys.exit(1)
    if int(os.environ.get("RANK", "0")) == 0:
        print("multi_gpu_check: all cases identical to the single-GPU chain")


if __name__ == "__main__":
    main()
