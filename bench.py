#!/usr/bin/env python
"""bench.py -- walker-steps/s of the ensemble-sampler hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c5] [--impl ours|reference]

A "step" is one fused sample_pt pass over the whole ensemble (plus swap_walkers before every 10th
step for the tempered workload).  Default workload = BASELINE config 5, the one the metric and
the 70 %-of-HBM target are quoted on: 64-D Rosenbrock, 2^22 walkers, a = 2, Prob(0.5) flags as in
examples/test_emcee.rs:55-62, fp64.  One JSON line on stdout (rank 0); see DESIGN.md section 7.

  value        device-timed (CUDA events on the launch stream), ensemble resident in HBM
  e2e          same metric through the reference-shaped host call (scorus_sample_pt_host):
               pinned host buffers in, one step, host buffers out, copies inside the timed region
  roofline     algorithmic bytes (16*D+16 per walker-step) / measured launch time
               against MEASURED_PEAKS.json's HBM copy bandwidth
  cpu_baseline the oracle's reference-structured CPU port on this box's cores, bounded sample
  --impl reference   times only that CPU port (the reference is Rust; no toolchain in the image)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
ELEM = 8   # bytes per coordinate of the run (--dtype)

WORKLOADS = {
    # name: (description, kind, n, dim, nbeta, ufs, prob, init box, swap_every)
    "c1": ("c1_test_emcee: 2-D Gaussian, 16 walkers", 0, 16, 2, 1, "prob", 0.5, (0.9, 1.1), 0),
    "c2": ("c2_sample_high_dim: 100-D AR(1) rho=0.9 correlated Gaussian, 65536 walkers", 6, 65536, 100, 1, "all", 0.0, (-1.0, 1.0), 0),
    "c3": ("c3_sample_bimodal: 32-D two-component mixture, 2^20 walkers", 7, 1 << 20, 32, 1, "prob", 0.5, (-6.0, 6.0), 0),
    "c4": ("c4_sample_rastrigin: 10-D Rastrigin, 64 rungs x 16384 walkers, swap before every 10th step", 3, 1 << 20, 10, 64, "prob", 0.2, (-0.01, 0.01), 10),
    "c5": ("c5_scale_sweep: 64-D Rosenbrock, 2^22 walkers", 2, 1 << 22, 64, 1, "prob", 0.5, (0.9, 1.1), 0),
}


def workload_params(kind, dim):
    if kind == 0:
        return [0.5]
    if kind == 3:
        return [10.0]
    if kind == 6:  # AR(1), Sigma_ij = rho^|i-j|: L upper bidiagonal with L^T L = Sigma^-1
        rho = 0.9
        L = np.zeros((dim, dim))
        s = 1.0 / np.sqrt(1.0 - rho * rho)
        for i in range(dim - 1):
            L[i, i] = s
            L[i, i + 1] = -rho * s
        L[dim - 1, dim - 1] = 1.0
        return np.concatenate([np.zeros(dim), L.ravel()]).tolist()
    if kind == 7:
        m = np.zeros(dim)
        m[0] = 5.0
        return [1.0, 2.0] + m.tolist()
    return []


def betas(nbeta):
    # 8 rungs: the reference's 2^-i ladder (examples/sample_rastrigin.rs:29); 64 rungs: 2^(-i/8)
    k = max(1, nbeta // 8)
    return (2.0 ** -(np.arange(nbeta) / k)).astype(np.float64)


def b_alg(dim, elem=8):
    return 2 * elem * dim + 2 * elem


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (pynvml, else nvidia-smi), sampled by the main thread
    while the GPU works through the queued steps -- after the last launch has been enqueued, before the synchronise.
    Round 1 polled NVML from a second thread every 5 ms; with one process per GPU on eight GPUs those queries took
    tens of milliseconds each and stalled the launches of the timed region (the same 100 steps: 0.136 ms/step with the
    thread off, 0.546 with it on, profiles/r2_notes.md), and initialising NVML right before the first timed launch cost
    another 2 ms.  NVML is therefore initialised at start-up and never queried while launches are being enqueued."""
    NAMES = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
             "hw_power_brake": 0x80, "sync_boost": 0x10, "display_clocks": 0x100}

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._nvml = None
        if os.environ.get("BENCH_CLOCKS", "1") == "0":
            return
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nvml = None

    def sample(self):
        nv = self._nvml
        if nv is None:
            return
        try:
            self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
            try:
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
            except Exception:
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
            for k, bit in self.NAMES.items():
                if r & bit:
                    self.reasons.add(k)
        except Exception:
            pass

    def sample_until(self, event, period=0.002, limit=200):
        """Call after the timed steps have been enqueued: samples until `event` (recorded after them) has completed."""
        self.sample()
        while not event.query() and len(self.samples) < limit:
            time.sleep(period)
            self.sample()

    def stop(self):
        if not self.samples:  # fallback: one nvidia-smi query
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", "--query-gpu=clocks.sm,clocks.max.sm",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=10).stdout
                a, b = out.strip().split(",")
                self.samples, self.max_mhz = [int(a)], int(b)
            except Exception:
                pass
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def make_ensemble(n, dim, box, seed, kind):
    rng = np.random.default_rng(seed)
    ens = rng.random((n, dim))
    ens *= (box[1] - box[0])
    ens += box[0]
    if kind == 7:  # start near the +m mode
        ens[:, 0] = np.abs(ens[:, 0])
    return ens


def cpu_reference_leg(wl, steps, warmup, sample_walkers=None, target_seconds=12.0):
    """The reference's CPU path, restated (oracle/ref_structure_bench.c), on all host threads, on a
    bounded sample of the workload: the first `sample_walkers` walkers (whole rungs)."""
    from oracle import pyoracle as po
    po.build()
    desc, kind, n, dim, nbeta, ufs, prob, box, swap_every = WORKLOADS[wl]
    params = workload_params(kind, dim)
    if sample_walkers is None:
        sample_walkers = min(n, int(os.environ.get("SCORUS_BENCH_CPU_WALKERS", 1 << 17)))
    npb = n // nbeta
    if nbeta > 1:
        rungs = max(1, min(nbeta, sample_walkers // npb))
        n_s, nb_s = rungs * npb, rungs
    else:
        n_s, nb_s = sample_walkers, 1
    beta = betas(nbeta)[:nb_s]
    ens = make_ensemble(n_s, dim, box, 5, kind)
    lp = po.init_logprob(kind, params, ens)
    ukind = {"prob": po.UFS_PROB, "all": po.UFS_ALL, "onehot": po.UFS_ONE_HOT_CYCLE}[ufs]
    cores = po.max_threads()
    # warm-up + calibration: one step, then as many as fit the time budget (bounded by `steps`)
    t1, _ = po.ref_structure_run(kind, params, ens, lp, beta, 2.0, ukind, prob, 11, max(1, min(warmup, 1)), cores)
    k = int(max(1, min(steps, target_seconds / max(t1, 1e-6))))
    t, stage = po.ref_structure_run(kind, params, ens, lp, beta, 2.0, ukind, prob, 12, k, cores)
    value = n_s * k / t
    names = ["pairing", "z", "flags", "proposal", "logprob", "accept"]
    # the best-effort flat CPU variant: contiguous buffers, every stage threaded, the device's counter-based
    # streams -- so that the GPU is not only compared with an allocation-bound structure
    flat = None
    if ufs != "onehot":
        e2, l2 = ens.copy(), lp.copy()
        tf1 = po.flat_threaded_run(kind, params, e2, l2, beta, 2.0, ukind, prob, 13, 0, 1, cores)
        kf = int(max(1, min(steps, 0.5 * target_seconds / max(tf1, 1e-6))))
        tf = po.flat_threaded_run(kind, params, e2, l2, beta, 2.0, ukind, prob, 13, 1, kf, cores)
        if tf > 0:
            flat = {"value": n_s * kf / tf, "unit": "walker-steps/s", "cores": cores, "steps": kf,
                    "note": "flat buffers, all stages threaded including the counter-based stream generation "
                            "(oracle/ref_structure_bench.c: orc_flat_threaded_run): what a CPU rewrite could do, not what the "
                            "reference does"}
    return {
        "value": value, "unit": "walker-steps/s", "cores": cores, "kind": "port",
        "sample": f"{k} steps of the first {n_s} walkers ({nb_s} rung(s)) of {desc}; reference-structured C port "
                  f"(per-walker heap vectors, serial stages, log-density on {cores} threads)",
        "seconds": t, "steps": k, "n_walkers": n_s,
        "stage_share": {nm: float(s / stage.sum()) for nm, s in zip(names, stage)},
        "flat_threaded": flat,
    }


def cpu_move_leg(wl, move, target_seconds=8.0):
    """CPU port of the other two move families (oracle/twalk.c, oracle/hmc.c: serial restatements, one core) on a
    bounded sample of the workload, for the GPU-to-CPU ratio of those paths."""
    from oracle import pyoracle as po
    po.build()
    desc, kind, n, dim, nbeta, ufs, prob, box, swap_every = WORKLOADS[wl]
    params = workload_params(kind, dim)
    npb = n // nbeta
    nb_s = max(1, min(nbeta, (1 << 14) // npb)) if nbeta > 1 else 1
    n_s = nb_s * npb if nbeta > 1 else min(n, 1 << 14)
    beta = betas(nbeta)[:nb_s]
    ens = make_ensemble(n_s, dim, box, 5, kind)
    lp = po.init_logprob(kind, params, ens)
    rng = po.Rng(3)
    tw = po.twalk_params(dim)
    eps = np.full(nb_s, {2: 0.002, 3: 0.02}.get(kind, 0.1))

    def one():
        if move == "twalk":
            po.twalk_sample(kind, params, ens, lp, rng, tw, beta)
        else:
            po.hmc_ensemble_pt(kind, params, ens, lp, rng, eps, beta, 10, 0.8, 0.0)

    t0 = time.perf_counter()
    one()
    t1 = time.perf_counter() - t0
    k = int(max(1, min(50, target_seconds / max(t1, 1e-6))))
    t0 = time.perf_counter()
    for _ in range(k):
        one()
    t = time.perf_counter() - t0
    return {"value": n_s * k / t, "unit": "walker-steps/s", "cores": 1, "kind": "port",
            "sample": f"{k} steps of the first {n_s} walkers ({nb_s} rung(s)) of {desc}; serial C restatement "
                      f"(oracle/{'twalk' if move == 'twalk' else 'hmc'}.c), one core"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--workload", default="c5", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ufs", default=None, choices=["prob", "all", "onehot"])
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--move", default="stretch", choices=["stretch", "twalk", "hmc"],
                    help="stretch = the headline metric (ensemble_sample::sample_pt); twalk = the t-walk move family "
                         "(twalk::sample, SURVEY 8(f) rank 2), hmc = ensemble-PT HMC (hmc::naive::sample_ensemble_pt, rank 3; "
                         "10 leapfrog steps, fixed step size) on the same workload: extra measurements")
    ap.add_argument("--n-walkers", type=int, default=0, help="experiments only: override the workload's ensemble size")
    ap.add_argument("--global-every", type=int, default=8,
                    help="N>1, --matching p2p/global: the whole-ensemble matching on every k-th step, shard-local (blocked) "
                         "matching in between; 1 = the reference's matching on every step")
    ap.add_argument("--matching", default="p2p", choices=["local", "global", "p2p"],
                    help="N>1, single temperature: p2p = the reference's whole-ensemble matching, one-sided over NVLink "
                         "(pair owner reads / writes the partner's row in the peer's HBM); global = the same over an NCCL "
                         "all-gather of the ensemble per step; local = shard-local matchings only (no exchange: the "
                         "replica upper bound)")
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"],
                    help="f32: the T = f32 instantiation of the step / init / swap kernels (one GPU, device-timed line only; "
                         "roofline bytes 8 D + 8 per walker-step)")
    ap.add_argument("--no-verify", action="store_true", help="N>1: skip the sharded-vs-single-GPU parity cases before timing")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    desc, kind, n, dim, nbeta, ufs, prob, box, swap_every = WORKLOADS[args.workload]
    if args.ufs:
        ufs = args.ufs
    if args.n_walkers:
        n = args.n_walkers
        desc += f" [n_walkers overridden to {n}]"
    W = max(args.warmup, 3)
    global ELEM
    ELEM = 4 if args.dtype == "f32" else 8
    if args.dtype == "f32" and world > 1:
        raise SystemExit("--dtype f32 is a single-GPU line")

    if args.move == "twalk":
        desc += " [move: t-walk, twalk::sample with TWalkParams::new(dim)]"
    if args.move == "hmc":
        desc += " [move: HMC, sample_ensemble_pt with l = 10 leapfrog steps, HmcParam::fixed]"
    if args.impl == "reference":
        if args.move != "stretch":
            raise SystemExit("--impl reference times the headline path (--move stretch)")
        if rank != 0:
            return 0
        r = cpu_reference_leg(args.workload, args.steps, W, target_seconds=60.0)
        line = {
            "impl": "reference", "metric": "walker_steps_per_sec", "value": r["value"], "unit": "walker-steps/s",
            "n_gpus": args.gpus, "steps": r["steps"], "warmup": 1, "ms_per_step": 1e3 * r["seconds"] / r["steps"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "n_walkers": n, "dim": dim, "nbeta": nbeta, "ufs": ufs, "prob": prob, "a": 2.0,
                       "sample_walkers": r["n_walkers"]},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "stage_share", "flat_threaded")},
            "e2e": {"value": r["value"], "unit": "walker-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line))
        return 0

    import torch
    import scorus_b200 as sb

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local_rank)
    clocks = ClockSampler(local_rank)   # NVML initialised here, long before the timed region
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # ---- shard the walkers: whole rungs per rank when tempered, contiguous walkers otherwise ----
    params = workload_params(kind, dim)
    spec = {"prob": sb.UpdateFlagSpec.Prob(prob), "all": sb.UpdateFlagSpec.All(),
            "onehot": sb.UpdateFlagSpec.OneHotCycle()}[ufs]
    stream = torch.cuda.Stream()  # the launch stream; the CUDA events below are recorded on it
    torch.cuda.set_stream(stream)
    beta = betas(nbeta)
    parity_check = None
    if world > 1 and not args.no_verify and args.move == "stretch":
        # sharded chain == single-GPU chain, in the modes the lines below are measured in (scorus_b200/selfcheck.py)
        from scorus_b200.selfcheck import sharded_parity_cases
        log = []
        cases, same = sharded_parity_cases(quick=True, log=log.append)
        parity_check = {"cases": cases, "identical": bool(same), "world": world, "log": log}
        if not same:
            if rank == 0:
                print("\n".join(log), file=sys.stderr)
            raise SystemExit("sharded run differs from the single-GPU chain: not timing it")
    if world > 1:
        from scorus_b200.distributed import ShardedSampler
        s = ShardedSampler(sb.LogProb(kind, tuple(params)), n, dim, nbeta=nbeta, seed=2024, matching=args.matching,
                           global_every=args.global_every)
        n_local, nb_local = s.n_local, s.nb_local
        if s.mode == "rungs":
            parallelism = (f"rungs->gpus x{world}: no per-step exchange; swap = all-gather(lp) + "
                           + ("one-sided NVLink pull" if s.p2p else "all-to-all") + " of crossing rows")
        elif args.matching == "global":
            parallelism = f"walker shards x{world}, whole-ensemble matching, NCCL all-gather of the ensemble per step"
        elif args.matching == "p2p":
            parallelism = (f"walker shards x{world}, cross-shard exchange: whole-ensemble matching, pair owner reads / writes "
                           "the remote partner's row one-sidedly over NVLink, one flag barrier per step")
        else:
            parallelism = f"walker shards x{world}, shard-local matching only (no exchange: replica upper bound)"
        if s.mode == "walkers" and args.matching != "local" and args.global_every > 1:
            parallelism += f"; on every {args.global_every}th step, shard-local (blocked) matching in between"
        ens = make_ensemble(n_local, dim, box, 5 + rank, kind)
        s.upload(ens, None)
        s.sync()
        launch_count = lambda: s.sampler.launch_count
        swap = s.swap_walkers
        can_swap = s.mode == "rungs"
    else:
        n_local, nb_local = n, nbeta
        parallelism = "single gpu"
        ens = make_ensemble(n_local, dim, box, 5, kind)
        if args.dtype == "f32":
            if args.move != "stretch":
                raise SystemExit("--dtype f32 covers the stretch move (K1 / K2 / K5)")
            args.no_e2e, args.no_cpu = True, True      # host-buffer calls and the CPU port are f64
            ens = np.ascontiguousarray(ens.astype(np.float32))
            s = sb.Sampler(sb.LogProb(kind, tuple(params)), n_local, dim, seed=2024, device=local_rank, dtype=np.float32)
        else:
            s = sb.Sampler(sb.LogProb(kind, tuple(params)), n_local, dim, seed=2024, device=local_rank)
        s.set_stream(stream.cuda_stream)
        s.upload(ens, None)
        s.sync()
        launch_count = lambda: s.launch_count
        swap = s.swap_walkers
        can_swap = True

    tw = sb.TWalkParams.new(dim)

    hmc_eps = np.full(nbeta, {2: 0.002, 3: 0.02}.get(kind, 0.1))   # the whole ladder's step sizes (a rank uses its rungs')
    hmc_param = sb.HmcParam.fixed(0.8)

    def do_steps(count):
        if args.move == "hmc":
            s.hmc_sample_ensemble_pt(hmc_eps, beta, 10, hmc_param, count)   # walkers are independent: no exchange on any sharding
        elif args.move == "twalk":
            s.twalk_sample(tw, beta, count)
        else:
            s.sample_pt(2.0, spec, beta, count)

    def run_steps(k0, count):
        # consecutive steps between two swaps go down in ONE library call (nsteps): the launches are
        # queued back to back from C, without a Python round trip per step
        k, end = k0, k0 + count
        while k < end:
            if swap_every and nbeta > 1 and can_swap and k % swap_every == 0:
                swap(beta)
            nxt = min(end, (k // swap_every + 1) * swap_every) if (swap_every and nbeta > 1 and can_swap) else end
            if os.environ.get("BENCH_STEP_PER_CALL"):
                nxt = k + 1
            do_steps(nxt - k)
            k = nxt

    # ---- walker shards: the pieces of the schedule on their own, and the NVLink denominators ----
    extra = None
    if world > 1 and s.mode == "walkers" and args.move == "stretch":
        def timed(fn, k):
            fn(3)    # NVLink and clocks awake before the timed calls (the first exchange steps after an idle spell ran 30 % slow)
            torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream); fn(k); e1.record(stream)
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / k], device="cuda", dtype=torch.float64)
            if os.environ.get("BENCH_DIAG"):
                allt = [torch.zeros_like(t) for _ in range(world)]
                dist.all_gather(allt, t)
                diag.append([round(float(x.item()), 5) for x in allt])
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        diag = []
        k_x = max(8, min(args.steps, 40))
        t_local = timed(lambda k: s.sampler.sample_pt(2.0, spec, beta, k), k_x)
        extra = {"replicas_upper_bound": {"ms_per_step": t_local, "value": n / (t_local * 1e-3),
                                          "note": "shard-local matching on every step: no exchange at all"}}
        if s.p2p:
            t_glob = timed(lambda k: s.sampler.sample_pt_peer(2.0, spec, beta, k), k_x)
            st_x = s.stats()
            acc = st_x["accepted"] / max(1, st_x["proposed"])
            cross = (n_local / 2) * (world - 1) / world                 # owned pairs whose partner lives on a peer
            rd, wr = cross * (dim * 8 + 8), cross * acc * (dim * 8 + 8)
            ce, sm, rr = s.sampler.peer_read_bandwidth((rank + 1) % world, 0, 5)
            bw = torch.tensor([ce, sm, rr], device="cuda", dtype=torch.float64)
            dist.all_reduce(bw, op=dist.ReduceOp.MIN)
            ce, sm, rr = (float(v) for v in bw.tolist())
            # what crosses NVLink INTO a GPU per step: the partner rows it reads, the accepted rows its peers write into
            # it, and (not counted in the algorithmic figure) the request / acknowledge packets of the opposite streams
            extra["exchange_every_step"] = {
                "ms_per_step": t_glob, "value": n / (t_glob * 1e-3),
                "nvlink_read_bytes_per_gpu_step": rd, "nvlink_write_bytes_per_gpu_step": wr,
                "roofline": {"bound": "nvlink", "achieved": (rd + wr) / (t_glob * 1e-3) / 1e9, "peak": sm, "unit": "GB/s",
                             "frac": (rd + wr) / (t_glob * 1e-3) / 1e9 / sm,
                             "traffic": "inbound per GPU: remote partner rows read + accepted rows written by the peers",
                             "peak_source": "scorus_peer_read_bandwidth: all ranks reading the next rank's ensemble at once, "
                                            "SM 16-byte loads (min over ranks), measured in this run"},
                "note": "the reference's whole-ensemble matching on every step (k = 1)"}
            # NCCL's all-gather of the ensemble (what BASELINE.json's wording would do every step), for the record
            buf = torch.empty((n, s.pitch), dtype=torch.float64, device="cuda")
            def ag(k):
                for _ in range(k):
                    dist.all_gather_into_tensor(buf, s.ens)
            ag(1)
            t_ag = timed(ag, 3)
            extra["nvlink"] = {"peer_read_copy_engine_gbps": ce, "peer_read_sm_loads_gbps": sm,
                               "peer_read_scrambled_rows_gbps": rr,
                               "nccl_allgather_ms": t_ag,
                               "nccl_allgather_busbw_gbps": n_local * s.pitch * 8 * (world - 1) / (t_ag * 1e-3) / 1e9}
            del buf
            if os.environ.get("BENCH_DIAG"):
                kk = max(2, args.global_every)
                def cycle(c):
                    for _ in range(c):
                        s.sampler.sample_pt_peer(2.0, spec, beta, 1)
                        s.sampler.sample_pt(2.0, spec, beta, kk - 1)
                t_cyc = timed(cycle, 10)
                extra["diag"] = {"per_rank_ms": diag, "order": ["local", "exchange", "allgather", "cycle"],
                                 "cycle_ms_per_cycle": t_cyc, "cycle_steps": kk}

    # burn-in, untimed and on top of the W warm-up steps: the timed region of a multi-GPU run is a few milliseconds, shorter
    # than the clock ramp of a GPU that idled while the host built the ensemble (measured at 8 GPUs: the same 100 steps
    # took 0.16-0.21 ms/step cold and 0.14 ms/step after a second of work)
    t_burn = time.perf_counter()
    while time.perf_counter() - t_burn < float(os.environ.get("BENCH_BURN_IN_S", "0.4")):
        do_steps(8)
        torch.cuda.synchronize()
    run_steps(0, W)
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    run_steps(W, args.steps)
    ev1.record(stream)
    clocks.sample_until(ev1)     # the GPU is still working through the queue: clocks under load, no launch in its way
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    clk = clocks.stop()
    launches = launch_count() - l0
    ms = ev0.elapsed_time(ev1)
    if dist:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    stats = s.stats()
    value = n * args.steps / (ms * 1e-3)

    # ---- the ladder swap on its own (reported separately; amortised into `value` above) ----
    swap_stage = None
    if swap_every and nbeta > 1 and can_swap:
        k_sw = 20
        sw0 = stats["swapped"]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(k_sw):
            swap(beta)
            do_steps(1)              # a step in between, as in the drivers: the next swap sees fresh log-probs
        e1.record(stream)
        torch.cuda.synchronize()
        t_both = e0.elapsed_time(e1) / k_sw
        e0.record(stream)
        do_steps(k_sw)
        e1.record(stream)
        torch.cuda.synchronize()
        t_step = e0.elapsed_time(e1) / k_sw
        if dist:
            tt = torch.tensor([t_both, t_step], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            t_both, t_step = float(tt[0].item()), float(tt[1].item())
        moved = (s.stats()["swapped"] - sw0) / k_sw          # accepted exchanges per call = rows moved / 2
        sw_bytes = 16.0 * n + 2.0 * moved * 16.0 * dim       # lp read + write; every accepted exchange moves two rows
        t_swap = max(t_both - t_step, 1e-6)
        swap_stage = {"ms_per_call": t_swap, "accepted_exchanges_per_call": moved,
                      "algorithmic_bytes_per_call": sw_bytes, "gbps": sw_bytes / (t_swap * 1e-3) / 1e9,
                      "note": "swap + one step minus one step, 20 calls (the rows move inside the step that follows the "
                              "swap); bound by the sweep's dependent steps and scattered accesses, not by HBM"}
        stats = s.stats()

    # ---- e2e: the reference-shaped call on pinned host buffers, copies inside the timed region ----
    e2e = None
    if not args.no_e2e:
        h_ens = sb.pinned_empty((n_local, dim))
        h_lp = sb.pinned_empty((n_local,))
        shard_local = world > 1 and args.move == "stretch" and (s.mode == "rungs" or args.matching == "local")
        if world > 1 and args.move == "stretch":
            # every rank makes the reference-shaped call on its own shard: upload, the step the schedule prescribes
            # (shard-local, or one-sided whole-ensemble matching), write-back of the shard's accepted walkers
            def host_step():
                s.sample_pt_host(h_ens, h_lp, 2.0, spec, beta)
            s.sampler.download(h_ens, h_lp)
            api = ("per rank: ShardedSampler.sample_pt_host -> scorus_sample_pt_host / scorus_sample_pt_peer_host on its shard "
                   "(pinned host buffers, one step per call)")
        elif world > 1:
            def host_step():  # upload shard, one step, download shard
                s.upload(h_ens, h_lp)
                do_steps(1)
                s.sampler.download(h_ens, h_lp)
            s.sampler.download(h_ens, h_lp)
            api = "ShardedSampler.upload + step + download (pinned host buffers, one step per call)"
        else:
            def host_step():
                if args.move == "hmc":
                    s.hmc_sample_ensemble_pt_host(h_ens, h_lp, hmc_eps, beta, 10, hmc_param)
                elif args.move == "twalk":
                    s.twalk_sample_host(h_ens, h_lp, tw, beta)
                else:
                    s.sample_pt_host(h_ens, h_lp, 2.0, spec, beta)
            s.download(h_ens, h_lp)
            api = {"twalk": "Sampler.twalk_sample_host -> scorus_twalk_sample_host",
                   "hmc": "Sampler.hmc_sample_ensemble_pt_host -> scorus_hmc_sample_ensemble_pt_host",
                   "stretch": "Sampler.sample_pt_host -> scorus_sample_pt_host"}[args.move] + " (pinned host buffers in/out, one step per call)"
        host_step()  # warm-up
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        acc0 = s.stats()["accepted"]
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            host_step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        acc_per_step = (s.stats()["accepted"] - acc0) / args.e2e_steps
        if dist:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        # whole job: every rank copies its own shard (rank 0's accept count stands for the others')
        bytes_dir = (n_local * dim * 8 + n_local * 8) * world
        # host call on pinned buffers: only accepted walkers are written back (scatter download)
        scatter = (world == 1 or shard_local or (args.move == "stretch" and args.matching != "global")) \
            and os.environ.get("SCORUS_SCATTER_DOWNLOAD", "1") != "0"
        d2h = int(acc_per_step * world * (dim * 8 + 8)) if scatter else bytes_dir
        e2e = {"value": n * args.e2e_steps / dt, "unit": "walker-steps/s", "h2d_bytes_per_step": bytes_dir,
               "d2h_bytes_per_step": d2h, "steps": args.e2e_steps, "ms_per_step": 1e3 * dt / args.e2e_steps,
               "api": api}
        if world == 1 and args.move == "stretch":
            # extra, not the headline: a driver that records every 10th iteration (examples/test_emcee.rs:82-84)
            # makes one host round trip per 10 steps (scorus_sample_pt_host_n)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(3):
                s.sample_pt_host(h_ens, h_lp, 2.0, spec, beta, nsteps=10)
            torch.cuda.synchronize()
            dt10 = time.perf_counter() - t0
            e2e["thin10"] = {"value": n * 30 / dt10, "unit": "walker-steps/s", "ms_per_step": 1e3 * dt10 / 30,
                             "note": "10 steps per host round trip; bytes per step are a tenth of the above"}

    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel: algorithmic bytes (or flops) per launch / measured launch time ----
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    kernel_name = {6: "step_reg_kernel<LpCorrGaussian> (K1 + K3 DMMA tile)"}.get(kind, "step_reg_kernel (K1)")
    if args.move == "twalk":
        kernel_name = "twalk_half_kernel (both half-passes of a t-walk step fused in one launch)"
    if args.move == "hmc":
        kernel_name = "hmc_kernel (10 leapfrog steps per walker in registers: bound by fp64 arithmetic, not HBM)"
    traffic = None
    prof = os.path.join(ROOT, "profiles", f"traffic_{args.workload}.json")
    if os.path.exists(prof) and args.move == "stretch" and world == 1 and not args.n_walkers and args.dtype == "f64":  # measured for this exact launch shape
        traffic = json.load(open(prof)).get("dram_bytes_per_launch")
    if kind == 6:
        # dense quadratic form: 2 D^2 + 2 D flops per walker-step against an fp64 DGEMM
        # peak measured here (MEASURED_PEAKS.json has no fp64 entry): torch.matmul fp64, best of 5
        m = 4096
        a_ = torch.randn(m, m, device="cuda", dtype=torch.float64)
        b_ = torch.randn(m, m, device="cuda", dtype=torch.float64)
        best = 0.0
        for _ in range(6):
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record(); c_ = a_ @ b_; t1.record(); torch.cuda.synchronize()
            best = max(best, 2.0 * m ** 3 / (t0.elapsed_time(t1) * 1e-3) / 1e12)
        flops_per_launch = (2 * dim * dim + 2 * dim) * n_local
        achieved = flops_per_launch * args.steps / (ms * 1e-3) / 1e12
        roofline = {"bound": "tensor", "achieved": achieved, "peak": best, "unit": "TFLOP/s", "frac": achieved / best,
                    "traffic": traffic, "kernel": kernel_name,
                    "peak_source": "cuBLAS fp64 DGEMM 4096^3 via torch.matmul, best of 6, measured in this run",
                    "algorithmic_flops_per_walker_step": 2 * dim * dim + 2 * dim, "walkers_per_launch": n_local,
                    # the workload's L is an upper-triangular (Cholesky-type) factor: the kernel starts the k loop of
                    # column block n0 at n0, so it EXECUTES this fraction of the accounted dense 2 D^2 flops
                    # ("a triangular-factor formulation does half -- still account at 2 D^2")
                    "executed_fraction_of_accounted_flops": sum(dim - n0 for n0 in range(0, dim, 32)) / (dim * len(range(0, dim, 32))),
                    "hbm_frac": b_alg(dim, ELEM) * n_local * args.steps / (ms * 1e-3) / 1e9 / peak}
    else:
        bytes_per_launch = b_alg(dim, ELEM) * n_local
        achieved = bytes_per_launch * args.steps / (ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "kernel": kernel_name, "peak_source": peak_src,
                    "algorithmic_bytes_per_walker_step": b_alg(dim, ELEM), "walkers_per_launch": n_local,
                    "note": "launch time = device time of the timed region / steps (one step kernel launch per step"
                            + (", swap kernels amortised in)" if swap_every else ")")
                            + "; the algorithmic figure counts every row as written, the kernel writes accepted rows "
                              "only, so frac can exceed 1: traffic_frac is the ncu-measured DRAM traffic over the same "
                              "time against the same peak"}
        if traffic:
            roofline["traffic_frac"] = traffic * args.steps / (ms * 1e-3) / 1e9 / peak
        if world == 1 and args.move == "stretch" and args.dtype == "f64" and n_local * dim * 8 > 2 * 126e6:
            # what HBM delivers for the kernel's ACCESS PATTERN -- every row read once in a scrambled order -- measured
            # here on a second context of the same shape (a world of one rank reads its own ensemble)
            try:
                probe = sb.Sampler(sb.Gaussian(0.5), n_local, dim, seed=1, device=local_rank)
                probe.init_ensemble(0.0, 1.0)
                he, hl = probe.ipc_export()
                probe.ipc_attach(1, 0, [he], [hl])
                _, seq, rows = probe.peer_read_bandwidth(0, 0, 5)
                probe.close()
                roofline["access_pattern"] = {
                    "sequential_read_gbps": seq, "scrambled_rows_read_gbps": rows,
                    "traffic_over_scrambled_rows": (traffic * args.steps / (ms * 1e-3) / 1e9 / rows) if traffic else None,
                    "note": "pure reads, 64 warps per SM: the ceiling for partner rows of a random matching; "
                            "profiles/r2_notes.md section 4"}
            except Exception as ex:   # a measurement aid: never fails the line
                roofline["access_pattern"] = {"error": str(ex)}

    cpu = None
    if not args.no_cpu and world == 1 and args.move == "stretch":
        cpu = cpu_reference_leg(args.workload, 20, 1)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample", "stage_share", "flat_threaded")}
    elif not args.no_cpu and world == 1:
        cpu = cpu_move_leg(args.workload, args.move)

    line = {
        "metric": "walker_steps_per_sec", "value": value, "unit": "walker-steps/s", "n_gpus": world,
        "steps": args.steps, "warmup": W, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": desc, "n_walkers": n, "dim": dim, "nbeta": nbeta, "ufs": ufs, "prob": prob, "a": 2.0,
                   "move": args.move, "parallelism": parallelism, "swap_every": swap_every,
                   "l2": f"resident ensemble {n_local * dim * ELEM / 2**20:.0f} MiB per GPU vs 126 MB L2 (no flush needed)"
                         if n_local * dim * ELEM > 2 * 126e6 else "ensemble fits in L2: numbers are L2-resident, not HBM"},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clk,
        "accept_rate": stats["accepted"] / max(1, stats["proposed"]),
    }
    if swap_stage:
        line["swap_stage"] = swap_stage
    if extra:
        line["extra"] = extra
    if parity_check:
        line["parity_check"] = parity_check
    print(json.dumps(line))
    if dist:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
