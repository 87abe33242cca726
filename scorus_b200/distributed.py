"""Walker / rung sharding over the GPUs of one box: one process per GPU, torch.distributed (NCCL) for
the plumbing, the library's own kernels for the work.

Two layouts, chosen from (nbeta, world size):

* ``rungs``   -- tempered runs with nbeta % world == 0: every rank holds whole rungs.  The pairing never
  leaves a rung (src/mcmc/ensemble_sample.rs:136-146), so ``sample_pt`` needs NO communication.
  ``swap_walkers`` couples adjacent rungs (src/mcmc/utils.rs:89-108): the log-probs are all-gathered
  (8*N bytes), every rank computes the same swap plan on its GPU (scorus_swap_plan_device), and only the
  rows that cross a rank boundary are exchanged (all_to_all_single).
* ``walkers`` -- single-temperature runs: contiguous walker shards.
    - matching="p2p": the reference's whole-rung matching, one-sided.  Every pair is owned by the rank holding its
      even-position member; the owner reads the partner's row straight from its home GPU over NVLink (CUDA IPC
      mappings), runs the fused step for both walkers and writes accepted rows back.  One flag barrier in peer
      memory per step, no NCCL call, nothing computed twice (scorus_sample_pt_peer).  Bit-identical to one GPU.
    - matching="global": the same matching, literally as BASELINE.json words it: an NCCL all-gather of the
      ensemble per step, then scorus_sample_pt_global.  NVLink-bound -- the baseline to beat.
    - matching="local": every shard pairs within itself, no exchange: a BLOCKED matching (one block per shard).
  ``global_every = k`` interleaves: steps whose index is a multiple of k use the whole-ensemble matching ("p2p" or
  "global"), the others the blocked one.  The schedule depends on the step index only, so every step's matching is
  independent of the state and each step is a valid stretch move; cross-shard mixing costs 1/k of the NVLink traffic.

Random streams are keyed by global walker / rung / block ids (Sampler.set_shard, set_matching_blocks), so EVERY mode
reproduces a single-GPU chain bit for bit: ``rungs`` and ``walkers`` with k = 1 the reference-matching chain,
``walkers`` with k > 1 or "local" the single-GPU chain with ``set_matching_blocks(world, 0, k)``.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np

from . import _lib as L
from .ensemble_sample import LogProb, Sampler, UpdateFlagSpec


# ------------------------------------------------------------------------------------------------
# pure host logic (unit-tested on CPU with gloo, tests/test_distributed_cpu.py)
# ------------------------------------------------------------------------------------------------
def shard_layout(n_walkers: int, nbeta: int, world: int, rank: int):
    """-> (mode, walker_offset, n_local, rung_offset, nbeta_local)."""
    if n_walkers % world:
        raise ValueError("the number of walkers must divide over the ranks")
    n_local = n_walkers // world
    if n_local % 2:
        raise ValueError("every shard needs an even number of walkers")
    if nbeta > 1 and nbeta % world == 0:
        nb_local = nbeta // world
        return "rungs", rank * n_local, n_local, rank * nb_local, nb_local
    return "walkers", rank * n_local, n_local, 0, nbeta


def owned_pairs(order: np.ndarray, n_local: int, rank: int) -> np.ndarray:
    """The pairs of a whole-ensemble matching that `rank` computes in the one-sided exchange step, as [k, 2] walker ids.
    order[j] = walker at matching position j (positions 2k, 2k+1 are partners).  Rule (api_shard.cu:
    shard_pairs_kernel): a pair belongs to the rank holding its EVEN-position member -- every pair on exactly one
    rank, which updates both of its walkers; positions are a keyed bijection of the walkers, so the split is balanced
    to O(sqrt(n))."""
    order = np.asarray(order, dtype=np.int64)
    a, b = order[0::2], order[1::2]
    mine = a // n_local == rank
    return np.stack([a[mine], b[mine]], axis=1)


def is_exchange_step(step_index: int, global_every: int) -> bool:
    """The k-schedule: steps whose index is a multiple of k use the whole-ensemble matching (the same rule as
    scorus_set_matching_blocks on one GPU)."""
    return global_every <= 1 or step_index % global_every == 0


def exchange_plan(src_all: np.ndarray, n_local: int, world: int, rank: int):
    """Who sends which rows to whom after a ladder swap.

    src_all[slot] = global slot whose (old) row must end in `slot`.  Returns
      local_dst, local_src : moves inside this rank (local indices; read old rows before writing)
      send_idx[q]          : local indices of my old rows that rank q needs, in q's slot order
      recv_dst[q]          : local indices of my slots filled from rank q, in my slot order
    Every rank derives the same plan from the same src_all, so no metadata is exchanged."""
    lo, hi = rank * n_local, (rank + 1) * n_local
    src_all = np.asarray(src_all, dtype=np.int64)
    mine = src_all[lo:hi]
    slots = np.arange(lo, hi)
    moved = mine != slots
    owner = mine // n_local
    loc = moved & (owner == rank)
    local_dst = (slots[loc] - lo).astype(np.int64)
    local_src = (mine[loc] - lo).astype(np.int64)
    send_idx, recv_dst = [], []
    for q in range(world):
        if q == rank:
            send_idx.append(np.zeros(0, dtype=np.int64))
            recv_dst.append(np.zeros(0, dtype=np.int64))
            continue
        theirs = src_all[q * n_local:(q + 1) * n_local]
        send_idx.append((theirs[(theirs >= lo) & (theirs < hi)] - lo).astype(np.int64))
        recv_dst.append((slots[moved & (owner == q)] - lo).astype(np.int64))
    return local_dst, local_src, send_idx, recv_dst


def exchange_plan_torch(src_all, n_local: int, world: int, rank: int):
    """exchange_plan with torch ops on whatever device src_all lives on (the GPU in production: the
    plan never visits the host except for the world x world matrix of row counts that
    all_to_all_single needs).  Returns (local_dst, local_src, send_idx, recv_dst, send_counts,
    recv_counts): send_idx / recv_dst are concatenated in peer order."""
    import torch
    n = src_all.numel()
    dev = src_all.device
    src = src_all.to(torch.int64) & 0xFFFFFFFF
    slot = torch.arange(n, device=dev, dtype=torch.int64)
    moved = src != slot
    owner = torch.div(src, n_local, rounding_mode="floor")       # rank holding the row that moves in
    holder = torch.div(slot, n_local, rounding_mode="floor")     # rank holding the destination slot
    cross = moved & (owner != holder)
    counts = torch.bincount(holder[cross] * world + owner[cross], minlength=world * world).view(world, world).cpu()
    recv_counts = [int(c) for c in counts[rank]]                 # rows I receive from each peer
    send_counts = [int(c) for c in counts[:, rank]]              # rows each peer receives from me
    lo = rank * n_local
    # rows I send: slots of other ranks fed from my range, in (peer, slot) = ascending slot order
    send_idx = src[cross & (owner == rank)] - lo
    # slots I fill from peers, grouped by peer (stable sort keeps slot order inside a peer)
    my = slice(lo, lo + n_local)
    rmask = cross[my]
    rslots = slot[my][rmask] - lo
    order = torch.argsort(owner[my][rmask], stable=True)
    recv_dst = rslots[order]
    lmask = moved[my] & (owner[my] == rank)
    local_dst = slot[my][lmask] - lo
    local_src = src[my][lmask] - lo
    return local_dst, local_src, send_idx, recv_dst, send_counts, recv_counts


# ------------------------------------------------------------------------------------------------
# device views (zero-copy torch tensors over the context's buffers)
# ------------------------------------------------------------------------------------------------
class _CudaArray:
    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (ptr, False),
                                         "version": 2, "strides": None}


def device_views(s: Sampler, device):
    import torch
    ens_ptr, pitch, lp_ptr = s.device_buffers()
    ens = torch.as_tensor(_CudaArray(ens_ptr, (s.n, pitch), "<f8"), device=device)
    lp = torch.as_tensor(_CudaArray(lp_ptr, (s.n,), "<f8"), device=device)
    return ens, lp, pitch


class ShardedSampler:
    """The sharded counterpart of Sampler: same calls, global beta list, local data."""

    def __init__(self, flogprob: LogProb, n_walkers: int, dim: int, nbeta: int = 1, seed: int = 0,
                 matching: str = "global", device: Optional[int] = None, group=None, p2p_swap: bool = True,
                 global_every: int = 1):
        import torch
        import torch.distributed as dist
        self.dist, self.torch, self.group = dist, torch, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.n_global, self.dim, self.nbeta = n_walkers, dim, nbeta
        self.mode, self.w_off, self.n_local, self.rung_off, self.nb_local = shard_layout(n_walkers, nbeta, self.world,
                                                                                        self.rank)
        if matching not in ("global", "p2p", "local"):
            raise ValueError("matching must be 'global', 'p2p' or 'local'")
        self.matching = matching
        # global_every = k > 1 (matching "global" / "p2p"): only the steps whose index is a multiple of k use the
        # whole-ensemble matching, the others pair inside each shard
        self.global_every = max(1, int(global_every))
        if self.mode == "walkers" and nbeta > 1:
            # a tempered ensemble whose rungs do not divide over the ranks: shard boundaries cut through rungs, so there
            # is no shard-local step -- every step uses the whole-rung matching, one-sided, and so does the ladder swap
            if matching != "p2p":
                raise ValueError("tempered walker shards (nbeta % world != 0) need matching='p2p'")
            self.global_every = 1
        dev = torch.cuda.current_device() if device is None else device
        self.device = torch.device("cuda", dev)
        self._flogprob, self._seed, self._full = flogprob, seed, None
        self.sampler = Sampler(flogprob, self.n_local, dim, seed=seed, device=dev)
        self.sampler.set_shard(self.w_off, self.rung_off, n_walkers)
        if self.mode == "walkers":
            # a shard is matching block `rank` of the whole ensemble: its local steps take that block's stream keys
            self.sampler.set_matching_blocks(1, self.rank, 0)
        # every launch and every collective of this object goes to ONE stream, captured here
        self._stream = torch.cuda.current_stream(self.device)
        self.sampler.set_stream(self._stream.cuda_stream)
        self.ens, self.lp, self.pitch = device_views(self.sampler, self.device)
        self._ens_all = self._lp_all = self._src_all = None
        self.p2p = False
        if self.world > 1 and ((self.mode == "walkers" and matching == "p2p") or (self.mode == "rungs" and p2p_swap)):
            self.p2p = True
            # map every peer's ensemble / log-probs (CUDA IPC) for one-sided partner reads over NVLink
            mine = self.sampler.ipc_export()
            handles = [None] * self.world
            dist.all_gather_object(handles, mine, group=group)
            self.sampler.ipc_attach(self.world, self.rank, [h[0] for h in handles], [h[1] for h in handles])

    # -- data
    def upload(self, ens_local: np.ndarray, lp_local: Optional[np.ndarray] = None):
        self.sampler.upload(ens_local, lp_local)

    def download(self):
        return self.sampler.download()

    def _gather_buffers(self):
        t = self.torch
        if self._ens_all is None:
            self._ens_all = t.empty((self.n_global, self.pitch), dtype=t.float64, device=self.device)
            self._lp_all = t.empty((self.n_global,), dtype=t.float64, device=self.device)
            self._src_all = t.empty((self.n_global,), dtype=t.int32, device=self.device)
        return self._ens_all, self._lp_all, self._src_all

    # -- the hot path
    def sample_pt(self, a: float, ufs: UpdateFlagSpec, beta_list: Sequence[float], nsteps: int = 1):
        beta = np.ascontiguousarray(beta_list, dtype=np.float64)
        if ufs.kind == L.UFS_MASK and self.world > 1:
            # UpdateFlagSpec::Func (ensemble_sample.rs:31,52,150-153): the closure is called once per walker of the WHOLE
            # ensemble, in order, every step.  Every rank holds the same closure in the same state, so every rank makes
            # the same n_global calls -- the flags are then exactly the single-GPU run's -- and hands its shard (local
            # step) or the whole matrix (whole-ensemble step: the owner of a pair needs both walkers' flags) on.
            # UpdateFlagSpec.Mask: one step's flags of the whole ensemble, already evaluated.
            if ufs.func is None and nsteps != 1:
                raise ValueError("a mask covers exactly one step")
            for _ in range(nsteps):
                flags = ufs.evaluate(self.n_global) if ufs.func is not None else ufs.mask
                if flags.shape[0] != self.n_global:
                    raise ValueError(f"mask has {flags.shape[0]} rows for {self.n_global} walkers")
                if self._step_is_local():
                    flags = flags[self.w_off:self.w_off + self.n_local]
                self._sample_pt(a, UpdateFlagSpec.Mask(flags), beta, 1)
            return
        self._sample_pt(a, ufs, beta, nsteps)

    def _sample_pt(self, a: float, ufs: UpdateFlagSpec, beta: np.ndarray, nsteps: int):
        if self.mode == "rungs":
            local_beta = beta[self.rung_off:self.rung_off + self.nb_local]
            self.sampler.sample_pt(a, ufs, local_beta, nsteps)      # no communication
            return
        if self.matching == "local" or self.world == 1:
            self.sampler.sample_pt(a, ufs, beta, nsteps)
            return
        done, k = 0, self.global_every
        while done < nsteps:
            index = self.sampler.counters[0]                         # the step counter drives the schedule
            if not is_exchange_step(index, k):                       # shard-local steps up to the next global one
                run = min(nsteps - done, k - index % k)
                self.sampler.sample_pt(a, ufs, beta, run)
                done += run
                continue
            if self.matching == "p2p":
                run = nsteps - done if k == 1 else 1                 # consecutive global steps: one library call
                self.sampler.sample_pt_peer(a, ufs, beta, run)
                done += run
                continue
            with self.torch.cuda.stream(self._stream):               # all-gather + global-matching step
                ens_all, lp_all, _ = self._gather_buffers()
                self.dist.all_gather_into_tensor(ens_all, self.ens, group=self.group)
                self.dist.all_gather_into_tensor(lp_all, self.lp, group=self.group)
            self.sampler.sample_pt_global(a, ufs, beta, ens_all.data_ptr(), lp_all.data_ptr())
            done += 1

    def _step_is_local(self) -> bool:
        """Whether the step the schedule prescribes next pairs walkers inside this rank's shard (and so wants a shard's
        worth of host-evaluated flags) rather than across the whole ensemble."""
        if self.mode == "rungs" or self.matching == "local" or self.world == 1:
            return True
        return not is_exchange_step(self.sampler.counters[0], self.global_every)

    def sample_pt_host(self, ens_local: np.ndarray, lp_local: np.ndarray, a: float, ufs: UpdateFlagSpec,
                       beta_list: Sequence[float]):
        """The reference-shaped call (host buffers in, ONE step, host buffers out) on this rank's shard; collective.
        The step is the one the schedule prescribes for the current step index: shard-local, or the whole-ensemble
        matching one-sided.  With pinned buffers only the shard's accepted walkers are written back."""
        beta = np.ascontiguousarray(beta_list, dtype=np.float64)
        if self.mode == "rungs":
            self.sampler.sample_pt_host(ens_local, lp_local, a, ufs, beta[self.rung_off:self.rung_off + self.nb_local])
            return
        index = self.sampler.counters[0]
        if self.matching == "local" or self.world == 1 or index % self.global_every != 0:
            self.sampler.sample_pt_host(ens_local, lp_local, a, ufs, beta)
        elif self.matching == "p2p":
            self.sampler.sample_pt_peer_host(ens_local, lp_local, a, ufs, beta)
        else:
            self.upload(ens_local, lp_local)
            self.sample_pt(a, ufs, beta, 1)
            self.sampler.download(ens_local, lp_local)

    def twalk_sample(self, param, beta_list: Sequence[float], nsteps: int = 1):
        """twalk::sample (src/mcmc/twalk.rs:586-648) on this rank's shard.  Rung shards (whole rungs per rank):
        the matching never leaves a rung, so no communication and the same chain as one GPU.  Walker shards:
        the matching is drawn inside the shard (`matching="local"`), or over the whole ensemble with the compute
        replicated on every rank (_twalk_replicated)."""
        beta = np.ascontiguousarray(beta_list, dtype=np.float64)
        if self.mode == "rungs":
            self.sampler.twalk_sample(param, beta[self.rung_off:self.rung_off + self.nb_local], nsteps)
            return
        if self.matching != "local" and self.world > 1:
            self._twalk_replicated(param, beta, nsteps)
            return
        self.sampler.twalk_sample(param, beta, nsteps)

    def _twalk_replicated(self, param, beta: np.ndarray, nsteps: int):
        """The t-walk under the reference's whole-rung matching (twalk.rs:600-612) across walker shards, REPLICATED:
        the shards are all-gathered once per call, every rank steps the whole ensemble with the single-GPU kernel (same
        seed and counters, so all ranks compute the same chain) and keeps its own rows.  Identical to the single-GPU
        chain and no faster than it -- the t-walk kernels have no owner-computes form yet; rung shards (tempered runs)
        and matching="local" are the paths that scale."""
        t = self.torch
        if self._full is None:
            self._full = Sampler(self._flogprob, self.n_global, self.dim, seed=self._seed, device=self.device.index)
            self._full.set_stream(self._stream.cuda_stream)
            self._full.init_ensemble(0.0, 1.0)           # (marks the context as holding an ensemble; overwritten below)
        self.ens, self.lp, self.pitch = device_views(self.sampler, self.device)
        with t.cuda.stream(self._stream):
            ens_all, lp_all, _ = self._gather_buffers()
            self.dist.all_gather_into_tensor(ens_all, self.ens, group=self.group)
            self.dist.all_gather_into_tensor(lp_all, self.lp, group=self.group)
            fe, fl, _ = device_views(self._full, self.device)
            fe.copy_(ens_all)
            fl.copy_(lp_all)
        self._full.counters = self.sampler.counters
        self._full.twalk_sample(param, beta, nsteps)
        with t.cuda.stream(self._stream):
            fe, fl, _ = device_views(self._full, self.device)
            self.ens.copy_(fe[self.w_off:self.w_off + self.n_local])
            self.lp.copy_(fl[self.w_off:self.w_off + self.n_local])
        self.sampler.counters = self._full.counters

    def hmc_sample_ensemble_pt(self, epsilon, beta_list: Sequence[float], l: int, param, nsteps: int = 1):
        """hmc::naive::sample_ensemble_pt (src/mcmc/hmc/naive.rs:122-292) on this rank's shard: walkers are
        independent, so there is no communication.  Rung shards: `epsilon` is the whole ladder's array, this rank
        updates (in place) and returns the counts of its own rungs only.  Walker shards of one rung adapt ONE step size:
        the ranks' net adaptation counts are summed every step (below)."""
        beta = np.ascontiguousarray(beta_list, dtype=np.float64)
        if self.mode == "rungs":
            lo, hi = self.rung_off, self.rung_off + self.nb_local
            return self.sampler.hmc_sample_ensemble_pt(epsilon[lo:hi], beta[lo:hi], l, param, nsteps)
        if self.world > 1 and param.adj_factor != 0.0:
            # walker shards of ONE rung with an adapting step size: every rank runs the step with the step size frozen,
            # the nets (#grow - #shrink, naive.rs:274-283) are summed over the ranks and epsilon (1 + adj)^net is applied
            # once -- what one GPU does for a rung above 4096 walkers, so epsilon matches it bit for bit
            if len(beta) != 1:
                raise NotImplementedError("adaptive HMC over walker shards: one rung (rung shards cover tempered runs)")
            from .hmc import HmcParam
            frozen = HmcParam(param.target_accept_ratio, 0.0)
            total = np.zeros(1, dtype=np.uint64)
            for _ in range(nsteps):
                total += self.sampler.hmc_sample_ensemble_pt(epsilon, beta, l, frozen, 1)
                net = self.torch.tensor([self.sampler.hmc_adapt_net()], dtype=self.torch.int64, device=self.device)
                self.dist.all_reduce(net, group=self.group)
                m, base, pw = abs(int(net.item())), 1.0 + param.adj_factor, 1.0
                while m:                                   # repeated squaring, as hmc_adapt_count_kernel does
                    if m & 1:
                        pw = pw * base
                    base = base * base
                    m >>= 1
                epsilon[0] = epsilon[0] / pw if int(net.item()) < 0 else epsilon[0] * pw
            return total
        return self.sampler.hmc_sample_ensemble_pt(epsilon, beta, l, param, nsteps)

    def _barrier(self):
        """Stream-ordered barrier across ranks: flags in peer memory (scorus_peer_barrier), no NCCL call, no host
        synchronisation."""
        self.sampler.peer_barrier()

    def swap_walkers(self, beta_list: Sequence[float]):
        beta = np.ascontiguousarray(beta_list, dtype=np.float64)
        if self.world == 1:
            self.sampler.swap_walkers(beta)
            # the swap leaves the rows in the context's second buffer (the two change roles): refresh the views
            self.ens, self.lp, self.pitch = device_views(self.sampler, self.device)
            return
        if self.mode != "rungs":
            # walker shards: the one-sided swap works on global slots, wherever the rung boundaries fall
            if len(beta) > 1:
                self.sampler.swap_walkers_peer(beta)
            return
        t, dist = self.torch, self.dist
        if self.p2p:
            # one-sided, one library call: barrier, the plan from log-probs read out of the peers' memory, incoming rows
            # pulled over NVLink, barrier, commit.  A step only writes the rank's own rows: no barrier after the commit.
            self.sampler.swap_walkers_peer(beta)
            return
        with t.cuda.stream(self._stream):
            _, lp_all, src_all = self._gather_buffers()
            dist.all_gather_into_tensor(lp_all, self.lp, group=self.group)
        self.sampler.swap_plan_device(beta, lp_all.data_ptr(), src_all.data_ptr())
        with t.cuda.stream(self._stream):
            local_dst, local_src, send_idx, recv_dst, send_counts, recv_counts = exchange_plan_torch(
                src_all, self.n_local, self.world, self.rank)
            send = self.ens.index_select(0, send_idx)                # old rows, gathered before any write
            local_rows = self.ens.index_select(0, local_src)
            recv = t.empty((sum(recv_counts), self.pitch), dtype=t.float64, device=self.device)
            dist.all_to_all_single(recv, send, recv_counts, send_counts, group=self.group)
            if local_dst.numel():
                self.ens.index_copy_(0, local_dst, local_rows)
            if recv_dst.numel():
                self.ens.index_copy_(0, recv_dst, recv)
            self.lp.copy_(lp_all[self.w_off:self.w_off + self.n_local])

    def sync(self):
        self.sampler.sync()

    def close(self):
        """Releases the context; every rank waits until all are done with the buffers it exported."""
        if self.sampler is None:
            return
        self.sampler.sync()
        if self.world > 1:
            self.dist.barrier(group=self.group)
        self.ens = self.lp = self._ens_all = self._lp_all = self._src_all = None
        if self._full is not None:
            self._full.close()
            self._full = None
        self.sampler.close()
        self.sampler = None

    def stats(self):
        return self.sampler.stats()
