"""Host-side logic of the multi-GPU path on CPU: shard layout, the swap exchange plan, and a
world_size-2 gloo run that moves rows exactly as ShardedSampler.swap_walkers does (with the plan
computed from an oracle-made swap), checked against the oracle's result on the whole ensemble."""
import os
import socket

import numpy as np
import pytest

from scorus_b200.distributed import exchange_plan, exchange_plan_torch, is_exchange_step, owned_pairs, shard_layout


def test_shard_layout():
    assert shard_layout(64, 8, 2, 1) == ("rungs", 32, 32, 4, 4)
    assert shard_layout(64, 1, 4, 3) == ("walkers", 48, 16, 0, 1)
    assert shard_layout(60, 3, 2, 0)[0] == "walkers"      # rungs do not divide over the ranks
    with pytest.raises(ValueError):
        shard_layout(10, 1, 4, 0)
    with pytest.raises(ValueError):
        shard_layout(12, 1, 4, 0)                          # 3 walkers per shard: odd


def _swap_source_map(oracle, ens, lp, beta, seed):
    """src[slot] for one oracle swap: tag every walker with its slot in an extra column."""
    n = ens.shape[0]
    tagged = np.ascontiguousarray(np.c_[ens, np.arange(n, dtype=np.float64)])
    lp2 = lp.copy()
    jvec, r = oracle.draw_swap_streams(oracle.Rng(seed), n, len(beta))
    rc, _ = oracle.swap_walkers_fixed(tagged, lp2, beta, jvec, r)
    assert rc == 0
    return tagged[:, -1].astype(np.int64), np.ascontiguousarray(tagged[:, :-1]), lp2


def test_exchange_plan_is_consistent_between_ranks(oracle):
    g = np.random.default_rng(0)
    n, d, nb, world = 64, 3, 8, 4
    ens, lp = g.normal(size=(n, d)), g.normal(size=n) * 3
    beta = 2.0 ** -np.arange(nb, dtype=np.float64)
    src, want, _ = _swap_source_map(oracle, ens, lp, beta, 5)
    assert sorted(src.tolist()) == list(range(n)) and np.any(src != np.arange(n))
    n_local = n // world
    out = ens.copy()
    plans = [exchange_plan(src, n_local, world, r) for r in range(world)]
    for r, (ld, ls, send, recv) in enumerate(plans):
        lo = r * n_local
        out[lo + ld] = ens[lo + ls]
        for q in range(world):
            assert len(recv[q]) == len(plans[q][2][r])            # what r expects from q == what q sends to r
            out[lo + recv[q]] = ens[q * n_local + plans[q][2][r]]
    assert np.array_equal(out, want)


def test_torch_plan_equals_numpy_plan(oracle):
    import torch
    g = np.random.default_rng(2)
    n, d, nb, world = 96, 2, 12, 4
    ens, lp = g.normal(size=(n, d)), g.normal(size=n) * 3
    src, _, _ = _swap_source_map(oracle, ens, lp, 2.0 ** -np.arange(nb, dtype=np.float64), 3)
    n_local = n // world
    for rank in range(world):
        ld, ls, send, recv = exchange_plan(src, n_local, world, rank)
        tld, tls, tsend, trecv, sc, rc = exchange_plan_torch(torch.from_numpy(src.astype(np.int32)), n_local, world, rank)
        assert np.array_equal(tld.numpy(), ld) and np.array_equal(tls.numpy(), ls)
        assert np.array_equal(tsend.numpy(), np.concatenate(send)) and np.array_equal(trecv.numpy(), np.concatenate(recv))
        assert sc == [len(x) for x in send] and rc == [len(x) for x in recv]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, src, ens, want, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_local = ens.shape[0] // world
    lo = rank * n_local
    mine = torch.from_numpy(ens[lo:lo + n_local].copy())
    ld, ls, send, recv = exchange_plan(src, n_local, world, rank)
    # same order of operations as ShardedSampler.swap_walkers: gather old rows, exchange, then write
    send_rows = [mine[torch.from_numpy(s)] for s in send]
    local_rows = mine[torch.from_numpy(ls)]
    recv_rows = [torch.empty((len(r), ens.shape[1]), dtype=torch.float64) for r in recv]
    reqs = []
    for qq in range(world):
        if qq == rank:
            continue
        if len(send[qq]):
            reqs.append(dist.isend(send_rows[qq].contiguous(), qq))
        if len(recv[qq]):
            reqs.append(dist.irecv(recv_rows[qq], qq))
    for r in reqs:
        r.wait()
    mine[torch.from_numpy(ld)] = local_rows
    for qq in range(world):
        if len(recv[qq]):
            mine[torch.from_numpy(recv[qq])] = recv_rows[qq]
    ok = bool(np.array_equal(mine.numpy(), want[lo:lo + n_local]))
    gathered = [torch.zeros(1, dtype=torch.int32) for _ in range(world)]
    dist.all_gather(gathered, torch.tensor([int(ok)], dtype=torch.int32))
    if rank == 0:
        q.put([int(t.item()) for t in gathered])
    dist.destroy_process_group()


def test_row_exchange_world2_gloo(oracle):
    import torch.multiprocessing as mp
    g = np.random.default_rng(1)
    n, d, nb, world = 48, 4, 6, 2
    ens, lp = g.normal(size=(n, d)), g.normal(size=n) * 2
    beta = 2.0 ** -np.arange(nb, dtype=np.float64)
    src, want, _ = _swap_source_map(oracle, ens, lp, beta, 9)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, src, ens, want, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=10) == [1] * world


def test_pair_ownership_partitions_the_matching(oracle):
    """Every pair of the whole-ensemble matching is owned by exactly one rank, the owner holds its even-position member,
    and the split is balanced; the schedule rule is the single-GPU one."""
    n, world = 1 << 14, 8
    n_local = n // world
    partner, _, _, _ = oracle.philox_sample_streams(3, 5, n, 4, 1, 2.0, oracle.UFS_ALL, 0.5, 0)
    # rebuild an order array from the partner map (any order of the pairs: ownership only needs "even member")
    firsts = np.flatnonzero(partner > np.arange(n))
    rng = np.random.default_rng(0)
    flip = rng.random(firsts.size) < 0.5
    a = np.where(flip, partner[firsts], firsts)
    b = np.where(flip, firsts, partner[firsts])
    order = np.stack([a, b], axis=1).reshape(-1)
    owned = [owned_pairs(order, n_local, r) for r in range(world)]
    allp = np.concatenate(owned)
    assert allp.shape[0] == n // 2 and np.array_equal(np.sort(allp.reshape(-1)), np.arange(n))
    for r, pr in enumerate(owned):
        assert np.all(pr[:, 0] // n_local == r)
        assert abs(pr.shape[0] - n // 2 // world) < 6 * np.sqrt(n / 2 / world)
    assert [is_exchange_step(s, 4) for s in range(6)] == [True, False, False, False, True, False]
    assert all(is_exchange_step(s, 1) for s in range(4))


def _owner_worker(rank, world, port, kind, params, ens, lp, beta, streams, want_e, want_lp, q):
    """The one-sided exchange step emulated on CPU ranks: each rank runs the oracle on the pairs it OWNS (reading both rows
    from the gathered pre-step state, as the device reads a remote partner over NVLink) and the results go back to the
    walkers' home ranks."""
    import torch
    import torch.distributed as dist
    from oracle import pyoracle as po
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, dim = ens.shape
    n_local = n // world
    partner, z, flags, u = streams
    firsts = np.flatnonzero(partner > np.arange(n))
    order = np.stack([firsts, partner[firsts]], axis=1).reshape(-1)       # even position: the lower id (any rule works)
    mine = owned_pairs(order, n_local, rank)
    ids = mine.reshape(-1)
    sub_e, sub_lp = np.ascontiguousarray(ens[ids]), np.ascontiguousarray(lp[ids])
    local = np.arange(ids.size, dtype=np.uint32) ^ 1
    rc, acc = po.sample_pt_fixed(kind, params, sub_e, sub_lp, beta, local, z[ids], flags[ids], u[ids])
    assert rc == 0
    # write-back to the home ranks: (walker id, row, lp) of everything this rank computed
    out = [None] * world
    dist.all_gather_object(out, (ids, sub_e, sub_lp))
    new_e, new_lp = ens.copy(), lp.copy()
    for i, e, l in out:
        new_e[i], new_lp[i] = e, l
    lo = rank * n_local
    ok = bool(np.array_equal(new_e[lo:lo + n_local], want_e[lo:lo + n_local]) and
              np.array_equal(new_lp[lo:lo + n_local], want_lp[lo:lo + n_local]))
    gathered = [torch.zeros(1, dtype=torch.int32) for _ in range(world)]
    dist.all_gather(gathered, torch.tensor([int(ok)], dtype=torch.int32))
    if rank == 0:
        q.put([int(t.item()) for t in gathered])
    dist.destroy_process_group()


def test_owner_computes_exchange_step_world2_gloo(oracle):
    """world_size 2 on CPU: pair ownership + write-back reproduce the oracle's whole-ensemble step exactly."""
    import torch.multiprocessing as mp
    g = np.random.default_rng(4)
    n, dim, world = 256, 6, 2
    ens = np.ascontiguousarray(g.uniform(0.9, 1.1, (n, dim)))
    lp = oracle.init_logprob(2, [], ens)
    beta = np.ones(1)
    streams = oracle.philox_sample_streams(21, 0, n, dim, 1, 2.0, oracle.UFS_PROB, 0.5, 0)
    want_e, want_lp = ens.copy(), lp.copy()
    rc, acc = oracle.sample_pt_fixed(2, [], want_e, want_lp, beta, *streams)
    assert rc == 0 and acc.sum() > 0
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_owner_worker, args=(r, world, port, 2, [], ens, lp, beta, streams, want_e, want_lp, q))
             for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=10) == [1] * world


def test_update_flag_spec_func_is_evaluated_once_per_walker_and_mask_carries_one_step():
    """UpdateFlagSpec::Func (ensemble_sample.rs:31,52,150-153) over shards: every rank evaluates the closure for the whole
    ensemble (`evaluate`) and hands rows of that one evaluation on (`Mask`)."""
    import pytest
    from scorus_b200 import UpdateFlagSpec
    calls = []

    def f():
        calls.append(len(calls))
        return [len(calls) % 2 == 0, True, False]
    spec = UpdateFlagSpec.Func(f)
    m = spec.evaluate(5)
    assert m.shape == (5, 3) and m.dtype == np.uint8 and len(calls) == 5
    assert m[:, 0].tolist() == [0, 1, 0, 1, 0] and m[:, 1].all() and not m[:, 2].any()
    shard = UpdateFlagSpec.Mask(m[2:4])
    st, keep = shard._c(2, 4)
    assert st.mask_len == 3 and keep.shape == (2, 3) and len(calls) == 5      # nothing is evaluated again
    with pytest.raises(ValueError):
        shard._c(3, 4)                                                         # rows != walkers
    with pytest.raises(IndexError):
        shard._c(2, 2)                                                         # flags longer than a walker (:70)
    with pytest.raises(ValueError):
        UpdateFlagSpec.Mask(np.ones(4))
