"""Host-side logic of the multi-GPU path on CPU: shard layout, the swap exchange plan, and a
world_size-2 gloo run that moves rows exactly as ShardedSampler.swap_walkers does (with the plan
computed from an oracle-made swap), checked against the oracle's result on the whole ensemble."""
import os
import socket

import numpy as np
import pytest

from scorus_b200.distributed import exchange_plan, exchange_plan_torch, shard_layout


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
