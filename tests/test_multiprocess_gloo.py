"""N > 1 host logic on CPU: world_size-2 gloo.  The chain ensemble shards into contiguous slices; the only
exchange on the path is the small sum of [sum ll, sum ll°, accept counts] (fetch_ll / accpt_rate, src/block_ensemble.jl:140,
152,175-179).  Here each rank holds its slice of synthetic per-chain values; the reduced statistics must equal the unsharded
ones and the slices must tile the ensemble and see the same synthetic data as the unsharded problem."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import dmt_b200
from dmt_b200 import configs
from dmt_b200 import host as H


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, M, nb, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)                      # same stream on every rank: the "unsharded" per-chain values
    ll = rng.normal(size=(nb, M)); llo = rng.normal(size=(nb, M)); acc = rng.random(size=(nb, M)) < 0.4
    lo, hi = H.shard_slice(M, rank, world)
    local = np.concatenate([[ll[:, lo:hi].sum(), llo[:, lo:hi].sum()], acc[:, lo:hi].sum(axis=1)])
    tot = H.dist_sum(local)
    full = configs.make_problem("lv", M, K=2, seed=9)
    part = configs.make_problem("lv", hi - lo, K=2, seed=9, chain_offset=lo)
    same_data = bool(np.array_equal(full.v[:, :, lo:hi], part.v) and np.array_equal(full.x0[:, lo:hi], part.x0))
    q.put((rank, lo, hi, tot, same_data))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_stats_allreduce_gloo():
    world, M, nb = 2, 37, 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, M, nb, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(0)
    ll = rng.normal(size=(nb, M)); llo = rng.normal(size=(nb, M)); acc = rng.random(size=(nb, M)) < 0.4
    want = np.concatenate([[ll.sum(), llo.sum()], acc.sum(axis=1)])
    assert res[0][1] == 0 and res[-1][2] == M and all(res[i][2] == res[i + 1][1] for i in range(world - 1))
    for rank, lo, hi, tot, same in res:
        assert same
        assert np.allclose(tot, want, rtol=1e-13, atol=1e-12)


def test_shard_slice_tiles_any_size():
    for n in (1, 7, 8, 4096, 4097):
        for w in (1, 2, 3, 8):
            sl = [H.shard_slice(n, r, w) for r in range(w)]
            assert sl[0][0] == 0 and sl[-1][1] == n and all(a[1] == b[0] for a, b in zip(sl, sl[1:]))
            assert max(h - l for l, h in sl) - min(h - l for l, h in sl) <= 1
