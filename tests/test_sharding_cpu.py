"""CPU, world_size 2 over gloo: pair sharding and the single end-of-step gather of per-pair
results.  The data path has no collective; this is the whole multi-GPU logic."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from umeregrobust_b200.engine import gather_results, shard_range


def test_shard_range_covers_all_pairs_contiguously():
    for n_pairs in (0, 1, 7, 64, 512, 4096, 4097):
        for world in (1, 2, 3, 8):
            blocks = [shard_range(n_pairs, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n_pairs
            for (a0, a1), (b0, b1) in zip(blocks, blocks[1:]):
                assert a1 == b0 and a0 <= a1
            per = -(-n_pairs // world) if n_pairs else 0
            assert all(b1 - b0 <= per for b0, b1 in blocks)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_pairs, n_kp, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(n_pairs, rank, world)
    g = torch.Generator().manual_seed(1234)                      # every rank can rebuild the full job
    T_all = torch.randn(n_pairs, n_kp, 4, 4, generator=g)
    m_all = torch.randint(0, n_kp, (n_pairs, n_kp), generator=g)
    d_all = torch.rand(n_pairs, n_kp, generator=g)
    local = dict(T=T_all[lo:hi].clone(), match=m_all[lo:hi].clone(), dmin=d_all[lo:hi].clone())
    full = gather_results(local)
    ok = (torch.equal(full["T"], T_all) and torch.equal(full["match"], m_all) and torch.equal(full["dmin"], d_all))
    # max-over-ranks timing reduction used by bench.py
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok = ok and float(t.item()) == float(world)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_gather_results_world2_gloo():
    world, n_pairs, n_kp = 2, 6, 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_pairs, n_kp, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]
