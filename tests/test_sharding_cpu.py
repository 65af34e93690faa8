"""CPU, world_size 2 over gloo: pair sharding and the single end-of-step gather of per-pair
results.  The data path has no collective; this is the whole multi-GPU logic."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from umeregrobust_b200.engine import PackedPairs, ResultPack, gather_packed, gather_results, shard_range


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
    # the engine's single-collective path: every rank fills its ResultPack.local, ONE all-gather
    pack = ResultPack(hi - lo, n_kp, "cpu", world=world)
    v = pack.views()
    v["T"].copy_(T_all[lo:hi])
    v["argmin"].copy_(m_all[lo:hi])
    v["dmin"].copy_(d_all[lo:hi])
    v["best"].copy_(torch.arange(lo, hi))
    av = gather_packed(pack)
    ok = ok and torch.equal(av["T"].reshape(n_pairs, n_kp, 4, 4), T_all) and torch.equal(av["argmin"].reshape(n_pairs, n_kp), m_all)
    ok = ok and torch.equal(av["dmin"].reshape(n_pairs, n_kp), d_all) and torch.equal(av["best"].reshape(-1), torch.arange(n_pairs))
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


def test_packed_buffers_layout():
    rng = np.random.default_rng(0)
    b = {k: rng.normal(size=(3, 50, 3 if "feat" not in k else 8)).astype(np.float32) for k in ("src_pts", "src_feat", "tgt_pts", "tgt_feat")}
    b["src_kp"] = rng.normal(size=(3, 7, 3)).astype(np.float32)
    b["tgt_kp"] = rng.normal(size=(3, 7, 3)).astype(np.float32)
    p = PackedPairs.from_arrays(b, device="cpu")
    assert (p.pairs, p.N, p.n, p.C) == (3, 50, 7, 8)
    assert all(np.array_equal(p[k].numpy(), b[k]) for k in b)
    assert all(p[k].data_ptr() % 256 == p.raw.data_ptr() % 256 for k in b)        # 256-byte aligned sections
    assert p.payload_bytes() == sum(v.nbytes for v in b.values()) <= p.nbytes
    # one allocation: a copy of `raw` carries all six arrays
    q = PackedPairs(3, 50, 7, 8, device="cpu")
    q.raw.copy_(p.raw)
    assert all(torch.equal(q[k], p[k]) for k in b)
    r = ResultPack(5, 11, "cpu", world=3)
    v = r.views()
    assert tuple(v["T"].shape) == (5, 11, 4, 4) and v["argmin"].dtype == torch.int64 and tuple(v["T_best"].shape) == (5, 4, 4)
    v["dmin"].fill_(2.5)
    r.raw.view(3, r.nbytes)[1].copy_(r.local)
    av = r.all_views()
    assert tuple(av["dmin"].shape) == (3, 5, 11) and float(av["dmin"][1].min()) == 2.5 and float(av["dmin"][0].max()) == 0.0


def test_reference_arm_uses_every_core_under_torchrun(tmp_path):
    """torchrun exports OMP_NUM_THREADS=1 to its workers; the reference arm must
    still run on all host cores (rank 0 only; the other rank exits 0 without work) and report them."""
    import json
    import subprocess
    import sys
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(repo, "bench.py"), "--impl", "reference", "--gpus", "2",
           "--steps", "1", "--warmup", "1", "--workload", "tiny"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=repo)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    cores = len(os.sched_getaffinity(0))
    assert line["impl"] == "reference" and line["cpu_baseline"]["cores"] == cores
    assert line["cpu_baseline"]["kind"] in ("reference", "port")
    assert line["steps"] == 1 and line["warmup"] == 1 and line["n_gpus"] == 2
    assert "%d threads" % cores in line["cpu_baseline"]["how"] or line["cpu_baseline"]["kind"] == "port"
