#!/usr/bin/env python
"""Throughput of the UME descriptor-and-registration hot path (BASELINE.json metric: pairs/sec,
KITTI-shape pairs, 1024 keypoints) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W]              # this repo's CUDA path
    python bench.py --impl reference [--gpus N] ...                  # the CPU reference arm

A step = one pass of the hot path (evaluate.py:206-257: UME matrices of both clouds, all-pairs
subspace distance + arg-min, one rigid hypothesis per match) over one batch of synthetic pairs.
Workloads (`--workload`):
  kitti_b64_n1024_c32 (default)  BASELINE config #3: 64 KITTI-shape pairs x 120k points per GPU, 1024
                                 keypoints, 32 channels; weak scaling (every rank owns its own 64 pairs)
  kitti_b1_n512_c32              config #2: one pair, 512 keypoints (CUDA-graph replay)
  nuscenes_b512_n1024_c32        config #4: 512 nuScenes-shape pairs (35k points) in total, sharded over
                                 the ranks (strong scaling), micro-steps of 64 pairs
  rotkitti_stream4096_n2048_c64  config #5: 4096 pairs in total STREAMED from pinned host memory in
                                 micro-batches, 2048 keypoints, 64 channels, ground truth cycled from
                                 the reference's RotKITTI transforms (tests/golden/rotkitti_gt_tforms.npy),
                                 large-rotation recovery checked against it
Every step ends with ONE NCCL all-gather of the packed per-pair results (N > 1).
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def _host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


if "--impl" in sys.argv and "reference" in sys.argv:
    # the CPU arm uses every host core it may run on, also under torchrun (which exports
    # OMP_NUM_THREADS=1 to its workers): set before numpy / torch / libgomp initialise
    for _v in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
        os.environ[_v] = str(_host_threads())

import numpy as np  # noqa: E402

WORKLOADS = {
    # pairs: per GPU per step (weak) or in total (strong / stream); micro: pairs per kernel batch
    "kitti_b64_n1024_c32": dict(pairs=64, N=120000, n_kp=1024, C=32, model="KITTI", scaling="weak", micro=64),
    "kitti_b1_n512_c32": dict(pairs=1, N=120000, n_kp=512, C=32, model="KITTI", scaling="weak", micro=1),
    "nuscenes_b64_n1024_c32": dict(pairs=64, N=35000, n_kp=1024, C=32, model="NUSCENES", scaling="weak", micro=64),
    "nuscenes_b512_n1024_c32": dict(pairs=512, N=35000, n_kp=1024, C=32, model="NUSCENES", scaling="strong", micro=64, pool=4),
    "rotkitti_b32_n2048_c64": dict(pairs=32, N=120000, n_kp=2048, C=64, model="KITTI", scaling="weak", micro=32),
    "rotkitti_stream4096_n2048_c64": dict(pairs=4096, N=120000, n_kp=2048, C=64, model="KITTI", scaling="strong",
                                          micro=16, stream=True, pool=2, gt="rotkitti"),
    "tiny": dict(pairs=4, N=20000, n_kp=256, C=32, model="KITTI", scaling="weak", micro=4),
    "tiny_stream": dict(pairs=16, N=20000, n_kp=256, C=32, model="KITTI", scaling="strong", micro=2, stream=True,
                        pool=2, gt="rotkitti"),
}
K_NN, RADIUS = 750, 5.0
METRIC, UNIT = "pairs_per_sec_kitti_shape_1024kp", "pairs/s"
KEYS = ("src_pts", "src_feat", "src_kp", "tgt_pts", "tgt_feat", "tgt_kp")

_JSON_FD = None


def protect_stdout():
    """Libraries (NCCL prints its version banner) write to fd 1; the driver wants exactly ONE JSON
    line there.  Everything else is sent to stderr; the JSON line goes to the saved descriptor."""
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="kitti_b64_n1024_c32", choices=sorted(WORKLOADS))
    ap.add_argument("--cdist-impl", type=int, default=None, help="0 = SIMT fp32, 1 = tcgen05")
    ap.add_argument("--cell-div2", type=int, default=None)
    ap.add_argument("--cta-moments", type=int, default=None, help="1 = force the CTA-per-keypoint moment kernel")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-baseline-pairs", type=int, default=16)
    ap.add_argument("--ref-pairs-per-step", type=int, default=2, help="--impl reference: pairs of the workload per step")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--full-reg-pairs", type=int, default=4, help="pairs for the full-registration (with hypothesis selection) figure; 0 = skip")
    ap.add_argument("--chunk-pairs", type=int, default=8)
    ap.add_argument("--numa-bind", type=int, default=1)
    ap.add_argument("--graph", type=int, default=None, help="replay the step from a CUDA graph (default: on for <= 8 pairs)")
    return ap.parse_args()


def base_config(name, wl):
    """The part of `config` both arms share (what the driver compares between them)."""
    per = "per GPU per step" if wl["scaling"] == "weak" else "in total per step, sharded over the GPUs"
    inputs_gb = wl["pairs"] * 2 * wl["N"] * (wl["C"] + 3) * 4 / 1e9
    return {"workload": name, "pairs": wl["pairs"], "pairs_are": per, "points": wl["N"], "keypoints": wl["n_kp"],
            "channels": wl["C"], "K": K_NN, "radius": RADIUS, "lidar_model": wl["model"],
            "l2_policy": "inputs (%.1f GB per step) exceed the 126 MB L2" % inputs_gb}


def rotkitti_gt():
    """The reference's RotKITTI ground-truth transforms (600 x 4 x 4, rotations 28.8-180 deg), a fixture
    copied from datasets/kitti/metadata/rotkitti_gt_tforms.npy."""
    return np.load(os.path.join(REPO, "tests", "golden", "rotkitti_gt_tforms.npy")).astype(np.float64)


def make_pairs(wl, n_pairs, seed0):
    """n_pairs synthetic pairs of the workload as stacked numpy arrays (+ 'gt')."""
    from umeregrobust_b200 import synth
    model = getattr(synth, wl["model"])
    kw = dict(N=wl["N"], C=wl["C"], n_kp=wl["n_kp"], model=model)
    if wl.get("gt") != "rotkitti":
        return synth.make_batch(n_pairs, seed0=seed0, n_base=min(4, n_pairs), **kw)
    gts = rotkitti_gt()
    # ground truth is checked on these pairs: features are a smooth field (as a backbone's output is) and the
    # target keypoints are the source keypoints seen in the other cloud, see synth.make_pair
    kw.update(feat_model="field", kp_mode="corresponding")
    bases = [synth.make_pair(seed0 + i, **kw) for i in range(min(2, n_pairs))]
    pairs = []
    for p in range(n_pairs):
        base = bases[p % len(bases)]
        want = gts[(seed0 + p) % len(gts)]
        extra = want @ np.linalg.inv(base["gt"].astype(np.float64))      # extra @ base.gt == want
        pairs.append(synth.rederive_pair(base, seed0 + 1000 + p, n_kp=wl["n_kp"], gt_extra=extra, kp_mode="corresponding"))
    return {k: np.stack([q[k] for q in pairs], 0) for k in pairs[0]}


# ----------------------------------------------------------------------------- CPU reference leg
class CpuArm:
    """The reference's CPU implementation of the path on the host cores.

    kind "reference": the reference's OWN functions (`evaluate.my_ume_generation` -> `utils.loc_utils.ume_cdist`
    -> the arg-min / gather lines of evaluate.py:224-231 -> `batch_estimate_transform_ume_old`), imported from the
    staged copy under baseline/_ref through the stubs of oracle/ref_import.py, with `pytorch3d.ops.ball_query`
    bound to the OpenMP C restatement (pytorch3d itself is not installable offline).  The reference hard-codes
    32 channels (evaluate.py:55,230-231), so other channel counts — and boxes without the staged copy — fall back
    to kind "port": the oracle's numpy restatement of the same four steps."""

    def __init__(self, C, threads):
        self.threads = int(threads)
        self.kind, self.note = "port", "oracle port: numpy + OpenMP C ball_query"
        from oracle import pytorch3d_ops as p3d
        p3d._load()
        self.p3d = p3d
        if C == 32:
            try:
                import torch
                from oracle import ref_import
                if ref_import.reference_available():
                    torch.set_num_threads(self.threads)
                    self.ev, self.loc, _ = ref_import.import_reference(num_threads=self.threads)
                    self.torch = torch
                    self.kind = "reference"
                    self.note = ("the reference's own my_ume_generation / ume_cdist / arg-min / batch_estimate_transform_ume_old "
                                 "(%s copy; torch %s CPU, %d threads; ball_query = OpenMP C restatement of pytorch3d's)"
                                 % (ref_import.reference_kind(), torch.__version__, torch.get_num_threads()))
            except Exception as e:                                   # staged copy missing / import error: port
                self.note += " (reference import failed: %s)" % (str(e)[:80],)

    def run(self, batch, idxs):
        """Seconds for pairs `idxs` of `batch` (numpy arrays)."""
        t0 = time.perf_counter()
        if self.kind == "reference":
            from types import SimpleNamespace
            torch, ev, loc = self.torch, self.ev, self.loc
            args = SimpleNamespace(ume_max_nn=K_NN, ume_r_nn=RADIUS)
            with torch.no_grad():
                for p in idxs:
                    t = {k: torch.from_numpy(batch[k][p:p + 1]) for k in KEYS}
                    ume_src = ev.my_ume_generation(t["src_pts"], t["src_kp"], t["src_feat"], args)        # evaluate.py:206
                    ume_tgt = ev.my_ume_generation(t["tgt_pts"], t["tgt_kp"], t["tgt_feat"], args)        # :207
                    D = loc.ume_cdist(ume_src, ume_tgt)                                                   # :215
                    m = D.min(dim=-1)[1]                                                                  # :224
                    m = torch.cat([torch.arange(D.shape[1])[None, :, None], m[..., None]], dim=-1)        # :225
                    ume_t = torch.gather(ume_tgt, 1, m[..., 1][..., None, None].expand(-1, -1, 32, 4))    # :230
                    ume_s = torch.gather(ume_src, 1, m[..., 0][..., None, None].expand(-1, -1, 32, 4))    # :231
                    G, H = ume_s.unsqueeze(2), ume_t.unsqueeze(1)                                         # :248-252
                    G, H = G.reshape(-1, *G.shape[3:]), H.reshape(-1, *H.shape[3:])
                    loc.batch_estimate_transform_ume_old(G, H)                                            # :253
        else:
            from oracle import ume_oracle as orc
            for p in idxs:
                orc.register_pair_hypotheses(batch["src_pts"][p:p + 1], batch["src_feat"][p:p + 1], batch["src_kp"][p:p + 1],
                                             batch["tgt_pts"][p:p + 1], batch["tgt_feat"][p:p + 1], batch["tgt_kp"][p:p + 1],
                                             K_NN, RADIUS, dtype=np.float32)
        return time.perf_counter() - t0

    def describe(self, value, sample):
        return {"value": value, "unit": UNIT, "cores": self.threads, "kind": self.kind, "sample": sample, "how": self.note}


def run_reference(args, name, wl):
    """--impl reference: W warm-up steps and K timed steps like the CUDA arm; a step is a bounded
    sample (`--ref-pairs-per-step` pairs) of the same workload, so the run ends within minutes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = _host_threads()
    arm = CpuArm(wl["C"], threads)
    per_step = max(1, min(args.ref_pairs_per_step, wl["pairs"]))
    batch = make_pairs(wl, per_step, seed0=0)
    idxs = list(range(per_step))
    for _ in range(args.warmup):
        arm.run(batch, idxs)
    times = [arm.run(batch, idxs) for _ in range(args.steps)]
    sec = float(np.mean(times))
    value = per_step / sec
    sample = "%d pair(s) of the workload per step, %d warm-up + %d timed steps" % (per_step, args.warmup, args.steps)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
            "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": base_config(name, wl),
            "cpu_baseline": arm.describe(value, sample),
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ----------------------------------------------------------------------------- CUDA arm
def algorithmic_bytes(counts, n_kp, C):
    """per cloud  sum_i cnt_i (4C + 12) + n (12 + 16C)."""
    return float(counts.sum()) * (4 * C + 12) + counts.shape[0] * n_kp * (12 + 16 * C)


def measure_dense_peak(torch, dev, dtype="tf32"):
    """Dense tensor-core peak the way MEASURED_PEAKS.json measures bf16: cuBLAS matmul 8192^3, best of 10.
    dtype "tf32" (fp32 operands, TF32 math) or "f16"."""
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        dt = torch.float16 if dtype == "f16" else torch.float32
        a = torch.randn(8192, 8192, device=dev, dtype=dt)
        b = torch.randn(8192, 8192, device=dev, dtype=dt)
        best = 1e9
        for i in range(12):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b)
            e1.record()
            e1.synchronize()
            if i >= 2:
                best = min(best, e0.elapsed_time(e1))
        return 2.0 * 8192 ** 3 / (best * 1e-3) / 1e12
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


def measure_h2d(torch, dev, barrier, nbytes=1 << 30):
    """Raw pinned host->device copy bandwidth of this rank while every rank copies at once (GB/s)."""
    h = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    h.zero_()
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    d.copy_(h, non_blocking=True)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        d.copy_(h, non_blocking=True)
    e1.record()
    e1.synchronize()
    return 3 * nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9


def gt_check(torch, T, gt, rot_deg=1.5, trans_m=0.6):
    """Large-rotation recovery: per pair, the share of hypotheses T (P,m,4,4) within (rot_deg, trans_m) of
    the ground truth gt (P,4,4) (the reference's N.P thresholds, evaluate.py:304)."""
    import umeregrobust_b200 as ume
    P, m = T.shape[0], T.shape[1]
    R_gt = gt[:, None, :3, :3].expand(P, m, 3, 3).reshape(-1, 3, 3).contiguous()
    rre = ume.relative_rotation_error(R_gt, T.reshape(-1, 4, 4)[:, :3, :3]).view(P, m)
    rte = (T[..., :3, 3] - gt[:, None, :3, 3]).norm(dim=-1)
    ok = (rre <= rot_deg) & (rte <= trans_m)
    frac = ok.float().mean(dim=1)
    return {"pairs_checked": int(P), "thresholds": [rot_deg, trans_m],
            "pairs_with_a_correct_hypothesis": int((frac > 0).sum().item()),
            "median_share_of_correct_hypotheses": float(frac.median().item()),
            "gt_rotation_deg_min_max": None}


def run_b200(args, name, wl):
    import torch
    import torch.distributed as dist
    import umeregrobust_b200 as ume
    from umeregrobust_b200 import _lib
    from umeregrobust_b200.engine import RegistrationEngine, PackedPairs, bind_to_gpu_numa_node, shard_range

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(0)
    dev_index = local if world > 1 else 0
    dev = torch.device("cuda", dev_index)
    numa = bind_to_gpu_numa_node(dev_index) if args.numa_bind else None
    _lib.lib()
    if args.cdist_impl is not None:
        ume.config["cdist_impl"] = args.cdist_impl
    if args.cell_div2 is not None:
        ume.config["cell_div2"] = bool(args.cell_div2)
    if args.cta_moments is not None:
        ume.config["cta_moments"] = bool(args.cta_moments)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- the rank's share of the step and its synthetic data
    strong = wl["scaling"] == "strong"
    stream = bool(wl.get("stream"))
    if strong:
        lo, hi = shard_range(wl["pairs"], rank, world)
        my_pairs = hi - lo
    else:
        my_pairs = wl["pairs"]
    micro = min(wl["micro"], max(my_pairs, 1))
    n_micro = (my_pairs + micro - 1) // micro
    pool_n = min(wl.get("pool", n_micro), n_micro) if (strong or stream) else 1
    pool_np = [make_pairs(wl, micro, seed0=10000 * rank + 100 * j) for j in range(pool_n)]
    pool_dev = [{k: torch.from_numpy(b[k]).to(dev) for k in KEYS} for b in pool_np]
    gt_dev = [torch.from_numpy(b["gt"]).to(dev) for b in pool_np]
    pairs_per_step = n_micro * micro                      # (a strong shard is padded to whole micro-batches)
    job_pairs = pairs_per_step * world
    eng = RegistrationEngine(K=K_NN, radius=RADIUS, device=dev, want_D=not stream, chunk_pairs=args.chunk_pairs)
    use_graph = (micro <= 8 and not strong) if args.graph is None else bool(args.graph)

    def step():
        pack = None
        for j in range(n_micro):
            b = pool_dev[j % pool_n]
            if world > 1:
                pack = eng.register_and_gather(b)
            elif use_graph:
                eng.register_graphed(b)
            else:
                eng.register(b, slot=j % 2)
        return pack

    for _ in range(max(args.warmup, 3)):
        step()
    if world > 1:
        eng.finish_gathers()
    barrier()

    # neighbour counts (outside the timed region) -> algorithmic bytes of the gather+moment kernel
    b0 = pool_dev[0]
    _, cnt_s = ume.ume_moments(b0["src_pts"], b0["src_kp"], b0["src_feat"], K_NN, RADIUS, return_count=True)
    _, cnt_t = ume.ume_moments(b0["tgt_pts"], b0["tgt_kp"], b0["tgt_feat"], K_NN, RADIUS, return_count=True)
    cnt_s, cnt_t = cnt_s.cpu().numpy(), cnt_t.cpu().numpy()
    bytes_both_sides = algorithmic_bytes(cnt_s, wl["n_kp"], wl["C"]) + algorithmic_bytes(cnt_t, wl["n_kp"], wl["C"])   # one micro-batch

    from umeregrobust_b200.clocks import ClockSampler
    try:
        dev_uuid = str(torch.cuda.get_device_properties(dev).uuid)
    except Exception:
        dev_uuid = None
    sampler = ClockSampler(dev_index, uuid=dev_uuid)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    _lib.profile_reset()
    _lib.profile_enable(True)
    barrier()
    launches0 = _lib.launch_count()
    t_host0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    if world > 1:
        eng.finish_gathers()
    e1.record()
    barrier()
    t_host1 = time.time()
    launches = _lib.launch_count() - launches0
    _lib.profile_enable(False)
    prof = _lib.profile_read()
    elapsed_ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    clocks = sampler.stop(t_host0, t_host1) if rank == 0 else None
    ms_per_step = elapsed_ms / args.steps
    value = job_pairs / (ms_per_step * 1e-3)

    # ---- large-rotation ground-truth check (config #5) on the last registered micro-batch of the pool
    gtc = None
    if wl.get("gt") == "rotkitti":
        out = eng.register(pool_dev[0], slot=0)
        gtc = gt_check(torch, out["T"], gt_dev[0])
        ang = np.degrees(np.arccos(np.clip((np.trace(pool_np[0]["gt"][:, :3, :3], axis1=1, axis2=2) - 1) / 2, -1, 1)))
        gtc["gt_rotation_deg_min_max"] = [float(ang.min()), float(ang.max())]

    # ---- end to end through the public API with HOST buffers (pinned), H2D + D2H inside the timed region
    e2e = None
    if args.e2e_steps:
        h2d_peak = measure_h2d(torch, dev, barrier)
        if stream or strong:
            host_pool = [PackedPairs.from_arrays(b) for b in pool_np]
            host_batches = [host_pool[j % pool_n] for j in range(n_micro)]
            run_host = lambda: eng.register_stream(host_batches, total_pairs=pairs_per_step)   # noqa: E731
            h2d, d2h = eng.host_bytes(host_batches)
            how = "register_stream: %d pinned micro-batches of %d pairs, one H2D copy each, two streams" % (n_micro, micro)
        else:
            host = {k: torch.from_numpy(pool_np[0][k]).pin_memory() for k in KEYS}
            run_host = lambda: eng.register_host(host)                                          # noqa: E731
            h2d, d2h = eng.host_bytes(host)
            how = "register_host: chunks of %d pairs alternating on two streams" % args.chunk_pairs
        for _ in range(2):
            run_host()
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(args.e2e_steps):
            res = run_host()
            if world > 1:                                  # the step's single collective: per-pair results to every rank
                from umeregrobust_b200.engine import gather_results
                gather_results({k: v.to(dev, non_blocking=True) for k, v in res.items()})
        s1.record()
        barrier()
        e2e_ms = s0.elapsed_time(s1)
        stats = torch.tensor([e2e_ms, -h2d_peak, h2d_peak], device=dev, dtype=torch.float64)
        if world > 1:
            mx = stats.clone()
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            sm = stats.clone()
            dist.all_reduce(sm, op=dist.ReduceOp.SUM)
            e2e_ms, h2d_min, h2d_sum = float(mx[0].item()), -float(mx[1].item()), float(sm[2].item())
        else:
            h2d_min = h2d_sum = h2d_peak
        per_step_ms = e2e_ms / args.e2e_steps
        e2e = {"value": job_pairs / (per_step_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d * world,
               "d2h_bytes_per_step": d2h * world, "steps": args.e2e_steps, "ms_per_step": per_step_ms,
               "h2d_gbs": h2d * world / (per_step_ms * 1e-3) / 1e9,
               "h2d_peak_gbs": h2d_sum, "h2d_peak_gbs_slowest_gpu": h2d_min,
               "h2d_peak_how": "1 GiB pinned->device copies, all %d rank(s) copying at once, CUDA events" % world,
               "how": how, "numa": numa}

    # ---- full registration: hypothesis selection included, ONE (R,t) per pair (north_star's per-pair result)
    full = None
    if args.full_reg_pairs > 0 and wl["C"] in (32, 64):
        # pairs with a smooth feature field and RotKITTI ground truth, so that the selected (R,t) can be checked
        fp = max(1, min(args.full_reg_pairs, micro))
        fnp = make_pairs(dict(wl, gt="rotkitti"), fp, seed0=777 + 10000 * rank)
        fb = {k: torch.from_numpy(fnp[k]).to(dev) for k in KEYS}
        feng = RegistrationEngine(K=K_NN, radius=RADIUS, device=dev, select=True, corr_sigma=1.5, pc_corr_max_size=10000)
        feng.register(fb)
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        _lib.profile_reset()
        _lib.profile_enable(True)
        f0.record()
        fout = feng.register(fb)
        if world > 1:
            from umeregrobust_b200.engine import gather_results
            gather_results({"T_best": fout["T_best"], "best": fout["best"]})       # 64 B + 8 B per pair
        f1.record()
        barrier()
        _lib.profile_enable(False)
        fprof = _lib.profile_read()
        f_ms = f0.elapsed_time(f1)
        if world > 1:
            t = torch.tensor([f_ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            f_ms = float(t.item())
        g = torch.from_numpy(fnp["gt"]).to(dev)
        rre = ume.relative_rotation_error(g[:, :3, :3], fout["T_best"][:, :3, :3])
        rte = (fout["T_best"][:, :3, 3] - g[:, :3, 3]).norm(dim=-1)
        gt_best = {"rre_deg": [float(x) for x in rre.cpu()], "rte_m": [float(x) for x in rte.cpu()]}
        full = {"value": world * fp / (f_ms * 1e-3), "unit": UNIT, "pairs_per_gpu": fp, "ms_per_pair": f_ms / fp,
                "hypotheses_per_pair": wl["n_kp"], "corr_points": 10000, "corr_num_nn": 20,
                "corr_ms_per_pair": fprof["corr"][0] / fp, "knn_ms_per_pair": fprof["knn"][0] / fp,
                "result_bytes_per_pair": 72, "selected_vs_gt": gt_best,
                "what": "evaluate.py:206-296: hypotheses + voxel de-duplication, feature transfer, down-sampling and the "
                        "FeatureCorrelator pick; gathers one (R,t) per pair"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks, peaks_note = {}, "fallback"
    try:
        peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
        peaks_note = "measured"
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    mom_ms, mom_n = prof["moments"]
    mom_avg_ms = mom_ms / max(mom_n, 1)
    # launches per micro-batch: 1 when source and target share one launch (ume_moments_pair_f32), else 2
    mom_per_micro = max(1, int(round(mom_n / float(args.steps * n_micro)))) if mom_n else 1
    bytes_per_launch = bytes_both_sides / mom_per_micro
    achieved = bytes_per_launch / (mom_avg_ms * 1e-3) / 1e9 if mom_n else None
    traffic = l2_bytes = None
    try:
        tr = json.load(open(os.path.join(REPO, "profiles", "moments_dram_traffic.json")))
        if tr.get("workload") == name and int(tr.get("launches_per_micro_batch", 2)) == mom_per_micro:
            traffic = tr.get("dram_bytes_per_launch")
            l2_bytes = tr.get("l2_bytes_per_launch")
    except Exception:
        pass
    stages = {k: {"ms_per_step": v[0] / args.steps, "launches_per_step": v[1] / args.steps}
              for k, v in prof.items() if v[1]}
    cd_ms, cd_n = prof["cdist"]
    cd_avg_ms = cd_ms / max(cd_n, 1)
    gemm_flops = 2.0 * (4 * wl["n_kp"]) ** 2 * wl["C"] * micro          # SURVEY §8d, per launch (one micro-batch)
    f16_peak = measure_dense_peak(torch, dev, "f16")
    tf32_peak = measure_dense_peak(torch, dev, "tf32")
    cfg = base_config(name, wl)
    detail = {}
    detail.update({"pairs_per_gpu_per_step": pairs_per_step, "pairs_per_kernel_batch": micro,
                "mean_neighbours": float(0.5 * (cnt_s.mean() + cnt_t.mean())),
                "cdist_impl": ume.config["cdist_impl"], "cell_div2": ume.config["cell_div2"], "cuda_graph": use_graph,
                "moment_kernel": "cta-per-keypoint" if ume.config["cta_moments"] else "warp-per-keypoint",
                "parallelism": "pairs sharded over %d GPU(s), one all-gather of the packed per-pair results per kernel batch "
                               "on a side stream" % world})
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": wl["scaling"],
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg,                 # identical in both arms (`--impl reference` prints the same dict)
        "detail": detail,
        "e2e": e2e,
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"kernel": "moments_%s_kernel (fused gather + UME moments)" % ("cta" if ume.config["cta_moments"] else "warp"),
                     "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                     "frac": (achieved / hbm_peak) if achieved else None, "traffic": traffic,
                     "dram_frac": (traffic / (mom_avg_ms * 1e-3) / 1e9 / hbm_peak) if (traffic and mom_n) else None,
                     "l2_gbs": (l2_bytes / (mom_avg_ms * 1e-3) / 1e9) if (l2_bytes and mom_n) else None,
                     "launch_covers": "source + target batch" if mom_per_micro == 1 else "one side",
                     "algorithmic_bytes_per_launch": bytes_per_launch, "avg_launch_ms": mom_avg_ms,
                     "launches_timed": mom_n, "peak_source": peaks_note + " (MEASURED_PEAKS.json hbm_gbs)",
                     "note": "frac counts SURVEY §8d's algorithmic bytes, most of which are served L2->SM (a cloud's features, "
                             "15 MB, stay in the 126 MB L2); dram_frac = measured DRAM bytes per launch (ncu, `traffic`) / time / peak "
                             "is the HBM-side fraction"},
        "roofline_gemm": {"kernel": "cdist_tc_kernel (all-pairs subspace distance, Gram form, tcgen05 kind::f16) + fused arg-min",
                          "bound": "tensor", "achieved": gemm_flops / (cd_avg_ms * 1e-3) / 1e12 if cd_n else None, "peak": f16_peak,
                          "unit": "TFLOP/s", "frac": (gemm_flops / (cd_avg_ms * 1e-3) / 1e12 / f16_peak) if cd_n else None,
                          "executed_frac": (3.0 * gemm_flops / (cd_avg_ms * 1e-3) / 1e12 / f16_peak) if cd_n else None,
                          "algorithmic_flops_per_launch": gemm_flops, "avg_launch_ms": cd_avg_ms, "launches_timed": cd_n,
                          "executed_over_algorithmic": 3.0, "tf32_peak": tf32_peak,
                          "tensor_pipe_active_pct_ncu": 39.8,
                          "peak_source": "measured live: cuBLAS fp16 matmul 8192^3, best of 10 (MEASURED_PEAKS.json bf16: %.0f); "
                                         "tf32_peak: the same with TF32 math" % float(peaks.get("bf16_tflops", 0.0)),
                          "note": "algorithmic = 2 (4n)^2 C per pair; every fp32 operand is split into two fp16 numbers and the kernel "
                                  "executes 3 fp16 products per term (hi*hi + hi*lo + lo*hi) for fp32-grade distances; K = C is tiny, "
                                  "so the kernel is bound by draining the accumulators (squares, 4x4 block sums, sqrt, D store, arg-min "
                                  "read every accumulator once from TMEM), not by the MMAs; tensor_pipe_active from "
                                  "profiles/r02_cdist_kernel_full.csv"},
        "stages": stages,
    }
    if gtc is not None:
        line["gt_check"] = gtc
    if full is not None:
        line["full_registration"] = full
    if not args.no_cpu_baseline and world == 1:          # the contract: rank 0 at N = 1 only
        arm = CpuArm(wl["C"], _host_threads())
        n_cpu = max(1, min(args.cpu_baseline_pairs, micro))
        arm.run(pool_np[0], [0])                                        # warm-up pair
        sec = arm.run(pool_np[0], list(range(n_cpu)))
        line["cpu_baseline"] = arm.describe(n_cpu / sec, "%d pairs of the same batch, %.1f s" % (n_cpu, sec))
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    protect_stdout()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, args.workload, wl)
    else:
        run_b200(args, args.workload, wl)


if __name__ == "__main__":
    main()
