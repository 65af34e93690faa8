"""Batched registration engine: the call a user makes for many pairs at once.

`register(batch)` runs evaluate.py:206-257 (UME generation for both clouds, subspace distances with
fused arg-min, optional distance-weighted sub-sampling of the matches, one rigid hypothesis per
match) for a batch of pairs that is already resident on the device; `select=True` adds
evaluate.py:259-296 (hypothesis selection by feature correlation) and returns ONE (R,t) per pair.
`register_host(batch)` takes PINNED HOST buffers: the batch is cut into chunks that alternate between
two CUDA streams, so the host->device copy of one chunk overlaps the kernels of the previous one, and
results come back into pinned host buffers.  `register_stream(micro_batches)` is the same pipeline
over an iterator of host micro-batches (BASELINE config #5: thousands of pairs streamed through).

Multi-GPU: pairs are independent (no cross-pair state anywhere in evaluate.py:175-299), so ranks
take contiguous blocks of pairs (`shard_range`); the per-pair results of a step live in ONE packed
buffer that the kernels write directly (`ResultPack`), and `gather_step` all-gathers that buffer with
a single collective on a side stream, overlapped with the next step's kernels.
"""
import os

import torch

from . import api

_IN_KEYS = ("src_pts", "src_feat", "src_kp", "tgt_pts", "tgt_feat", "tgt_kp")


def _align(x, a=256):
    return (x + a - 1) // a * a


# ----------------------------------------------------------------------------- packed buffers
class PackedPairs:
    """The six input arrays of `pairs` registration pairs in ONE allocation (256-byte aligned
    sections), pinned on the host or resident on a device: a micro-batch moves host->device with a
    single copy instead of six.  `views[k]` are ordinary tensors into the allocation."""

    def __init__(self, pairs, N, n, C, device="pinned"):
        self.pairs, self.N, self.n, self.C = int(pairs), int(N), int(n), int(C)
        shapes = dict(src_pts=(pairs, N, 3), src_feat=(pairs, N, C), src_kp=(pairs, n, 3),
                      tgt_pts=(pairs, N, 3), tgt_feat=(pairs, N, C), tgt_kp=(pairs, n, 3))
        off, self._layout = 0, {}
        for k in _IN_KEYS:
            nbytes = 4 * int(torch.Size(shapes[k]).numel())
            self._layout[k] = (off, nbytes, shapes[k])
            off = _align(off + nbytes)
        self.nbytes = off
        if device == "pinned":
            self.raw = torch.empty(self.nbytes, dtype=torch.uint8, pin_memory=torch.cuda.is_available())
        else:
            self.raw = torch.empty(self.nbytes, dtype=torch.uint8, device=device)
        self.views = {k: self.raw[o:o + b].view(torch.float32).view(shape) for k, (o, b, shape) in self._layout.items()}

    @classmethod
    def from_arrays(cls, batch, device="pinned"):
        """batch: dict of numpy arrays / CPU tensors with the six keys, shapes (B, ...)."""
        t = {k: torch.as_tensor(batch[k]) for k in _IN_KEYS}
        B, N, _ = t["src_pts"].shape
        p = cls(B, N, t["src_kp"].shape[1], t["src_feat"].shape[2], device=device)
        for k in _IN_KEYS:
            p.views[k].copy_(t[k])
        return p

    def payload_bytes(self):
        return sum(b for _, b, _ in self._layout.values())

    def __getitem__(self, k):
        return self.views[k]


class ResultPack:
    """Per-step results of `pairs` pairs with `m` hypotheses each in ONE buffer:
    [T (pairs,m,4,4) f32 | argmin (pairs,m) i64 | dmin (pairs,m) f32 | T_best (pairs,4,4) f32 | best (pairs) i64].
    The rigid-solve and distance kernels write straight into these views (they are pre-seeded into
    the arena the kernels take their outputs from), so the end-of-step collective is one all-gather of
    `raw` with no packing pass."""

    FIELDS = (("T", torch.float32, lambda p, m: (p, m, 4, 4)), ("argmin", torch.int64, lambda p, m: (p, m)),
              ("dmin", torch.float32, lambda p, m: (p, m)), ("T_best", torch.float32, lambda p, m: (p, 4, 4)),
              ("best", torch.int64, lambda p, m: (p,)))

    def __init__(self, pairs, m, device, world=1):
        self.pairs, self.m, self.world = int(pairs), int(m), int(world)
        off, self._layout = 0, {}
        for name, dt, shp in self.FIELDS:
            shape = shp(self.pairs, self.m)
            nbytes = int(torch.Size(shape).numel()) * torch.empty((), dtype=dt).element_size()
            self._layout[name] = (off, nbytes, dt, shape)
            off = _align(off + nbytes)
        self.nbytes = off
        self.local = torch.zeros(self.nbytes, dtype=torch.uint8, device=device)          # this rank's results
        self.raw = self.local if self.world == 1 else torch.zeros(self.world * self.nbytes, dtype=torch.uint8, device=device)

    def views(self):
        """This rank's result tensors (views of `local`)."""
        return {name: self.local[o:o + b].view(dt).view(shape) for name, (o, b, dt, shape) in self._layout.items()}

    def all_views(self):
        """{name: (world, pairs, ...)} over the gathered buffer: one strided view per field, no copy."""
        out = {}
        rows = self.raw.view(self.world, self.nbytes)
        for name, (o, b, dt, shape) in self._layout.items():
            out[name] = rows[:, o:o + b].view(dt).view((self.world,) + tuple(shape))
        return out


# ----------------------------------------------------------------------------- NUMA placement
def bind_to_gpu_numa_node(device_index):
    """Restrict this process to the CPUs of the NUMA node the GPU hangs off, so that pinned host
    buffers allocated afterwards are node-local (first touch).  Eight ranks streaming from one
    node's memory is what limited the 8-GPU end-to-end number in round 1.  Returns a short
    description, or None when the topology cannot be read (nothing is changed then)."""
    try:
        pr = torch.cuda.get_device_properties(device_index)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node_path = "/sys/bus/pci/devices/%s/numa_node" % bdf
        node = int(open(node_path).read().strip())
        if node < 0:
            return None
        cpus = []
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.extend(range(int(lo), int(hi or lo) + 1))
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return "numa node %d (%d cpus)" % (node, len(allowed))
    except Exception:
        return None


# ----------------------------------------------------------------------------- engine
class RegistrationEngine:
    def __init__(self, K=750, radius=5.0, device=None, chunk_pairs=8, want_D=False, centered=True, subsample=None,
                 tau=0.05, seed=0, select=False, corr_sigma=1.5, corr_ds=0.6, tgt_ds=0.3, pc_corr_max_size=10000,
                 corr_num_nn=20):
        self.K = int(K)
        self.radius = float(radius)
        self.device = torch.device(device if device is not None else ("cuda:%d" % torch.cuda.current_device()))
        self.chunk_pairs = int(chunk_pairs)
        self.want_D = bool(want_D)
        self.centered = bool(centered)
        self.subsample = None if subsample is None else int(subsample)
        self.tau = float(tau)
        self.seed = int(seed)
        self.select = bool(select)
        self.corr = dict(corr_sigma=float(corr_sigma), corr_ds=float(corr_ds), tgt_ds=float(tgt_ds),
                         pc_corr_max_size=int(pc_corr_max_size), corr_num_nn=int(corr_num_nn))
        self._calls = 0
        self._streams = None
        self._staging = {}
        self._host_out = {}
        self._arenas = [{}, {}]          # outputs + workspace of register(): ping-pong, reused every other call
        self._packs = [None, None]       # ResultPack per arena
        self._graphs = {}                # CUDA graphs of register(): key -> (graph, out, arena)
        self._chunk_arena = [{}, {}]     # per-stream outputs of register_host() / register_stream()
        self._comm = None                # side stream of gather_step
        self._gather_done = [None, None]
        self._step = 0

    # ------------------------------------------------------------------ device-resident batch
    def _hyp_per_pair(self, n):
        return n if self.subsample is None else min(self.subsample, n)

    def _seed_pack(self, arena, slot, pairs, n, world=1):
        """Make the arena's T / argmin / dmin outputs views of the step's ResultPack."""
        m = self._hyp_per_pair(n)
        pack = self._packs[slot]
        if pack is None or pack.pairs != pairs or pack.m != m or pack.world != world:
            pack = ResultPack(pairs, m, self.device, world)
            self._packs[slot] = pack
        v = pack.views()
        arena["T"] = v["T"]
        if self.subsample is None:
            arena["argmin"], arena["dmin"] = v["argmin"], v["dmin"]
        return pack, v

    def _register(self, batch, arena, pack_views=None):
        self._calls += 1
        out = api.register_hypotheses(batch["src_pts"], batch["src_feat"], batch["src_kp"], batch["tgt_pts"],
                                      batch["tgt_feat"], batch["tgt_kp"], self.K, self.radius, want_D=self.want_D,
                                      centered=self.centered, buf=arena, subsample=self.subsample, tau=self.tau,
                                      subsample_seed=(self.seed << 20) + self._calls)
        if pack_views is not None and self.subsample is not None:
            pack_views["argmin"].copy_(out["match"][..., 1])
            pack_views["dmin"].copy_(out["dmin"])
        if self.select:
            T_best, best = self._select(batch, out["T"], arena, pack_views)
            out["T_best"], out["best"] = T_best, best
        return out

    def _select(self, batch, T, arena, pack_views):
        """evaluate.py:259-296 for every pair of the batch: the clouds the features live on stand in
        for the raw clouds (voxel de-duplication at corr_ds / tgt_ds, nearest-row feature transfer,
        random down-sampling to pc_corr_max_size, correlator pick)."""
        B = T.shape[0]
        T_best = pack_views["T_best"] if pack_views is not None else api._out(arena, "T_best", (B, 4, 4), torch.float32, T.device)
        best = pack_views["best"] if pack_views is not None else api._out(arena, "best", (B,), torch.int64, T.device)
        for b in range(B):
            Tb, ib, _ = api.select_hypothesis(batch["src_pts"][b], batch["tgt_pts"][b], batch["src_pts"][b:b + 1],
                                              batch["tgt_pts"][b:b + 1], batch["src_feat"][b:b + 1],
                                              batch["tgt_feat"][b:b + 1], T[b], **self.corr)
            T_best[b].copy_(Tb)
            best[b].copy_(ib)
        return T_best, best

    def register(self, batch, slot=0):
        """batch: dict with src_pts (B,N,3), src_feat (B,N,C), src_kp (B,n,3) and tgt_* on the
        device.  Returns dict(T (B,m,4,4), match (B,m,2) int64, dmin (B,m), F_src, F_tgt, D|None
        [, T_best (B,4,4), best (B,)]); m = n, or `subsample` when set.  The outputs live in the engine's
        arena `slot`: they are overwritten by the next call on that slot (clone what must survive) —
        no allocation happens in steady state."""
        B, n = batch["src_kp"].shape[0], batch["src_kp"].shape[1]
        arena = self._arenas[slot]
        _, v = self._seed_pack(arena, slot, B, n, self._packs[slot].world if self._packs[slot] is not None else 1)
        return self._register(batch, arena, v)

    def register_graphed(self, batch):
        """Same as `register`, replayed from a CUDA graph: the ~25 launches of a step (grid build,
        moments, descriptors, distance GEMM, solve, for both clouds) become one graph launch, which
        matters for small batches (a single pair is launch-latency bound).  The graph is keyed on
        the input tensors' addresses and shapes — refill the same buffers between calls; new
        buffers trigger a new capture.  Every graph owns its arena (outputs, intermediates and the
        kernels' workspace), so graphs of different shapes never share or free each other's memory."""
        if self.select:
            raise NotImplementedError("register_graphed: hypothesis selection reads sizes back to the host and cannot be captured")
        key = tuple((batch[k].data_ptr(), tuple(batch[k].shape)) for k in _IN_KEYS) + \
            (self.K, self.radius, self.want_D, self.centered, self.subsample, tuple(sorted(api.config.items(), key=str)))
        entry = self._graphs.get(key)
        if entry is None:
            arena = {}
            with torch.cuda.device(self.device):
                self._register(batch, arena)                 # warm-up: sizes the arena and its workspace
                torch.cuda.synchronize(self.device)
                calls = self._calls
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    out = self._register(batch, arena)
                self._calls = calls + 1
            entry = (graph, out, arena)
            self._graphs[key] = entry
        entry[0].replay()
        return entry[1]

    # ------------------------------------------------------------------ end-of-step collective
    def register_and_gather(self, batch, group=None):
        """One step of the multi-GPU job: `register` into the ping-pong arena of this step, then ONE
        all-gather of the step's packed results on a side stream.  The gather of step k overlaps the
        kernels of step k+1 (which writes the other arena).  Returns the ResultPack whose
        `all_views()` hold the whole job's results once `finish_gathers()` (or a synchronize) has run."""
        import torch.distributed as dist
        world = dist.get_world_size(group)
        slot = self._step % 2
        self._step += 1
        main = torch.cuda.current_stream(self.device)
        if self._gather_done[slot] is not None:
            main.wait_event(self._gather_done[slot])          # the gather that last read this arena's pack
        B, n = batch["src_kp"].shape[0], batch["src_kp"].shape[1]
        arena = self._arenas[slot]
        pack, v = self._seed_pack(arena, slot, B, n, world)
        self._register(batch, arena, v)
        if self._comm is None:
            self._comm = torch.cuda.Stream(self.device)
        ready = main.record_event()
        self._comm.wait_event(ready)
        with torch.cuda.stream(self._comm):
            if world > 1:
                dist.all_gather_into_tensor(pack.raw, pack.local, group=group)
            self._gather_done[slot] = self._comm.record_event()
        return pack

    def finish_gathers(self):
        main = torch.cuda.current_stream(self.device)
        for ev in self._gather_done:
            if ev is not None:
                main.wait_event(ev)

    # ------------------------------------------------------------------ host-resident batch
    def _stage(self, slot, key, shape, dtype):
        k = (slot, key)
        t = self._staging.get(k)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            t = torch.empty(shape, dtype=dtype, device=self.device)
            self._staging[k] = t
        return t

    def _stage_packed(self, slot, like):
        k = (slot, "packed")
        p = self._staging.get(k)
        if p is None or (p.pairs, p.N, p.n, p.C) != (like.pairs, like.N, like.n, like.C):
            p = PackedPairs(like.pairs, like.N, like.n, like.C, device=self.device)
            self._staging[k] = p
        return p

    def _pinned_out(self, key, shape, dtype):
        t = self._host_out.get(key)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            t = torch.empty(shape, dtype=dtype, pin_memory=True)
            self._host_out[key] = t
        return t

    def _host_outputs(self, B, n):
        m = self._hyp_per_pair(n)
        out = dict(T=self._pinned_out("T", (B, m, 4, 4), torch.float32),
                   match=self._pinned_out("match", (B, m, 2), torch.int64),
                   dmin=self._pinned_out("dmin", (B, m), torch.float32))
        if self.select:
            out["T_best"] = self._pinned_out("T_best", (B, 4, 4), torch.float32)
            out["best"] = self._pinned_out("best", (B,), torch.int64)
        return out

    def _run_chunk(self, dev, slot, host_out, lo, hi, full_chunk):
        out = self._register(dev, self._chunk_arena[slot] if full_chunk else {})
        host_out["T"][lo:hi].copy_(out["T"], non_blocking=True)
        host_out["match"][lo:hi].copy_(out["match"], non_blocking=True)
        host_out["dmin"][lo:hi].copy_(out["dmin"], non_blocking=True)
        if self.select:
            host_out["T_best"][lo:hi].copy_(out["T_best"], non_blocking=True)
            host_out["best"][lo:hi].copy_(out["best"], non_blocking=True)

    def register_host(self, batch):
        """batch: the same dict with PINNED CPU tensors, or a list of `PackedPairs` micro-batches
        (one H2D copy each instead of six).  Returns dict(T (B,m,4,4), match (B,m,2), dmin (B,m)
        [, T_best, best]) as pinned CPU tensors (reused between calls) that are valid once the caller's
        current stream has been synchronised.  Bytes moved: see `host_bytes(batch)`."""
        if isinstance(batch, (list, tuple)):
            return self.register_stream(batch)
        for k in _IN_KEYS:
            t = batch[k]
            if t.is_cuda or not t.is_pinned():
                raise ValueError("register_host: %s must be a pinned CPU tensor" % k)
        B, n = batch["src_kp"].shape[0], batch["src_kp"].shape[1]
        with torch.cuda.device(self.device):
            if self._streams is None:
                self._streams = [torch.cuda.Stream(self.device), torch.cuda.Stream(self.device)]
            main = torch.cuda.current_stream(self.device)
            start = main.record_event()
            host_out = self._host_outputs(B, n)
            cp = max(1, min(self.chunk_pairs, B))
            for ci, lo in enumerate(range(0, B, cp)):
                hi = min(lo + cp, B)
                slot = ci % 2
                s = self._streams[slot]
                if ci < 2:
                    s.wait_event(start)
                with torch.cuda.stream(s):
                    dev = {}
                    for k in _IN_KEYS:
                        src = batch[k][lo:hi]
                        buf = self._stage(slot, k, (cp,) + tuple(src.shape[1:]), src.dtype)[: hi - lo]
                        buf.copy_(src, non_blocking=True)
                        dev[k] = buf
                    self._run_chunk(dev, slot, host_out, lo, hi, hi - lo == cp)
            for s in self._streams:
                main.wait_event(s.record_event())
        return host_out

    def register_stream(self, micro_batches, total_pairs=None):
        """Streams host micro-batches through the device (BASELINE config #5): `micro_batches` is an
        iterable of `PackedPairs` (pinned; all of the same geometry).  Micro-batch i+1 is copied
        host->device on one stream while micro-batch i is being registered on the other (two device
        staging slots, two streams); each micro-batch is ONE copy.  Results accumulate in pinned host
        buffers sized for `total_pairs` (default: the sum over a list).  Returns the same dict as
        `register_host`."""
        if total_pairs is None:
            micro_batches = list(micro_batches)
            total_pairs = sum(p.pairs for p in micro_batches)
        host_out = None
        with torch.cuda.device(self.device):
            if self._streams is None:
                self._streams = [torch.cuda.Stream(self.device), torch.cuda.Stream(self.device)]
            main = torch.cuda.current_stream(self.device)
            start = main.record_event()
            lo = 0
            for ci, mb in enumerate(micro_batches):
                if not isinstance(mb, PackedPairs) or mb.raw.is_cuda or (torch.cuda.is_available() and not mb.raw.is_pinned()):
                    raise ValueError("register_stream: micro-batches must be pinned PackedPairs")
                if host_out is None:
                    host_out = self._host_outputs(total_pairs, mb.n)
                hi = lo + mb.pairs
                if hi > total_pairs:
                    raise ValueError("register_stream: more pairs than total_pairs")
                slot = ci % 2
                s = self._streams[slot]
                if ci < 2:
                    s.wait_event(start)
                with torch.cuda.stream(s):
                    dev = self._stage_packed(slot, mb)
                    dev.raw.copy_(mb.raw, non_blocking=True)
                    self._run_chunk(dev.views, slot, host_out, lo, hi, True)
                lo = hi
            for s in self._streams:
                main.wait_event(s.record_event())
        if host_out is None:
            raise ValueError("register_stream: no micro-batches")
        return {k: v[:lo] for k, v in host_out.items()}

    def host_bytes(self, batch):
        """(h2d_bytes, d2h_bytes) one register_host / register_stream call moves."""
        if isinstance(batch, (list, tuple)):
            h2d = sum(p.nbytes for p in batch)
            B, n = sum(p.pairs for p in batch), batch[0].n
        else:
            h2d = sum(batch[k].numel() * batch[k].element_size() for k in _IN_KEYS)
            B, n = batch["src_kp"].shape[0], batch["src_kp"].shape[1]
        m = self._hyp_per_pair(n)
        d2h = B * m * (16 * 4 + 16 + 4) + (B * (64 + 8) if self.select else 0)
        return h2d, d2h


def shard_range(n_pairs, rank, world):
    """Contiguous block of pairs owned by `rank`: ceil(n_pairs / world) each."""
    per = (n_pairs + world - 1) // world
    lo = min(rank * per, n_pairs)
    return lo, min(lo + per, n_pairs)


def gather_results(result, group=None):
    """All-gather of a dict of per-pair results (one collective per entry) so that every rank — rank 0
    in particular — holds the whole job's output.  Equal shard sizes are required (pad the last
    shard).  Works with NCCL on GPUs and with gloo on CPU tensors (tests).  The engine's own steps use
    the packed single-collective path (`RegistrationEngine.register_and_gather`)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    out = {}
    for k in ("T", "match", "dmin", "T_best", "best"):
        if k not in result or result[k] is None:
            continue
        t = result[k].contiguous()
        full = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(full, t, group=group)
        out[k] = full
    return out


def gather_packed(pack, group=None):
    """ONE collective for a step's results: all-gather of every rank's ResultPack.local into
    ResultPack.raw.  Returns `pack.all_views()`."""
    import torch.distributed as dist
    if pack.world > 1:
        dist.all_gather_into_tensor(pack.raw, pack.local, group=group)
    return pack.all_views()
