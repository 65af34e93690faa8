"""Batched registration engine: the call a user makes for many pairs at once.

`register(batch)` runs evaluate.py:206-257 (UME generation for both clouds, subspace distances with
fused arg-min, one rigid hypothesis per match) for a batch of pairs that is already resident on
the device.  `register_host(batch)` takes PINNED HOST buffers: the batch is cut into chunks that
alternate between two CUDA streams, so the host->device copy of one chunk overlaps the kernels
of the previous one, and results come back into pinned host buffers.

Multi-GPU: pairs are independent (no cross-pair state anywhere in evaluate.py:175-299), so ranks
take contiguous blocks of pairs and `gather_results` does the single end-of-step NCCL all-gather
of the per-pair results.
"""
import torch

from . import api

_IN_KEYS = ("src_pts", "src_feat", "src_kp", "tgt_pts", "tgt_feat", "tgt_kp")


class RegistrationEngine:
    def __init__(self, K=750, radius=5.0, device=None, chunk_pairs=8, want_D=False, centered=True):
        self.K = int(K)
        self.radius = float(radius)
        self.device = torch.device(device if device is not None else ("cuda:%d" % torch.cuda.current_device()))
        self.chunk_pairs = int(chunk_pairs)
        self.want_D = bool(want_D)
        self.centered = bool(centered)
        self._streams = None
        self._staging = {}
        self._host_out = {}
        self._arena = {}                 # outputs of register(): reused every call
        self._graphs = {}                # CUDA graphs of register(), keyed on input addresses / shapes
        self._chunk_arena = [{}, {}]     # per-stream outputs of register_host()

    # ------------------------------------------------------------------ device-resident batch
    def register(self, batch):
        """batch: dict with src_pts (B,N,3), src_feat (B,N,C), src_kp (B,n,3) and tgt_* on the
        device.  Returns dict(T (B,n,4,4), match (B,n,2) int64, dmin (B,n), F_src, F_tgt, D|None).
        The outputs live in the engine's arena: they are overwritten by the next call (clone what
        must survive) — no allocation happens in steady state."""
        return api.register_hypotheses(batch["src_pts"], batch["src_feat"], batch["src_kp"], batch["tgt_pts"],
                                       batch["tgt_feat"], batch["tgt_kp"], self.K, self.radius, want_D=self.want_D,
                                       centered=self.centered, buf=self._arena)

    def register_graphed(self, batch):
        """Same as `register`, replayed from a CUDA graph: the ~25 launches of a step (grid build,
        moments, descriptors, distance GEMM, solve, for both clouds) become one graph launch, which
        matters for small batches (a single pair is launch-latency bound).  The graph is keyed on
        the input tensors' addresses and shapes — refill the same buffers between calls; new
        buffers trigger a new capture."""
        key = tuple((batch[k].data_ptr(), tuple(batch[k].shape)) for k in _IN_KEYS) + \
            (self.K, self.radius, self.want_D, self.centered, tuple(sorted(api.config.items(), key=str)))
        entry = self._graphs.get(key)
        if entry is None:
            with torch.cuda.device(self.device):
                self.register(batch)                      # warm-up: arena, function attributes
                torch.cuda.synchronize(self.device)
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    out = self.register(batch)
            entry = (graph, out)
            self._graphs[key] = entry
        entry[0].replay()
        return entry[1]

    # ------------------------------------------------------------------ host-resident batch
    def _stage(self, slot, key, shape, dtype):
        k = (slot, key)
        t = self._staging.get(k)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            t = torch.empty(shape, dtype=dtype, device=self.device)
            self._staging[k] = t
        return t

    def _pinned_out(self, key, shape, dtype):
        t = self._host_out.get(key)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            t = torch.empty(shape, dtype=dtype, pin_memory=True)
            self._host_out[key] = t
        return t

    def register_host(self, batch):
        """batch: the same dict with PINNED CPU tensors.  Returns dict(T, match, dmin) as pinned
        CPU tensors (reused between calls) that are valid once the caller's current stream has been
        synchronised.  Bytes moved: see `host_bytes(batch)`."""
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
            T = self._pinned_out("T", (B, n, 4, 4), torch.float32)
            match = self._pinned_out("match", (B, n), torch.int64)
            dmin = self._pinned_out("dmin", (B, n), torch.float32)
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
                    out = api.register_hypotheses(dev["src_pts"], dev["src_feat"], dev["src_kp"], dev["tgt_pts"],
                                                  dev["tgt_feat"], dev["tgt_kp"], self.K, self.radius, want_D=False,
                                                  centered=self.centered,
                                                  buf=self._chunk_arena[slot] if hi - lo == cp else None)
                    T[lo:hi].copy_(out["T"], non_blocking=True)
                    match[lo:hi].copy_(out["match"][..., 1], non_blocking=True)
                    dmin[lo:hi].copy_(out["dmin"], non_blocking=True)
            for s in self._streams:
                main.wait_event(s.record_event())
        return dict(T=T, match=match, dmin=dmin)

    @staticmethod
    def host_bytes(batch):
        """(h2d_bytes, d2h_bytes) one register_host call moves."""
        h2d = sum(batch[k].numel() * batch[k].element_size() for k in _IN_KEYS)
        B, n = batch["src_kp"].shape[0], batch["src_kp"].shape[1]
        d2h = B * n * (16 * 4 + 8 + 4)
        return h2d, d2h


def shard_range(n_pairs, rank, world):
    """Contiguous block of pairs owned by `rank`: ceil(n_pairs / world) each."""
    per = (n_pairs + world - 1) // world
    lo = min(rank * per, n_pairs)
    return lo, min(lo + per, n_pairs)


def gather_results(result, group=None):
    """The single end-of-step collective: all-gather of the per-pair results (T hypotheses,
    arg-min match, match distance) so that every rank — rank 0 in particular — holds the whole
    job's output.  Equal shard sizes are required (pad the last shard).  Works with NCCL on GPUs
    and with gloo on CPU tensors (tests)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    out = {}
    for k in ("T", "match", "dmin"):
        t = result[k].contiguous()
        full = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(full, t, group=group)
        out[k] = full
    return out
