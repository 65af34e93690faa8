"""Host-side mirror of the reference's call signatures for the UME hot path.

Every function here has the name, argument meaning and return layout of the reference function it
replaces (file:line cited per function) and does its work in libumereg_b200.so (hand-written
sm_100a CUDA behind the C ABI of include/umereg_b200.h).  torch is used for device memory,
streams and nothing else.  Inputs must be CUDA tensors: there is no CPU fallback.
"""
import collections
import ctypes
import threading

import torch

from . import _lib

_BallQuery = collections.namedtuple("_BallQuery", ["dists", "idx", "knn"])
_KNN = collections.namedtuple("_KNN", ["dists", "idx", "knn"])

# Global defaults (can be changed by callers / tests):
#   fma_dist : evaluate the squared distance with fused multiply-adds (what nvcc makes of
#              pytorch3d's CUDA kernel) instead of separately rounded mul/add (pytorch3d CPU build)
#   cell_div2: search grid with cell = radius / 2
#   cta_moments / warp_moments: force the CTA-per-keypoint / warp-per-keypoint gather+moment kernel
#              (default: warp kernel for C in {16,32,64,128} and launches of >= 3072 keypoints)
#   cdist_impl: 0 = fp32 SIMT distance kernel, 1 = tcgen05 tensor-core kernel
#               (C = 32 / 64 only), None = tensor cores whenever the channel count allows
config = {"fma_dist": False, "cell_div2": False, "cdist_impl": None, "cta_moments": False, "warp_moments": False}

_workspaces = {}
_workspaces_lock = threading.Lock()


def _flags():
    return ((_lib.UME_FLAG_FMA_DIST if config["fma_dist"] else 0) | (_lib.UME_FLAG_CELL_DIV2 if config["cell_div2"] else 0)
            | (_lib.UME_FLAG_CTA_MOMENTS if config["cta_moments"] is True else 0)
            | (_lib.UME_FLAG_WARP_MOMENTS if config["cta_moments"] is False and config.get("warp_moments") else 0))


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _workspace(nbytes, device, buf=None):
    """Scratch buffer for one call, grown on demand.  With an arena dict `buf` the workspace belongs
    to that arena (an engine, a captured graph): nothing else ever touches it.  Without, it is a
    per (device, stream) buffer — calls are stream-ordered, so reuse by consecutive calls on the same
    stream is safe; the table is guarded by a lock and a stream's buffer is only ever replaced by
    a larger one."""
    if buf is not None:
        ws = buf.get("_ws")
        if ws is None or ws.numel() < nbytes or ws.device != device:
            ws = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
            buf["_ws"] = ws
        return ws
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    with _workspaces_lock:
        ws = _workspaces.get(key)
        if ws is None or ws.numel() < nbytes:
            ws = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
            _workspaces[key] = ws
        return ws


def release_workspaces():
    """Drop every per-stream scratch buffer (they are otherwise kept for the life of the process)."""
    with _workspaces_lock:
        _workspaces.clear()


def _out(buf, key, shape, dtype, device):
    """Output tensor: fresh, or — when the caller passes an arena dict `buf` — the arena's tensor
    of that name, reused across calls (no allocator traffic in steady state, stable pointers)."""
    if buf is None:
        return torch.empty(shape, dtype=dtype, device=device)
    t = buf.get(key)
    if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype or t.device != device:
        t = torch.empty(shape, dtype=dtype, device=device)
        buf[key] = t
    return t


def _dev_f32(t, name, ndim=None):
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor" % name)
    if not t.is_cuda:
        raise RuntimeError("%s is on %s: umeregrobust_b200 only runs on CUDA devices (no CPU fallback)" % (name, t.device))
    if ndim is not None and t.dim() != ndim:
        raise ValueError("%s must have %d dims, got shape %s" % (name, ndim, tuple(t.shape)))
    if t.requires_grad and torch.is_grad_enabled():
        raise RuntimeError("%s requires grad: the inference kernels of umeregrobust_b200 are not differentiable; use the "
                           "autograd mirrors in umeregrobust_b200.training (or wrap the call in torch.no_grad())" % name)
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


# ----------------------------------------------------------------------------- pytorch3d names
def ball_query(p1, p2, lengths1=None, lengths2=None, K=500, radius=0.2, return_nn=True):
    """pytorch3d.ops.ball_query as called at evaluate.py:51 and utils/loc_utils.py:383-384:
    first K rows of p2 (row order) with dist^2 < radius^2; idx -1 padded (int64), dists 0 padded,
    knn zero padded.  Returns a namedtuple (dists, idx, knn).
    lengths1 / lengths2 (B,) (utils/loc_utils.py:113 passes lengths1): rows of p1 past lengths1[b] get
    the padding values, rows of p2 past lengths2[b] are never neighbours (they are handed to the
    kernel as NaN points, which the search grid drops)."""
    p1 = _dev_f32(p1, "p1", 3)
    p2 = _dev_f32(p2, "p2", 3)
    if lengths2 is not None:
        beyond = torch.arange(p2.shape[1], device=p2.device)[None] >= torch.as_tensor(lengths2, device=p2.device)[:, None]
        p2 = torch.where(beyond[..., None], torch.full_like(p2, float("nan")), p2)
    if p1.shape[0] != p2.shape[0] or p1.shape[2] != 3 or p2.shape[2] != 3:
        raise ValueError("ball_query: p1 %s and p2 %s must be (B,P,3) with equal B" % (tuple(p1.shape), tuple(p2.shape)))
    B, P1, _ = p1.shape
    P2 = p2.shape[1]
    K = int(K)
    idx = torch.empty((B, P1, K), dtype=torch.int64, device=p1.device)
    dists = torch.empty((B, P1, K), dtype=torch.float32, device=p1.device)
    nn = torch.empty((B, P1, K, 3), dtype=torch.float32, device=p1.device) if return_nn else None
    with torch.cuda.device(p1.device):
        L = _lib.lib()
        nbytes = L.ume_ball_query_workspace_bytes(B, P1, P2, K)
        ws = _workspace(nbytes, p1.device)
        rc = L.ume_ball_query_f32(_ptr(p1), _ptr(p2), B, P1, P2, K, float(radius), _flags(), _ptr(idx), _ptr(dists),
                                  _ptr(nn), None, _ptr(ws), ws.numel(), _stream())
    _lib.check(rc, "ball_query")
    if lengths1 is not None:
        beyond = torch.arange(P1, device=p1.device)[None] >= torch.as_tensor(lengths1, device=p1.device)[:, None]
        idx[beyond] = -1
        dists[beyond] = 0
        if nn is not None:
            nn[beyond] = 0
    return _BallQuery(dists, idx, nn)


def knn_points(p1, p2, lengths1=None, lengths2=None, norm=2, K=1, version=-1, return_nn=False, return_sorted=True):
    """pytorch3d.ops.knn_points for the K = 1 case the hot path uses (evaluate.py:272,274):
    (dists (B,P1,1) squared L2, idx (B,P1,1) int64, knn or None)."""
    if lengths1 is not None or lengths2 is not None:
        raise NotImplementedError("knn_points: heterogeneous lengths are not supported")
    if norm != 2:
        raise NotImplementedError("knn_points: only norm=2")
    p1 = _dev_f32(p1, "p1", 3)
    p2 = _dev_f32(p2, "p2", 3)
    if p1.shape[0] != p2.shape[0]:
        raise ValueError("knn_points: batch sizes differ")
    B, P1, _ = p1.shape
    P2 = p2.shape[1]
    K = int(K)
    if K != 1:
        # general K (utils/loc_utils.py:580,623): sorted ascending, lower row index on ties
        if K < 1 or K > 64:
            raise NotImplementedError("knn_points: K must be in [1, 64]")
        idx = torch.empty((B, P1, K), dtype=torch.int64, device=p1.device)
        d2 = torch.empty((B, P1, K), dtype=torch.float32, device=p1.device)
        with torch.cuda.device(p1.device):
            L = _lib.lib()
            ws = _workspace(L.ume_knn_workspace_bytes(B, P1, P2), p1.device)
            rc = L.ume_knn_f32(_ptr(p1), _ptr(p2), B, P1, P2, K, _flags(), _ptr(idx), _ptr(d2), _ptr(ws), ws.numel(),
                               _stream())
        _lib.check(rc, "knn_points")
        return _KNN(d2, idx, knn_gather(p2, idx) if return_nn else None)
    idx = torch.empty((B, P1, 1), dtype=torch.int64, device=p1.device)
    d2 = torch.empty((B, P1, 1), dtype=torch.float32, device=p1.device)
    with torch.cuda.device(p1.device):
        L = _lib.lib()
        ws = _workspace(L.ume_knn1_workspace_bytes(B, P1, P2), p1.device)
        rc = L.ume_knn1_gather_f32(_ptr(p1), _ptr(p2), None, B, P1, P2, 0, _flags(), _ptr(idx), _ptr(d2), None,
                                   _ptr(ws), ws.numel(), _stream())
    _lib.check(rc, "knn_points")
    knn = knn_gather(p2, idx) if return_nn else None
    return _KNN(d2, idx, knn)


def knn_gather(x, idx, lengths=None):
    """pytorch3d.ops.knn_gather (evaluate.py:273,275; utils/loc_utils.py:354,581): x (B,M,U),
    idx (B,L,K) -> (B,L,K,U).  Pure indexing (pytorch3d implements it as expand+gather too)."""
    B, M, U = x.shape
    _, Lq, K = idx.shape
    flat = idx.reshape(B, Lq * K, 1).expand(-1, -1, U)
    return torch.gather(x, 1, flat).reshape(B, Lq, K, U)


def knn1_transfer(q, p, x):
    """Fused evaluate.py:272-273 (and :274-275): features of the nearest row of `p` for every row
    of `q`: knn_gather(x, knn_points(q, p, K=1).idx)[:, :, 0, :] without the index round trip."""
    q = _dev_f32(q, "q", 3)
    p = _dev_f32(p, "p", 3)
    x = _dev_f32(x, "x", 3)
    B, P1, _ = q.shape
    P2, U = x.shape[1], x.shape[2]
    if p.shape[1] != P2 or p.shape[0] != B or x.shape[0] != B:
        raise ValueError("knn1_transfer: shapes do not agree")
    out = torch.empty((B, P1, U), dtype=torch.float32, device=q.device)
    with torch.cuda.device(q.device):
        L = _lib.lib()
        ws = _workspace(L.ume_knn1_workspace_bytes(B, P1, P2), q.device)
        rc = L.ume_knn1_gather_f32(_ptr(q), _ptr(p), _ptr(x), B, P1, P2, U, _flags(), None, None, _ptr(out),
                                   _ptr(ws), ws.numel(), _stream())
    _lib.check(rc, "knn1_transfer")
    return out


# ----------------------------------------------------------------------------- UME moments
def ume_moments(pts, kpts, feat, K, radius, return_centered=False, return_count=False, buf=None, tag="", raw=False):
    """Fused ball-query + gather + moment build.  F (B,n,C,4) exactly as evaluate.py:50-60
    produces it; optionally the keypoint-centred matrix Fc (same column space) and the
    neighbour count per keypoint.  raw=True leaves out the normalisation of evaluate.py:59
    (utils/loc_utils.py:157-161 with normalized_ume=False; C in {16,32,64,128})."""
    pts = _dev_f32(pts, "pts", 3)
    kpts = _dev_f32(kpts, "kpts", 3)
    feat = _dev_f32(feat, "feat", 3)
    B, N, _ = pts.shape
    if kpts.shape[0] != B or feat.shape[0] != B or feat.shape[1] != N or pts.shape[2] != 3 or kpts.shape[2] != 3:
        raise ValueError("ume_moments: pts %s, kpts %s, feat %s do not agree" %
                         (tuple(pts.shape), tuple(kpts.shape), tuple(feat.shape)))
    n, C = kpts.shape[1], feat.shape[2]
    dev = pts.device
    F = _out(buf, "F" + tag, (B, n, C, 4), torch.float32, dev)
    Fc = _out(buf, "Fc" + tag, (B, n, C, 4), torch.float32, dev) if return_centered else None
    cnt = _out(buf, "cnt" + tag, (B, n), torch.int32, dev) if return_count else None
    with torch.cuda.device(dev):
        L = _lib.lib()
        ws = _workspace(L.ume_moments_workspace_bytes(B, N, n, C, int(K)), dev, buf)
        rc = L.ume_moments_f32(_ptr(pts), _ptr(kpts), _ptr(feat), B, N, n, C, int(K), float(radius),
                               _flags() | (_lib.UME_FLAG_RAW_MOMENTS if raw else 0),
                               _ptr(F), _ptr(Fc), _ptr(cnt), _ptr(ws), ws.numel(), _stream())
    _lib.check(rc, "ume_moments")
    out = (F,)
    if return_centered:
        out += (Fc,)
    if return_count:
        out += (cnt,)
    return out[0] if len(out) == 1 else out


def ume_moments_pair(pts1, kpts1, feat1, pts2, kpts2, feat2, K, radius, return_centered=False, buf=None):
    """`ume_moments` of a source and a target batch in ONE search-grid build and ONE kernel launch
    (`ume_moments_pair_f32`): returns (F1, F2, F_both) or (F1, Fc1, F2, Fc2, Fc_both) — the per-side results are
    the halves of one (2B,n,C,4) allocation (the last element), bit-identical to two `ume_moments` calls.  Returns None when the pair entry does not apply (different
    shapes on the two sides, a channel count or a launch size the warp-per-keypoint kernel is not used for)."""
    if tuple(pts1.shape) != tuple(pts2.shape) or tuple(kpts1.shape) != tuple(kpts2.shape) \
            or tuple(feat1.shape) != tuple(feat2.shape) or feat1.dim() != 3:
        return None
    B, N, _ = pts1.shape
    n, C = kpts1.shape[1], feat1.shape[2]
    if C not in (16, 32, 64, 128) or 2 * B > 65535:
        return None
    pts1, kpts1, feat1 = _dev_f32(pts1, "pts1", 3), _dev_f32(kpts1, "kpts1", 3), _dev_f32(feat1, "feat1", 3)
    pts2, kpts2, feat2 = _dev_f32(pts2, "pts2", 3), _dev_f32(kpts2, "kpts2", 3), _dev_f32(feat2, "feat2", 3)
    dev = pts1.device
    F = _out(buf, "F_pair", (2 * B, n, C, 4), torch.float32, dev)
    Fc = _out(buf, "Fc_pair", (2 * B, n, C, 4), torch.float32, dev) if return_centered else None
    with torch.cuda.device(dev):
        L = _lib.lib()
        ws = _workspace(L.ume_moments_workspace_bytes(2 * B, N, n, C, int(K)), dev, buf)
        rc = L.ume_moments_pair_f32(_ptr(pts1), _ptr(kpts1), _ptr(feat1), _ptr(pts2), _ptr(kpts2), _ptr(feat2), B, N, n, C,
                                    int(K), float(radius), _flags(), _ptr(F), _ptr(Fc), None, _ptr(ws), ws.numel(), _stream())
    _lib.check(rc, "ume_moments_pair")
    if return_centered:
        return F[:B], Fc[:B], F[B:], Fc[B:], Fc
    return F[:B], F[B:], F


def ume_moments_backward(pts, kpts, grad_F, K, radius):
    """Gradient of the RAW moments with respect to the features: the same
    neighbourhoods as `ume_moments(pts, kpts, ., K, radius)`, every neighbour row j of keypoint i
    receives grad_F[i,c,0] + grad_F[i,c,1:4] . pts[j].  grad_F (B,n,C,4) -> (B,N,C)."""
    pts = _dev_f32(pts, "pts", 3)
    kpts = _dev_f32(kpts, "kpts", 3)
    g = _dev_f32(grad_F, "grad_F", 4)
    B, N, _ = pts.shape
    n, C = kpts.shape[1], g.shape[2]
    if tuple(g.shape) != (B, n, C, 4):
        raise ValueError("ume_moments_backward: grad_F %s does not match (B,n,C,4)" % (tuple(g.shape),))
    out = torch.zeros((B, N, C), dtype=torch.float32, device=pts.device)
    with torch.cuda.device(pts.device):
        L = _lib.lib()
        ws = _workspace(L.ume_moments_workspace_bytes(B, N, n, C, int(K)), pts.device)
        rc = L.ume_moments_backward_f32(_ptr(pts), _ptr(kpts), _ptr(g), B, N, n, C, int(K), float(radius), _flags(),
                                        _ptr(out), _ptr(ws), ws.numel(), _stream())
    _lib.check(rc, "ume_moments_backward")
    return out


def neighbor_count(pts, kpts, K, radius):
    """min(K, number of rows of pts within `radius` of each keypoint) -> (B,n) int32, without the
    (B,n,K) index tensor ((ball_query(...).idx > -1).sum(-1) as used at utils/loc_utils.py:103,119)."""
    pts = _dev_f32(pts, "pts", 3)
    kpts = _dev_f32(kpts, "kpts", 3)
    B, N, _ = pts.shape
    n = kpts.shape[1]
    cnt = torch.zeros((B, n), dtype=torch.int32, device=pts.device)
    with torch.cuda.device(pts.device):
        L = _lib.lib()
        ws = _workspace(L.ume_moments_workspace_bytes(B, N, n, 32, int(K)), pts.device)
        rc = L.ume_neighbor_count_f32(_ptr(pts), _ptr(kpts), B, N, n, int(K), float(radius), _flags(), _ptr(cnt),
                                      _ptr(ws), ws.numel(), _stream())
    _lib.check(rc, "neighbor_count")
    return cnt


def my_ume_generation(pts, kpts, feat, args):
    """evaluate.py:50-60 `my_ume_generation(pts, kpts, feat, args)` -> F (B,n,C,4); reads
    args.ume_max_nn and args.ume_r_nn.  (The reference hard-codes C = 32 at :55; any C works here.)"""
    return ume_moments(pts, kpts, feat, args.ume_max_nn, args.ume_r_nn)


def create_local_ume_matrix(nn_pts, nn_feat):
    """utils/loc_utils.py:434-445: un-normalised moments of pre-gathered neighbourhoods
    (bs, n_kp, pc, 3) / (bs, n_kp, pc, C).  Kept as plain batched torch matmuls: it is an unused
    helper in the reference (no caller in any script) and not part of the measured path."""
    F1 = nn_feat.transpose(-1, -2) @ nn_pts
    F0 = nn_feat.transpose(-1, -2).sum(dim=-1, keepdim=True)
    return torch.cat([F0, F1], dim=-1)


# ----------------------------------------------------------------------------- descriptors / distances
def ume_descriptors(ume, return_rank=False, buf=None, tag=""):
    """Orthonormal basis rows Qt (..., 4, C) of the column space of each (..., C, 4) UME matrix
    (the QR of utils/loc_utils.py:9,11 up to the choice of basis)."""
    ume = _dev_f32(ume, "ume")
    if ume.dim() < 2 or ume.shape[-1] != 4:
        raise ValueError("ume_descriptors: expected (..., C, 4), got %s" % (tuple(ume.shape),))
    C = ume.shape[-2]
    nmat = ume.numel() // (C * 4) if C > 0 else 0
    Qt = _out(buf, "Qt" + tag, tuple(ume.shape[:-2]) + (4, C), torch.float32, ume.device)
    rank = _out(buf, "rank" + tag, tuple(ume.shape[:-2]), torch.int32, ume.device) if return_rank else None
    with torch.cuda.device(ume.device):
        rc = _lib.lib().ume_orthonormalize_f32(_ptr(ume), nmat, C, _ptr(Qt), _ptr(rank), _stream())
    _lib.check(rc, "ume_descriptors")
    return (Qt, rank) if return_rank else Qt


def ume_descriptors_split(ume, buf=None, tag="", want_Qt=False):
    """Orthonormal bases as the tensor-core distance kernel's operand: Qh (..., 4, 2C) float16, rows
    [hi | lo] of 256 q (`ume_orthonormalize_split_f32`), optionally with the fp32 Qt as well.  Feeding
    `descriptor_cdist_split` with these skips the split pre-pass of the generic `descriptor_cdist`."""
    ume = _dev_f32(ume, "ume")
    if ume.dim() < 2 or ume.shape[-1] != 4:
        raise ValueError("ume_descriptors_split: expected (..., C, 4), got %s" % (tuple(ume.shape),))
    C = ume.shape[-2]
    nmat = ume.numel() // (C * 4) if C > 0 else 0
    Qh = _out(buf, "Qh" + tag, tuple(ume.shape[:-2]) + (4, 2 * C), torch.float16, ume.device)
    Qt = _out(buf, "Qt" + tag, tuple(ume.shape[:-2]) + (4, C), torch.float32, ume.device) if want_Qt else None
    with torch.cuda.device(ume.device):
        rc = _lib.lib().ume_orthonormalize_split_f32(_ptr(ume), nmat, C, _ptr(Qt), _ptr(Qh), None, _stream())
    _lib.check(rc, "ume_descriptors_split")
    return (Qh, Qt) if want_Qt else Qh


def descriptor_cdist_split(Qh1, Qh2, want_D=True, want_argmin=False, buf=None):
    """`descriptor_cdist` on pre-split float16 operands (B,n,4,2C) from `ume_descriptors_split`: the
    tcgen05 kernel without its split pre-pass and without a workspace.  C = 32 or 64."""
    if Qh1.dtype != torch.float16 or Qh2.dtype != torch.float16 or not Qh1.is_cuda or Qh1.dim() != 4 or Qh2.dim() != 4:
        raise ValueError("descriptor_cdist_split: expected float16 CUDA tensors (B,n,4,2C)")
    Qh1, Qh2 = Qh1.contiguous(), Qh2.contiguous()
    B, n1, _, C2 = Qh1.shape
    n2 = Qh2.shape[1]
    if Qh2.shape[0] != B or Qh2.shape[3] != C2 or Qh1.shape[2] != 4 or Qh2.shape[2] != 4:
        raise ValueError("descriptor_cdist_split: shapes %s and %s do not agree" % (tuple(Qh1.shape), tuple(Qh2.shape)))
    dev = Qh1.device
    D = _out(buf, "D", (B, n1, n2), torch.float32, dev) if want_D else None
    am = _out(buf, "argmin", (B, n1), torch.int64, dev) if want_argmin else None
    dm = _out(buf, "dmin", (B, n1), torch.float32, dev) if want_argmin else None
    with torch.cuda.device(dev):
        rc = _lib.lib().ume_cdist_split_f16(_ptr(Qh1), _ptr(Qh2), B, n1, n2, C2 // 2, _ptr(D), _ptr(am), _ptr(dm), _stream())
    _lib.check(rc, "descriptor_cdist_split")
    return D, am, dm


def descriptor_cdist(Qt1, Qt2, want_D=True, want_argmin=False, impl=None, buf=None):
    """All-pairs D = sqrt(4 - |Q1^T Q2|_F^2) between descriptor sets (B,n1,4,C) x (B,n2,4,C), with
    the row arg-min (first index on ties) fused.  Returns (D or None, argmin or None, dmin or None)."""
    Qt1 = _dev_f32(Qt1, "Qt1", 4)
    Qt2 = _dev_f32(Qt2, "Qt2", 4)
    B, n1, _, C = Qt1.shape
    n2 = Qt2.shape[1]
    if Qt2.shape[0] != B or Qt2.shape[3] != C or Qt1.shape[2] != 4 or Qt2.shape[2] != 4:
        raise ValueError("descriptor_cdist: shapes %s and %s do not agree" % (tuple(Qt1.shape), tuple(Qt2.shape)))
    impl = config["cdist_impl"] if impl is None else impl
    if impl is None:
        impl = 1 if C in (32, 64) else 0
    dev = Qt1.device
    D = _out(buf, "D", (B, n1, n2), torch.float32, dev) if want_D else None
    am = _out(buf, "argmin", (B, n1), torch.int64, dev) if want_argmin else None
    dm = _out(buf, "dmin", (B, n1), torch.float32, dev) if want_argmin else None
    with torch.cuda.device(dev):
        L = _lib.lib()
        nbytes = L.ume_cdist_workspace_bytes(B, n1, n2, C, impl)
        ws = _workspace(nbytes, dev, buf) if nbytes else None
        rc = L.ume_cdist_f32(_ptr(Qt1), _ptr(Qt2), B, n1, n2, C, impl, _ptr(D), _ptr(am), _ptr(dm), _ptr(ws),
                             ws.numel() if ws is not None else 0, _stream())
    _lib.check(rc, "descriptor_cdist")
    return D, am, dm


def ume_cdist(ume1, ume2):
    """utils/loc_utils.py:8-15 `ume_cdist(ume1, ume2)`: (bs,n1,C,4) x (bs,n2,C,4) -> D (bs,n1,n2),
    D_ij = |P1_i - P2_j|_F / sqrt(2) with P = Q Q^T."""
    if ume1.dim() != 4 or ume2.dim() != 4:
        raise ValueError("ume_cdist: expected (bs, n, C, 4) tensors")
    D, _, _ = descriptor_cdist(ume_descriptors(ume1), ume_descriptors(ume2), want_D=True, want_argmin=False)
    return D


# ----------------------------------------------------------------------------- rigid solve
def rigid_solve(G, H, gi=None, hi=None, offG=None, offH=None, buf=None):
    """Batched closed-form rigid hypotheses.  G (B,nG,C,4), H (B,nH,C,4); hypothesis (b,i) pairs
    G[b,gi[b,i]] with H[b,hi[b,i]] (identity when the index is None).  offG/offH: the points the
    moments are relative to (for the centred `Fc` matrices).  Returns T (B,nm,4,4)."""
    G = _dev_f32(G, "G", 4)
    H = _dev_f32(H, "H", 4)
    B, nG, C, _ = G.shape
    nH = H.shape[1]
    if H.shape[0] != B or H.shape[2] != C or G.shape[3] != 4 or H.shape[3] != 4:
        raise ValueError("rigid_solve: shapes %s and %s do not agree" % (tuple(G.shape), tuple(H.shape)))
    if gi is not None:
        gi = gi.to(torch.int64).contiguous()
    if hi is not None:
        hi = hi.to(torch.int64).contiguous()
    nm = gi.shape[1] if gi is not None else (hi.shape[1] if hi is not None else min(nG, nH))
    if offG is not None:
        offG = _dev_f32(offG, "offG", 3)
        offH = _dev_f32(offH, "offH", 3)
    T = _out(buf, "T", (B, nm, 4, 4), torch.float32, G.device)
    with torch.cuda.device(G.device):
        rc = _lib.lib().ume_rigid_solve_f32(_ptr(G), _ptr(H), _ptr(gi), _ptr(hi), _ptr(offG), _ptr(offH), B, nG, nH,
                                            nm, C, _ptr(T), _stream())
    _lib.check(rc, "rigid_solve")
    return T


def rigid_solve_backward(G, H, gT):
    """Gradient of `rigid_solve(G[None], H[None])[0]` (pairing G[i] <-> H[i]) wrt G and H: G, H (nb,C,4),
    gT (nb,4,4) -> (gG, gH).  The backward of utils/loc_utils.py:292-335 for loss.py:137-190."""
    G = _dev_f32(G, "G", 3)
    H = _dev_f32(H, "H", 3)
    gT = _dev_f32(gT, "gT", 3)
    nb, C, _ = G.shape
    if H.shape != G.shape or G.shape[2] != 4 or tuple(gT.shape) != (nb, 4, 4):
        raise ValueError("rigid_solve_backward: shapes %s, %s, %s do not agree" % (tuple(G.shape), tuple(H.shape), tuple(gT.shape)))
    gG, gH = torch.empty_like(G), torch.empty_like(H)
    with torch.cuda.device(G.device):
        rc = _lib.lib().ume_rigid_solve_backward_f32(_ptr(G), _ptr(H), _ptr(gT), nb, C, _ptr(gG), _ptr(gH), _stream())
    _lib.check(rc, "rigid_solve_backward")
    return gG, gH


def ume_cdist_backward(ume1, ume2, Qt1, Qt2, D, gD):
    """Gradient of `ume_cdist(ume1, ume2)` wrt both arguments: ume1 (B,n1,C,4), ume2 (B,n2,C,4), their bases
    Qt1 / Qt2 from `ume_descriptors`, the forward's D and its gradient gD (B,n1,n2) -> (g_ume1, g_ume2).
    The backward of utils/loc_utils.py:8-15 for loss.py:84-118."""
    ume1 = _dev_f32(ume1, "ume1", 4)
    ume2 = _dev_f32(ume2, "ume2", 4)
    Qt1, Qt2 = _dev_f32(Qt1, "Qt1", 4), _dev_f32(Qt2, "Qt2", 4)
    D, gD = _dev_f32(D, "D", 3), _dev_f32(gD, "gD", 3)
    B, n1, C, _ = ume1.shape
    n2 = ume2.shape[1]
    if (ume2.shape[0] != B or ume2.shape[2] != C or tuple(Qt1.shape) != (B, n1, 4, C) or tuple(Qt2.shape) != (B, n2, 4, C)
            or tuple(D.shape) != (B, n1, n2) or tuple(gD.shape) != (B, n1, n2)):
        raise ValueError("ume_cdist_backward: shapes do not agree")
    g1, g2 = torch.empty_like(ume1), torch.empty_like(ume2)
    L = _lib.lib()
    with torch.cuda.device(ume1.device):
        ws = _workspace(L.ume_cdist_backward_workspace_bytes(B, n1, n2, C), ume1.device)
        rc = L.ume_cdist_backward_f32(_ptr(ume1), _ptr(ume2), _ptr(Qt1), _ptr(Qt2), _ptr(D), _ptr(gD), B, n1, n2, C,
                                      _ptr(g1), _ptr(g2), _ptr(ws), ws.numel(), _stream())
    _lib.check(rc, "ume_cdist_backward")
    return g1, g2


def batch_estimate_transform_ume_old(G, H):
    """utils/loc_utils.py:292-350 `batch_estimate_transform_ume_old(G, H)`: G, H (bs, C, 4) ->
    (T (bs,4,4), D (bs,)) with T[:3,:3] = R^T, T[:3,3] = b2 and D = 0.707 |P_H - P_G|_F."""
    if G.dim() != 3 or H.dim() != 3 or G.shape != H.shape or G.shape[-1] != 4:
        raise ValueError("batch_estimate_transform_ume_old: G %s, H %s must both be (bs, C, 4)" %
                         (tuple(G.shape), tuple(H.shape)))
    G = _dev_f32(G, "G")
    H = _dev_f32(H, "H")
    bs, C, _ = G.shape
    T = rigid_solve(G[None], H[None])[0]
    Dp = torch.empty((bs,), dtype=torch.float32, device=G.device)
    if bs:
        Qg = ume_descriptors(G)
        Qh = ume_descriptors(H)
        with torch.cuda.device(G.device):
            rc = _lib.lib().ume_pair_dist_f32(_ptr(Qh), _ptr(Qg), bs, C, 0.707, _ptr(Dp), _stream())
        _lib.check(rc, "batch_estimate_transform_ume_old")
    return T, Dp


def _rot_operand(R, name):
    """(pointer-ready tensor, stride in floats) of a batch of 3x3 rotations: packed (n,3,3) arrays and
    the rotation blocks of (n,4,4) transforms (`T[:, :3, :3]`) are read in place."""
    if not isinstance(R, torch.Tensor) or R.dim() != 3 or tuple(R.shape[1:]) != (3, 3):
        raise ValueError("%s must be (B,3,3), got %s" % (name, tuple(getattr(R, "shape", ()))))
    if not R.is_cuda:
        raise RuntimeError("%s is on %s: umeregrobust_b200 only runs on CUDA devices (no CPU fallback)" % (name, R.device))
    if R.dtype == torch.float32 and R.stride(1) == 4 and R.stride(2) == 1 and (R.shape[0] <= 1 or R.stride(0) == 16):
        return R, 16
    return R.float().contiguous(), 9


def relative_rotation_error(R, R_hat):
    """utils/eval_utils.py:60-76: degrees, acos((clamp(tr(R_hat R^T), -1, 3) - 1) / 2) * 180/pi.
    R, R_hat (B,3,3) on the device (views into (B,4,4) transforms are read in place) -> (B,)."""
    R, sr = _rot_operand(R, "R")
    R_hat, sh = _rot_operand(R_hat, "R_hat")
    if R.shape[0] != R_hat.shape[0]:
        raise ValueError("relative_rotation_error: batch sizes differ")
    out = torch.empty((R.shape[0],), dtype=torch.float32, device=R.device)
    with torch.cuda.device(R.device):
        rc = _lib.lib().ume_rotation_error_deg_f32(_ptr(R), _ptr(R_hat), R.shape[0], sr, sh, _ptr(out), _stream())
    _lib.check(rc, "relative_rotation_error")
    return out


# ----------------------------------------------------------------------------- ume_kp_layer
def ball_query_gather(pts, idx):
    """utils/loc_utils.py:353-354: rows of pts at idx with a zero row for idx == -1."""
    pad = torch.cat((torch.zeros((pts.shape[0], 1, pts.shape[-1]), device=pts.device, dtype=pts.dtype), pts), 1)
    return knn_gather(pad, idx + 1)


class ume_kp_layer(torch.nn.Module):
    """utils/loc_utils.py:357-431.  Same constructor and forward signature; forward returns
    (T, D, G_kp, H_kp) with T (bs,n_kp[,n_kp],4,4) and D (bs,n_kp[,n_kp])."""

    def __init__(self, ume_knn, ume_desc_rad, diag_only=False, n_rand=None):
        super().__init__()
        self.ume_knn = ume_knn
        self.ume_desc_rad = ume_desc_rad
        self.diag_only = diag_only
        self.n_rand = n_rand

    def ume_mat(self, points, features, bs, n_kp):
        """:365-372 on pre-gathered, zero-padded neighbourhoods (bs*n_kp, K, 3/C)."""
        m0 = torch.sum(features, dim=1, keepdim=True)
        m1 = features.transpose(2, 1) @ points
        mat = torch.cat((m0.transpose(2, 1), m1), dim=2) / (torch.sum(m0, dim=-1, keepdim=True) + 1e-6)
        return mat.view(bs, n_kp, *mat.shape[1:])

    def batch_keypoints(self, points, idx):
        out = ball_query_gather(points, idx)
        return out.view(-1, *out.shape[2:])

    def forward(self, source_points, source_features, source_kp, target_points, target_features, target_kp):
        bs, n_kp = source_kp.shape[0], source_kp.shape[1]
        # :383-393 ball_query + gathers + ume_mat, fused (no (bs,n_kp,K,C) tensor)
        pair = ume_moments_pair(source_points, source_kp, source_features, target_points, target_kp, target_features,
                                self.ume_knn, self.ume_desc_rad)              # one grid build + one launch for both sides
        if pair is not None:
            G, H, _ = pair
        else:
            G = ume_moments(source_points, source_kp, source_features, self.ume_knn, self.ume_desc_rad)
            H = ume_moments(target_points, target_kp, target_features, self.ume_knn, self.ume_desc_rad)
        C = G.shape[-2]
        if self.n_rand is not None:
            # :406-410 random triplet sums (host RNG, "only valid for batch size of one")
            import numpy as np
            if not self.diag_only:
                Gf = G.unsqueeze(2).expand(bs, n_kp, n_kp, C, 4).reshape(-1, C, 4)
                Hf = H.unsqueeze(1).expand(bs, n_kp, n_kp, C, 4).reshape(-1, C, 4)
            else:
                Gf, Hf = G.reshape(-1, C, 4), H.reshape(-1, C, 4)
            tri = torch.from_numpy(np.random.choice(np.arange(Gf.shape[0]), (self.n_rand, 3))).to(G.device)
            Gf = Gf[tri[:, 0]] + Gf[tri[:, 1]] + Gf[tri[:, 2]]
            Hf = Hf[tri[:, 0]] + Hf[tri[:, 1]] + Hf[tri[:, 2]]
            T, D = batch_estimate_transform_ume_old(Gf, Hf)
            return T.view(bs, -1, 4, 4), D.view(bs, -1), G.unsqueeze(2).squeeze(), H.unsqueeze(1).squeeze()
        if self.diag_only:
            T = rigid_solve(G, H)                                          # pairs i <-> i
            Qg, Qh = ume_descriptors(G), ume_descriptors(H)
            D = torch.empty((bs, n_kp), dtype=torch.float32, device=G.device)
            with torch.cuda.device(G.device):
                rc = _lib.lib().ume_pair_dist_f32(_ptr(Qh), _ptr(Qg), bs * n_kp, C, 0.707, _ptr(D), _stream())
            _lib.check(rc, "ume_kp_layer")
        else:
            # all n_kp^2 pairs (:394-397) without materialising the broadcast (bs*n_kp^2, C, 4) tensors
            ar = torch.arange(n_kp, device=G.device)
            gi = ar.repeat_interleave(n_kp).unsqueeze(0).expand(bs, -1).contiguous()
            hi = ar.repeat(n_kp).unsqueeze(0).expand(bs, -1).contiguous()
            T = rigid_solve(G, H, gi, hi).view(bs, n_kp, n_kp, 4, 4)
            D, _, _ = descriptor_cdist(ume_descriptors(G), ume_descriptors(H))
            D = D * (0.707 * 2.0 ** 0.5)                                    # :344 uses 0.707, not 1/sqrt(2)
        return T, D, G.unsqueeze(2).squeeze(), H.unsqueeze(1).squeeze()


# ----------------------------------------------------------------------------- hypothesis selection (f1)
def feature_spatial_var(pts, feat, knn=10):
    """utils/loc_utils.py:579-585 `feature_spatial_var(pts, feat, knn)`: for every point the mean
    feature distance to its knn-1 nearest other points.  pts (B,N,3), feat (B,N,C) -> (B,N)."""
    pts = _dev_f32(pts, "pts", 3)
    feat = _dev_f32(feat, "feat", 3)
    B, N, _ = pts.shape
    C = feat.shape[2]
    if feat.shape[0] != B or feat.shape[1] != N:
        raise ValueError("feature_spatial_var: pts %s and feat %s do not agree" % (tuple(pts.shape), tuple(feat.shape)))
    out = torch.empty((B, N), dtype=torch.float32, device=pts.device)
    with torch.cuda.device(pts.device):
        L = _lib.lib()
        ws = _workspace(L.ume_feature_spatial_var_workspace_bytes(B, N), pts.device)
        rc = L.ume_feature_spatial_var_f32(_ptr(pts), _ptr(feat), B, N, C, int(knn), _flags(), _ptr(out), _ptr(ws),
                                           ws.numel(), _stream())
    _lib.check(rc, "feature_spatial_var")
    return out


def cauchy_kernel(e, k=0.1):
    """utils/loc_utils.py:588-589."""
    return 1 / (1 + (e / k) ** 2)


def correlation_scores(source_points, target_points, source_vals, target_vals, T, k=20, sigma=0.05):
    """Scores of ALL hypotheses T (n_hyp,4,4) in one launch: pc_corr_cost_pytorch3d
    (utils/loc_utils.py:621-631) + pc_corr (:592-619) without the (chunk, Ns, k, C) gathers.
    source_points (Ns,3), target_points (Nt,3), *_vals (N*,C).  Returns (scores (n_hyp,), best int64 ())."""
    sp = _dev_f32(source_points, "source_points", 2)
    tp = _dev_f32(target_points, "target_points", 2)
    sv = _dev_f32(source_vals, "source_vals", 2)
    tv = _dev_f32(target_vals, "target_vals", 2)
    T = _dev_f32(T, "T", 3)
    Ns, Nt, C, nh = sp.shape[0], tp.shape[0], sv.shape[1], T.shape[0]
    if sv.shape[0] != Ns or tv.shape[0] != Nt or tv.shape[1] != C or tuple(T.shape[1:]) != (4, 4):
        raise ValueError("correlation_scores: shapes do not agree")
    score = torch.empty((nh,), dtype=torch.float32, device=sp.device)
    best = torch.zeros((), dtype=torch.int64, device=sp.device)
    with torch.cuda.device(sp.device):
        L = _lib.lib()
        ws = _workspace(L.ume_corr_scores_workspace_bytes(Ns, Nt, nh), sp.device)
        rc = L.ume_corr_scores_f32(_ptr(sp), _ptr(tp), _ptr(sv), _ptr(tv), _ptr(T), Ns, Nt, C, nh, int(k), float(sigma),
                                   _flags(), _ptr(score), _ptr(best), _ptr(ws), ws.numel(), _stream())
    _lib.check(rc, "correlation_scores")
    return score, best


def pc_corr_cost_pytorch3d(x1, x2, source_points, target_points, k, source_vals, target_vals, sigma, P=None,
                           use_norm=False, src_norm=None, tgt_norm=None, dev="cpu"):
    """utils/loc_utils.py:621-631: scores of the hypotheses (R = x1 (b,3,3), t = x2 (b,3))."""
    if P is not None or use_norm:
        raise NotImplementedError("pc_corr_cost_pytorch3d: P / use_norm are never used by the reference's scripts")
    b = x1.shape[0]
    T = torch.zeros((b, 4, 4), dtype=torch.float32, device=source_points.device)
    T[:, :3, :3] = x1
    T[:, :3, 3] = x2
    T[:, 3, 3] = 1
    return correlation_scores(source_points, target_points, source_vals, target_vals, T, k, sigma)[0]


def weighted_features(feat, mean, weight):
    """(feat - mean) * weight[..., None]  (utils/loc_utils.py:649-650); feat (N,C), mean (C,), weight (N,)."""
    feat = _dev_f32(feat, "feat", 2)
    mean = _dev_f32(mean, "mean", 1)
    weight = _dev_f32(weight, "weight", 1)
    out = torch.empty_like(feat)
    with torch.cuda.device(feat.device):
        rc = _lib.lib().ume_weight_features_f32(_ptr(feat), _ptr(mean), _ptr(weight), feat.shape[0], feat.shape[1],
                                                _ptr(out), _stream())
    _lib.check(rc, "weighted_features")
    return out


class FeatureCorrelator:
    """utils/loc_utils.py:634-681, same constructor and `feature_corr_hypothesis_test` signature.
    `batch` (the reference's chunk size) is accepted and ignored: all hypotheses are scored by one
    launch."""

    def __init__(self, n_clusters=8, batch=1, n_hypotheses=1, sigma=0.05, P=None, corr_num_nn=20):
        self.n_clusters = n_clusters
        self.batch = batch
        self.sigma = sigma
        self.n_hypotheses = n_hypotheses
        self.P = P
        self.corr_num_nn = corr_num_nn

    def scores(self, source_pc, target_pc, source_feat, target_feat, T_kp):
        """Scores of every hypothesis (the `mmf_score` of :663) and the index of the best."""
        if self.P is not None:
            raise NotImplementedError("FeatureCorrelator: P is never set by the reference's scripts")
        # :646 mean feature over both clouds (a C-vector; one torch reduction)
        m = torch.mean(torch.cat((source_feat, target_feat), dim=1), dim=1)[0]
        src_w = feature_spatial_var(source_pc, source_feat, knn=50)[0]          # :647
        tgt_w = feature_spatial_var(target_pc, target_feat, knn=50)[0]          # :648
        wsf = weighted_features(source_feat[0], m, src_w)                        # :649
        wtf = weighted_features(target_feat[0], m, tgt_w)                        # :650
        return correlation_scores(source_pc[0], target_pc[0], wsf, wtf, T_kp, self.corr_num_nn, self.sigma)

    def feature_corr_hypothesis_test(self, source_pc, target_pc, source_feat, target_feat, T_kp, src_norm=None,
                                     tgt_norm=None):
        """source_pc (1,Ns,3), target_pc (1,Nt,3), *_feat (1,N*,C), T_kp (n_hyp,4,4) -> best T (4,4)."""
        score, best = self.scores(source_pc, target_pc, source_feat, target_feat, T_kp)
        return T_kp[best]


# ----------------------------------------------------------------------------- voxel de-duplication (f4)
def sparse_quantize(coordinates, features=None, return_index=False, quantization_size=None):
    """MinkowskiEngine `ME.utils.sparse_quantize` as evaluate.py:261-264 calls it: integer voxel
    coordinates floor(coordinates / quantization_size) with duplicates removed — the first row of
    every occupied voxel survives, survivors keep their row order.  coordinates (N,3) float32 on
    the device.  Returns `unique_coords (M,3) int32` and, with return_index=True, the surviving
    rows `(M,) int64` (ascending).  Reading M back is a sync point, as in the reference."""
    if features is not None:
        raise NotImplementedError("sparse_quantize: the feature-averaging mode is not used by evaluate.py")
    if quantization_size is None:
        quantization_size = 1.0
    c = _dev_f32(coordinates, "coordinates", 2)
    if c.shape[1] != 3:
        raise ValueError("sparse_quantize: coordinates must be (N,3)")
    N = c.shape[0]
    index = torch.empty((max(N, 1),), dtype=torch.int64, device=c.device)
    coords = torch.empty((max(N, 1), 3), dtype=torch.int32, device=c.device)
    count = torch.empty((1,), dtype=torch.int32, device=c.device)
    with torch.cuda.device(c.device):
        L = _lib.lib()
        ws = _workspace(L.ume_voxel_unique_workspace_bytes(N), c.device)
        rc = L.ume_voxel_unique_f32(_ptr(c), N, float(quantization_size), _ptr(index), _ptr(coords), _ptr(count),
                                    _ptr(ws), ws.numel(), _stream())
    _lib.check(rc, "sparse_quantize")
    M = int(count.item())
    if M < 0:
        raise ValueError("sparse_quantize: a coordinate is NaN or outside +-2^20 voxels")
    return (coords[:M], index[:M]) if return_index else coords[:M]


def select_hypothesis(src_pts_raw, tgt_pts_raw, src_pts, tgt_pts, src_feat, tgt_feat, hypotheses, corr_sigma,
                      corr_ds=0.3, tgt_ds=0.3, pc_corr_max_size=30000, corr_num_nn=20, src_rows=None, tgt_rows=None,
                      generator=None):
    """evaluate.py:259-296 for one pair, stream-ordered on the device: voxel de-duplication of the
    raw clouds (:261-264), nearest-row feature transfer from the voxel clouds the backbone saw
    (:272-275), random down-sampling to `pc_corr_max_size` (:278-285) and the correlator's pick
    among the hypotheses (pc_fcht :20-47).
      src_pts_raw (Nr,3), tgt_pts_raw (Nr',3): raw clouds; src_pts (1,N,3), src_feat (1,N,C) (and
      tgt_*): the clouds the features live on; hypotheses (n_hyp,4,4).
      src_rows / tgt_rows: the down-sampling draw as data (the reference uses the host RNG,
      np.random.choice, which cannot be reproduced bit for bit); default: torch.randperm.
    Returns (T (4,4), best index (0-dim int64), scores (n_hyp,))."""
    def prep(raw, q, pts, feat, rows):
        raw = _dev_f32(raw, "raw cloud", 2)
        _, keep = sparse_quantize(raw, return_index=True, quantization_size=q)
        ds = raw[keep][None]
        f = knn1_transfer(ds, pts, feat)
        n = min(int(pc_corr_max_size), ds.shape[1])
        if rows is None:
            rows = torch.randperm(ds.shape[1], device=ds.device, generator=generator)[:n]
        return ds[:, rows].contiguous(), f[:, rows].contiguous()
    sp, sf = prep(src_pts_raw, corr_ds, src_pts, src_feat, src_rows)
    tp, tf = prep(tgt_pts_raw, tgt_ds, tgt_pts, tgt_feat, tgt_rows)
    corr = FeatureCorrelator(sigma=corr_sigma, corr_num_nn=corr_num_nn)
    score, best = corr.scores(sp, tp, sf, tf, hypotheses)
    return hypotheses[best], best, score


# ----------------------------------------------------------------------------- Hungarian option (f4)
def linear_sum_assignment(cost):
    """scipy.optimize.linear_sum_assignment for one (n1,n2) float32 cost matrix on the HOST (a CPU
    tensor or numpy array), as evaluate.py:219 calls it: (row_ind, col_ind) int64 numpy arrays,
    row_ind ascending."""
    import numpy as np
    c = np.ascontiguousarray(cost.detach().cpu().numpy() if torch.is_tensor(cost) else cost, dtype=np.float32)
    if c.ndim != 2:
        raise ValueError("linear_sum_assignment: expected a matrix")
    k = min(c.shape)
    rows, cols = np.empty(k, np.int64), np.empty(k, np.int64)
    rc = _lib.lib().ume_linear_sum_assignment_host_f32(c.ctypes.data, c.shape[0], c.shape[1], rows.ctypes.data, cols.ctypes.data)
    _lib.check(rc, "linear_sum_assignment")
    return rows, cols


def hungarian_match(D):
    """evaluate.py:216-222 (`hungarian_matching_flag`): one-to-one matches minimising the summed
    distance.  D (B,n1,n2) on the device -> m (B, min(n1,n2), 2) int64 on the device.  Like the
    reference, the assignment itself is solved on the host from a copy of D (one D2H per call)."""
    Dh = D.detach().float().cpu()
    out = torch.empty((D.shape[0], min(D.shape[1], D.shape[2]), 2), dtype=torch.int64)
    for b in range(D.shape[0]):
        r, c = linear_sum_assignment(Dh[b])
        out[b, :, 0] = torch.from_numpy(r)
        out[b, :, 1] = torch.from_numpy(c)
    return out.to(D.device)


# ----------------------------------------------------------------------------- match sub-sampling (f2)
_subsample_calls = [0]


def weighted_match_subsample(ume_d, tau, num_samples, generator=None, u=None, seed=None, buf=None):
    """Device-side equivalent of evaluate.py:233-245: draw `num_samples` of the n matches WITHOUT
    replacement with probability proportional to exp((1 - d) / tau).  The reference does this on
    the host with np.random.choice (a D2H sync per pair); here it is the Gumbel-top-k trick in one
    kernel (`ume_gumbel_topk_f32`: Philox counter RNG, exact radix select per pair) — the same
    distribution (successive sampling without replacement == top-k of log-weight + Gumbel noise), but
    not the same random stream, so parity tests feed the uniforms `u` (same shape as ume_d) in as data.
    ume_d (n,) or (B,n) -> int64 indices (num_samples,) or (B,num_samples), ascending.
    `seed`: Philox key; default: drawn from `generator` (a CPU torch.Generator) or derived from
    torch.initial_seed() and a call counter, so that torch.manual_seed() makes runs repeatable."""
    d = _dev_f32(ume_d if ume_d.dim() == 2 else ume_d[None], "ume_d", 2)
    B, n = d.shape
    k = min(int(num_samples), n)
    if u is not None:
        u = _dev_f32(u if u.dim() == 2 else u[None], "u", 2)
        if tuple(u.shape) != (B, n):
            raise ValueError("weighted_match_subsample: u %s does not match ume_d %s" % (tuple(u.shape), (B, n)))
    if seed is None:
        if generator is not None:
            seed = int(torch.randint(0, 2 ** 62, (1,), generator=generator if generator.device.type == "cpu" else None).item())
        else:
            _subsample_calls[0] += 1
            seed = (torch.initial_seed() * 0x9E3779B97F4A7C15 + _subsample_calls[0]) & 0xFFFFFFFFFFFFFFFF
    idx = _out(buf, "subsample_idx", (B, k), torch.int64, d.device)
    with torch.cuda.device(d.device):
        rc = _lib.lib().ume_gumbel_topk_f32(_ptr(d), _ptr(u), B, n, k, float(tau), int(seed) & 0xFFFFFFFFFFFFFFFF, _ptr(idx),
                                            _stream())
    _lib.check(rc, "weighted_match_subsample")
    return idx if ume_d.dim() == 2 else idx[0]


# ----------------------------------------------------------------------------- fused hot path
def register_hypotheses(src_pts, src_feat, src_kp, tgt_pts, tgt_feat, tgt_kp, K, radius, want_D=False,
                        centered=True, buf=None, matching="argmin", subsample=None, tau=0.05, subsample_u=None,
                        subsample_seed=None):
    """evaluate.py:206-257 for a whole batch, without the host-RNG sub-sampling (:233-245): UME
    matrices for both clouds, subspace distances with fused arg-min, one rigid hypothesis per
    source keypoint from its best-matching target keypoint.

    centered=True runs descriptors and the solve on the keypoint-centred moments (same column
    space / same transform, smaller numbers, closer to the exact answer than the reference's own
    fp32); centered=False follows the reference's absolute-coordinate arithmetic.
    `buf`: an arena dict; when given, every output lives in it and is REUSED by the next call with
    the same shapes (the returned tensors are then only valid until that next call).
    matching="hungarian" (evaluate.py:216-222, `hungarian_matching_flag`): one-to-one matches from the
    host assignment solver instead of the row arg-min (one D2H copy of D per call, like the reference).
    subsample=m (evaluate.py:233-245, `filter_by_ume_dist_cond` with ume_n_samples = m): the solve runs on
    m of the n matches drawn without replacement with probability ~ exp((1 - d)/tau) on the device
    (`weighted_match_subsample`); match / dmin / T then have m rows per pair.
    Returns dict(F_src, F_tgt, match (B,n,2) int64, dmin (B,n), T (B,n,4,4), D or None)."""
    # both sides in one grid build and one moment launch where the pair entry applies (same shapes, warp kernel)
    pair = ume_moments_pair(src_pts, src_kp, src_feat, tgt_pts, tgt_kp, tgt_feat, K, radius, return_centered=centered, buf=buf)
    if centered:
        if pair is not None:
            F_src, Fc_src, F_tgt, Fc_tgt, both = pair
        else:
            F_src, Fc_src = ume_moments(src_pts, src_kp, src_feat, K, radius, return_centered=True, buf=buf, tag="_src")
            F_tgt, Fc_tgt = ume_moments(tgt_pts, tgt_kp, tgt_feat, K, radius, return_centered=True, buf=buf, tag="_tgt")
        A, Bm = Fc_src, Fc_tgt
    else:
        if pair is not None:
            F_src, F_tgt, both = pair
        else:
            F_src = ume_moments(src_pts, src_kp, src_feat, K, radius, buf=buf, tag="_src")
            F_tgt = ume_moments(tgt_pts, tgt_kp, tgt_feat, K, radius, buf=buf, tag="_tgt")
        A, Bm = F_src, F_tgt
    if matching not in ("argmin", "hungarian"):
        raise ValueError("register_hypotheses: matching must be 'argmin' or 'hungarian'")
    C = A.shape[-2]
    if C in (32, 64) and config["cdist_impl"] in (None, 1):
        # tensor-core distance kernel, operands written pre-split by the orthonormalisation kernel
        if pair is not None:
            # (the two sides are halves of one allocation: one orthonormalisation launch for both)
            nb = A.shape[0]
            Qh = ume_descriptors_split(both, buf=buf, tag="_pair")
            Qh_src, Qh_tgt = Qh[:nb], Qh[nb:]
        else:
            Qh_src, Qh_tgt = ume_descriptors_split(A, buf=buf, tag="_src"), ume_descriptors_split(Bm, buf=buf, tag="_tgt")
        D, am, dm = descriptor_cdist_split(Qh_src, Qh_tgt, want_D=want_D or matching == "hungarian", want_argmin=True, buf=buf)
    else:
        D, am, dm = descriptor_cdist(ume_descriptors(A, buf=buf, tag="_src"), ume_descriptors(Bm, buf=buf, tag="_tgt"),
                                     want_D=want_D or matching == "hungarian", want_argmin=True, buf=buf)
    if matching == "hungarian":
        m = hungarian_match(D)
        gi, hi = m[..., 0].contiguous(), m[..., 1].contiguous()
        T = rigid_solve(A, Bm, gi, hi, src_kp if centered else None, tgt_kp if centered else None, buf=buf)
        dsel = torch.gather(torch.gather(D, 1, gi[..., None].expand(-1, -1, D.shape[2])), 2, hi[..., None])[..., 0]
        return dict(F_src=F_src, F_tgt=F_tgt, match=m, dmin=dsel, T=T, D=D)
    if subsample is not None and int(subsample) < am.shape[1]:
        sel = weighted_match_subsample(dm, tau, int(subsample), u=subsample_u, seed=subsample_seed, buf=buf)   # (B,m) ascending
        hi_sel = torch.gather(am, 1, sel)
        T = rigid_solve(A, Bm, sel, hi_sel, src_kp if centered else None, tgt_kp if centered else None, buf=buf)
        match = _out(buf, "match_sub", tuple(sel.shape) + (2,), torch.int64, am.device)
        match[..., 0] = sel
        match[..., 1] = hi_sel
        return dict(F_src=F_src, F_tgt=F_tgt, match=match, dmin=torch.gather(dm, 1, sel), T=T, D=D)
    if centered:
        T = rigid_solve(A, Bm, None, am, src_kp, tgt_kp, buf=buf)
    else:
        T = rigid_solve(A, Bm, None, am, buf=buf)
    B, n = am.shape
    if buf is not None and "arange" in buf and tuple(buf["arange"].shape) == (B, n) and buf["arange"].device == am.device:
        ar = buf["arange"]
    else:
        ar = torch.arange(n, device=am.device).expand(B, n)
        if buf is not None:
            buf["arange"] = ar
    match = _out(buf, "match", (B, n, 2), torch.int64, am.device)
    match[..., 0] = ar
    match[..., 1] = am
    return dict(F_src=F_src, F_tgt=F_tgt, match=match, dmin=dm, T=T, D=D)
