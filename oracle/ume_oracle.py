"""ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy restatement of the reference's UME descriptor-and-registration hot path
(yuvalH9/UMERegRobust).  Every function cites the reference file:line it follows.  The
restatement is written against numpy (fp32 or fp64 selectable) instead of torch so that it can
serve as an fp64 ground truth as well as an fp32 mirror of the reference.

Parity status: the reference ships no tests, fixtures or golden vectors for this path, so the
oracle is pinned by (1) tests/golden/*.npz — outputs of the UNMODIFIED reference functions,
imported in the build container by tests/golden/make_golden.py — and (2) analytic known-answer
tests (tests/test_oracle.py).  pytorch3d ops underneath are restated in pytorch3d_ops.py.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package must never route through it.
"""
import numpy as np

from . import pytorch3d_ops as p3d


# ----------------------------------------------------------------------------- moments
def moments_from_neighbors(nn_pts, nn_feat, normalise=True, eps=1e-6):
    """UME moment matrix from already-gathered (zero-padded) neighbourhoods.

    Follows evaluate.py:56-59 (same math: utils/loc_utils.py:365-372 `ume_kp_layer.ume_mat`,
    :434-445 `create_local_ume_matrix` with normalise=False).
    nn_pts (..., K, 3), nn_feat (..., K, C)  ->  F (..., C, 4) = [sum_k f | sum_k f x^T] / s,
    s = sum_c sum_k f_kc + 1e-6.
    """
    ft = np.swapaxes(nn_feat, -1, -2)                      # (..., C, K)
    first = ft @ nn_pts                                    # (..., C, 3)
    zeroth = ft.sum(axis=-1, keepdims=True)                # (..., C, 1)
    F = np.concatenate([zeroth, first], axis=-1)
    if normalise:
        F = F / (zeroth.sum(axis=-2, keepdims=True) + np.asarray(eps, dtype=F.dtype))
    return F


def ume_moments(pts, kpts, feat, K, radius, dtype=np.float32, fma=False, use_c=True,
                return_idx=False, chunk=256):
    """evaluate.py:50-60 `my_ume_generation`: ball-query + pad/gather + moment build.

    pts (B,N,3), kpts (B,n,3), feat (B,N,C) -> F (B,n,C,4).  Neighbour search is always fp32
    (that is what defines the neighbour set); the sums run in `dtype`.  Work is chunked over
    keypoints so the (n,K,C) gather the reference materialises stays small.
    """
    pts32 = np.ascontiguousarray(pts, dtype=np.float32)
    kp32 = np.ascontiguousarray(kpts, dtype=np.float32)
    B, N, _ = pts32.shape
    n = kp32.shape[1]
    C = feat.shape[-1]
    bq = (p3d.ball_query_c(kp32, pts32, K, radius, return_nn=False, fma=fma) if use_c
          else p3d.ball_query_np(kp32, pts32, K, radius, return_nn=False))
    idx = bq.idx
    # evaluate.py:52-53: -1 -> index of an appended all-zero feature row
    feat_pad = np.concatenate([feat, np.zeros_like(feat[:, :1])], axis=1).astype(dtype)
    pts_pad = np.concatenate([pts32, np.zeros_like(pts32[:, :1])], axis=1).astype(dtype)
    safe = np.where(idx < 0, N, idx)
    F = np.empty((B, n, C, 4), dtype=dtype)
    for b in range(B):
        for s in range(0, n, chunk):
            sel = safe[b, s:s + chunk]                        # (c,K)
            F[b, s:s + chunk] = moments_from_neighbors(pts_pad[b][sel], feat_pad[b][sel])
    return (F, idx) if return_idx else F


def ume_moments_backward(pts, kpts, grad_F, K, radius, dtype=np.float64, fma=False):
    """Gradient of the RAW moments F[i,c,:] = sum_{j in nbr(i)} feat[j,c] [1, pts[j]] with respect to
    feat: what torch autograd returns through the gather + matmul of utils/loc_utils.py:150-161.
    grad_F (B,n,C,4) -> (B,N,C)."""
    pts32 = np.ascontiguousarray(pts, dtype=np.float32)
    kp32 = np.ascontiguousarray(kpts, dtype=np.float32)
    idx = p3d.ball_query_c(kp32, pts32, K, radius, return_nn=False, fma=fma).idx
    B, N, _ = pts32.shape
    g = np.asarray(grad_F, dtype=dtype)
    out = np.zeros((B, N, g.shape[2]), dtype=dtype)
    for b in range(B):
        for i in range(idx.shape[1]):
            rows = idx[b, i][idx[b, i] >= 0]
            out[b, rows] += g[b, i, :, 0][None] + pts32[b, rows].astype(dtype) @ g[b, i, :, 1:].T
    return out


def normaliser_condition(feat, idx):
    """kappa[b,i] = sum|f| / |sum f + 1e-6| over keypoint i's neighbourhood: how much the division at
    evaluate.py:59 amplifies rounding (features are signed, so the sum cancels; SURVEY §7 'Normaliser
    can be <= 0').  Tests scale their F tolerance by it."""
    B, N, C = feat.shape
    row_abs = np.concatenate([np.abs(feat.astype(np.float64)).sum(-1), np.zeros((B, 1))], 1)
    row_sum = np.concatenate([feat.astype(np.float64).sum(-1), np.zeros((B, 1))], 1)
    safe = np.where(idx < 0, N, idx)
    b = np.arange(B)[:, None, None]
    return row_abs[b, safe].sum(-1) / np.abs(row_sum[b, safe].sum(-1) + 1e-6)


# ----------------------------------------------------------------------------- descriptor distance
def subspace_projector(F):
    """utils/loc_utils.py:9-12: thin QR of (..., C, 4), P = Q Q^T."""
    Q = np.linalg.qr(F, mode="reduced")[0]
    return Q @ np.swapaxes(Q, -1, -2), Q


def cdist_mm(a, b):
    """torch.cdist's matmul form (used by torch for >25 rows): sqrt(clamp(|a|^2+|b|^2-2ab, 0))."""
    a2 = (a * a).sum(-1)[..., :, None]
    b2 = (b * b).sum(-1)[..., None, :]
    d2 = a2 + b2 - 2.0 * (a @ np.swapaxes(b, -1, -2))
    return np.sqrt(np.maximum(d2, 0))


def ume_cdist(F1, F2, dtype=None):
    """utils/loc_utils.py:8-15 `ume_cdist`: D[b,i,j] = |P1_i - P2_j|_F / sqrt(2)."""
    if dtype is not None:
        F1, F2 = F1.astype(dtype), F2.astype(dtype)
    P1, _ = subspace_projector(F1)
    P2, _ = subspace_projector(F2)
    a = P1.reshape(*P1.shape[:-2], -1)
    b = P2.reshape(*P2.shape[:-2], -1)
    return cdist_mm(a, b) / np.sqrt(np.asarray(2.0, dtype=a.dtype))


def ume_cdist_gram(F1, F2, dtype=np.float64):
    """Same quantity through the Gram identity D^2 = 4 - |Q1^T Q2|_F^2."""
    Q1 = np.linalg.qr(F1.astype(dtype), mode="reduced")[0]
    Q2 = np.linalg.qr(F2.astype(dtype), mode="reduced")[0]
    B, n1, C, M = Q1.shape
    n2 = Q2.shape[1]
    A = np.swapaxes(Q1, -1, -2).reshape(B, n1 * M, C)
    Bm = np.swapaxes(Q2, -1, -2).reshape(B, n2 * M, C)
    S = (A @ np.swapaxes(Bm, -1, -2)).reshape(B, n1, M, n2, M)
    s = (S * S).sum(axis=(2, 4))
    return np.sqrt(np.maximum(M - s, 0))


def match_argmin(D):
    """evaluate.py:224-225: m = D.min(-1)[1] (first index on ties) paired with arange."""
    m = np.argmin(D, axis=-1)
    B, n = m.shape
    return np.stack([np.broadcast_to(np.arange(n), (B, n)), m], axis=-1).astype(np.int64)


# ----------------------------------------------------------------------------- rigid solve
def rigid_from_ume(G, H, dtype=None, with_distance=True):
    """utils/loc_utils.py:292-350 `batch_estimate_transform_ume_old(G, H)`.

    G, H (b, C, 4).  Column 0 holds the zeroth-order weights, columns 1..3 the first-order
    moments.  Returns T (b,4,4) and D (b,) = 0.707 |P_H - P_G|_F.
    (Empirically tgt ~= T[:3,:3] src + T[:3,3] when G = UME(src), H = UME(tgt))
    """
    if dtype is not None:
        G, H = G.astype(dtype), H.astype(dtype)
    dt = G.dtype
    b = G.shape[0]
    mg, mh = G[:, :, :1], H[:, :, :1]                         # :304-305
    g, h = G[:, :, 1:], H[:, :, 1:]                           # :308-309
    mg_sq = (mg * mg).sum(axis=1, keepdims=True) + dt.type(1e-16)   # :312
    mg_mh = (mg * mh).sum(axis=1, keepdims=True)              # :313
    gmg = (g * mg).sum(axis=1, keepdims=True)                 # :314
    hmg = (h * mg).sum(axis=1, keepdims=True)                 # :315
    wlc = gmg / (mg_sq + dt.type(1e-16))                      # :319
    wrc = hmg / (mg_mh + dt.type(1e-16))                      # :320
    left = g - wlc * mg                                       # :322
    right = h - wrc * mh                                      # :323
    M = np.swapaxes(right, 1, 2) @ left                       # :325
    U, S, VH = np.linalg.svd(np.swapaxes(M, 1, 2))            # :326
    fix = np.tile(np.eye(3, dtype=dt), (b, 1, 1))
    fix[:, 2, 2] = np.sign(np.linalg.det(U @ VH))             # :327-328
    R = U @ fix @ VH                                          # :329
    b2 = wrc - wlc @ R                                        # :332
    T = np.tile(np.eye(4, dtype=dt), (b, 1, 1))
    T[:, :3, :3] = np.swapaxes(R, 1, 2)                       # :348
    T[:, :3, 3] = b2[:, 0, :]                                 # :349
    if not with_distance:
        return T, None
    PH, _ = subspace_projector(H)                             # :338-339
    PG, _ = subspace_projector(G)                             # :341-342
    diff = PH - PG
    D = dt.type(0.707) * np.sqrt((diff * diff).sum(axis=(1, 2)))   # :344
    return T, D


def relative_rotation_error(R, R_hat):
    """utils/eval_utils.py:60-76: degrees, acos((clamp(tr(R_hat R^T), -1, 3) - 1) / 2)."""
    delta = R_hat @ np.swapaxes(R, 1, 2)
    tr = np.clip(np.trace(delta, axis1=1, axis2=2), -1, 3)
    return np.arccos((tr - 1) / 2) * (180.0 / 3.141592653589793)


def rotation_angle_rad(Ra, Rb):
    """Angle of Ra Rb^T, stable for tiny angles (|Ra - Rb|_F / sqrt(2) ~ angle)."""
    d = Ra - Rb
    return np.sqrt((d * d).sum(axis=(-1, -2)) / 2.0)


# ----------------------------------------------------------------------------- ume_kp_layer
def ume_kp_layer_forward(src_pts, src_feat, src_kp, tgt_pts, tgt_feat, tgt_kp, ume_knn,
                         ume_desc_rad, diag_only=False, dtype=np.float32, fma=False):
    """utils/loc_utils.py:380-431 `ume_kp_layer.forward` (n_rand=None).

    ball_query(return_nn=False) -> ball_query_gather (zero row for -1, :353-354) -> ume_mat
    (:365-372) -> all-pairs or diagonal rigid solve.  Returns (T, D, G_kp, H_kp).
    """
    bs, n_kp = src_kp.shape[0], src_kp.shape[1]
    G = ume_moments(src_pts, src_kp, src_feat, ume_knn, ume_desc_rad, dtype=dtype, fma=fma)
    H = ume_moments(tgt_pts, tgt_kp, tgt_feat, ume_knn, ume_desc_rad, dtype=dtype, fma=fma)
    C = G.shape[-2]
    if diag_only:
        Gf, Hf = G.reshape(-1, C, 4), H.reshape(-1, C, 4)
    else:
        Gb = np.broadcast_to(G[:, :, None], (bs, n_kp, n_kp, C, 4))
        Hb = np.broadcast_to(H[:, None, :], (bs, n_kp, n_kp, C, 4))
        Gf, Hf = Gb.reshape(-1, C, 4), Hb.reshape(-1, C, 4)
    T, D = rigid_from_ume(Gf, Hf)
    if diag_only:
        return T.reshape(bs, n_kp, 4, 4), D.reshape(bs, n_kp), G, H
    return T.reshape(bs, n_kp, n_kp, 4, 4), D.reshape(bs, n_kp, n_kp), G, H


# ----------------------------------------------------------------------------- hypothesis selection (f1)
def feature_spatial_var(pts, feat, knn=10, dtype=np.float32, fma=False):
    """utils/loc_utils.py:579-585: mean over the knn-1 nearest OTHER rows of |f_i - f_j|_2.
    pts (B,N,3), feat (B,N,C) -> (B,N).  Neighbour search in fp32 (pytorch3d), norms in `dtype`."""
    nn = p3d.knn_points_c(np.ascontiguousarray(pts, np.float32), np.ascontiguousarray(pts, np.float32), knn, fma=fma)
    f = feat.astype(dtype)
    nn_feat = p3d.knn_gather_np(f, nn.idx[:, :, 1:])                       # :581 drops the nearest (the point itself)
    diff = f[:, :, None, :] - nn_feat
    return np.sqrt((diff * diff).sum(-1)).mean(-1)


def cauchy_kernel(e, k=0.1):
    """utils/loc_utils.py:588-589."""
    return 1 / (1 + (e / k) ** 2)


def pc_corr_scores(src_pts, tgt_pts, vals_p, vals_q, T, k, sigma, dtype=np.float32, fma=False, chunk=16):
    """utils/loc_utils.py:621-631 + :592-619 for hypotheses T (n_hyp,4,4):
    transformed = src @ R^T + t (fp32, like the reference's batched matmul), K nearest target rows
    (pytorch3d knn, fp32), dist = |p - q| , weight = cauchy(dist; sigma), score = sum w <vp, vq> / Ns."""
    src32 = np.ascontiguousarray(src_pts, np.float32)
    tgt32 = np.ascontiguousarray(tgt_pts, np.float32)
    vp, vq = vals_p.astype(dtype), vals_q.astype(dtype)
    out = np.empty(len(T), dtype=dtype)
    for s in range(0, len(T), chunk):
        Tc = np.asarray(T[s:s + chunk], np.float32)
        moved = (src32[None] @ np.swapaxes(Tc[:, :3, :3], 1, 2) + Tc[:, None, :3, 3]).astype(np.float32)   # :626
        nn = p3d.knn_points_c(moved, np.broadcast_to(tgt32[None], (len(Tc),) + tgt32.shape).copy(), k, fma=fma)
        q = tgt32[nn.idx].astype(dtype)                                         # (c, Ns, k, 3)
        d = moved.astype(dtype)[:, :, None, :] - q
        dist = np.sqrt((d * d).sum(-1))                                         # :593
        w = cauchy_kernel(dist, dtype(sigma))                                   # :596
        prod = (vp[None, :, None, :] * vq[nn.idx]).sum(-1)                      # :604
        out[s:s + chunk] = (w * prod).sum(axis=(1, 2)) / dtype(vp.shape[0])     # :613-615
    return out


def feature_corr_hypothesis_test(src_pc, tgt_pc, src_feat, tgt_feat, T_kp, sigma=0.05, corr_num_nn=20,
                                 dtype=np.float32, fma=False):
    """utils/loc_utils.py:640-681 `FeatureCorrelator.feature_corr_hypothesis_test` (P=None, no
    normals): inputs (1,N,3)/(1,N,C)/(n_hyp,4,4).  Returns (best_T (4,4), scores (n_hyp,))."""
    m = np.concatenate([src_feat, tgt_feat], axis=1).astype(dtype).mean(axis=1)          # :646
    sw = feature_spatial_var(src_pc, src_feat, knn=50, dtype=dtype, fma=fma)                # :647
    tw = feature_spatial_var(tgt_pc, tgt_feat, knn=50, dtype=dtype, fma=fma)                # :648
    wsf = (src_feat.astype(dtype) - m) * sw[..., None]                                      # :649
    wtf = (tgt_feat.astype(dtype) - m) * tw[..., None]                                      # :650
    scores = pc_corr_scores(src_pc[0], tgt_pc[0], wsf[0], wtf[0], T_kp, corr_num_nn, sigma, dtype=dtype, fma=fma)
    return T_kp[int(np.argmax(scores))], scores                                            # :663-681


# ----------------------------------------------------------------------------- voxel de-duplication
def sparse_quantize(coords, quantization_size):
    """MinkowskiEngine 0.5.4 `ME.utils.sparse_quantize(coordinates, return_index=True,
    quantization_size=q)` (un-vendored dependency, requirements.txt; called at evaluate.py:261-264
    and kitti_dataset.py:416): discrete = floor(coordinates / q) in the input's float32 arithmetic,
    then the rows of the first occurrence of every distinct voxel.  ME's CPU coordinate map inserts
    rows in order, so the returned index list is ascending.  Returns (unique (M,3) int32, index (M,) int64)."""
    c = np.asarray(coords, dtype=np.float32)
    disc = np.floor(c / np.float32(quantization_size)).astype(np.int32)
    _, first = np.unique(disc, axis=0, return_index=True)
    first = np.sort(first).astype(np.int64)
    return disc[first], first


# ----------------------------------------------------------------------------- whole hot path
def register_pair_hypotheses(src_pts, src_feat, src_kp, tgt_pts, tgt_feat, tgt_kp, K, radius,
                             dtype=np.float32, fma=False):
    """evaluate.py:206-257 with the flags every shipped config uses except the host-RNG
    sub-sampling (:233-245): UME generation for both clouds, ume_cdist, arg-min matching, gather
    of matched UME matrices, one rigid hypothesis per match.

    Returns dict(F_src, F_tgt, D, match (B,n,2), T (B,n,4,4)).
    """
    F_src = ume_moments(src_pts, src_kp, src_feat, K, radius, dtype=dtype, fma=fma)
    F_tgt = ume_moments(tgt_pts, tgt_kp, tgt_feat, K, radius, dtype=dtype, fma=fma)
    D = ume_cdist(F_src, F_tgt)
    m = match_argmin(D)
    B, n = m.shape[:2]
    C = F_src.shape[-2]
    bidx = np.arange(B)[:, None]
    Gm = F_src[bidx, m[..., 0]]
    Hm = F_tgt[bidx, m[..., 1]]
    T, _ = rigid_from_ume(Gm.reshape(-1, C, 4), Hm.reshape(-1, C, 4), with_distance=False)
    return dict(F_src=F_src, F_tgt=F_tgt, D=D, match=m, T=T.reshape(B, n, 4, 4))
