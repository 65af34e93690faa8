"""Training-time UME generation and losses: mirrors of
`utils.loc_utils.generate_ume_from_keypoints2` (utils/loc_utils.py:86-188), `loss.UMEContrastiveLoss`
(loss.py:49-118) and `loss.CubeRegistrationLoss` (loss.py:121-190) with the same signatures and
return values.

What runs where:
  * the neighbourhood work — the K = 1 intersection queries, the dense-neighbourhood filter and the
    UME moment build with up to `max_nn` = 5000 neighbours per keypoint — runs in this repo's CUDA
    kernels; the moment build is a `torch.autograd.Function` whose backward is the scatter kernel
    `ume_moments_backward_f32`, so the reference's (bs, n, max_nn, C) gather tensor (2.6 GB at the
    shipped training config) exists neither in the forward nor in the backward pass;
  * the projector distance (thin QR -> P = Q Q^T -> cdist) and the 3x3 SVD solve downstream of the
    (bs, n, C, 4) moment matrices are `torch.autograd.Function`s too: forward = the inference kernels,
    backward = `ume_cdist_backward_f32` / `ume_rigid_solve_backward_f32` (closed forms on the 4x4 Gram
    blocks and on the proper singular frames of the 3x3 cross moment; no (bs,n,C,C) projector, no
    cuSOLVER autograd).  Only the soft-max / mean reductions of the losses themselves are torch ops.
"""
import numpy as np
import torch

from . import api


class _RawMoments(torch.autograd.Function):
    """F[b,i,c,:] = sum over the first K rows j of pts[b] within `radius` of kpts[b,i] of
    feat[b,j,c] * [1, pts[b,j,:]].  Differentiable in `feat` only (points are data)."""

    @staticmethod
    def forward(ctx, feat, pts, kpts, K, radius):
        ctx.save_for_backward(pts, kpts)
        ctx.K, ctx.radius = int(K), float(radius)
        return api.ume_moments(pts, kpts, feat.detach(), K, radius, raw=True)

    @staticmethod
    def backward(ctx, grad_F):
        pts, kpts = ctx.saved_tensors
        return api.ume_moments_backward(pts, kpts, grad_F.contiguous(), ctx.K, ctx.radius), None, None, None, None


def ume_moments_autograd(pts, kpts, feat, K, radius, normalized=True):
    """Differentiable `my_ume_generation` / utils/loc_utils.py:157-161: (B,n,C,4)."""
    F = _RawMoments.apply(feat, pts.detach(), kpts.detach(), K, radius)
    if normalized:
        F = F / (F[..., :1].sum(dim=-2, keepdim=True) + 1e-6)
    return F


def _descending(cond):
    """Rows where `cond` (bs,L) holds, per batch entry in DESCENDING order, padded with 0, and how
    many there are (the where / scatter / sort idiom of utils/loc_utils.py:106-111)."""
    ar = torch.arange(cond.shape[1], device=cond.device).expand_as(cond)
    srt = torch.where(cond, ar, torch.full_like(ar, -1)).sort(dim=1, descending=True).values
    return srt.clamp_min(0), (srt > -1).sum(dim=-1)


def generate_ume_from_keypoints2(velo_pts, velo_seg, velo_feat, ref_pts, ref_feat, gt_tform, nn_r=10, max_nn=5000,
                                 min_nn=1000, num_samples=1024, flat_labels=[9], normalized_ume=False,
                                 nn_intersection_r=0.6):
    """utils/loc_utils.py:86-188.  Returns (F_velo, F_ref, velo_keypoint_pts, ref_keypoint_pts,
    matched_nn_intersection_ratio, with_kpts_batch_cond)."""
    R_gt, t_gt = gt_tform[:, :3, :3], gt_tform[:, :3, 3]
    labels = torch.as_tensor(flat_labels, device=velo_seg.device)
    not_flat = (velo_seg != labels).all(dim=-1).flatten(1)                                   # :93
    moved = velo_pts @ R_gt.transpose(-1, -2) + t_gt[:, None]
    in_both = api.neighbor_count(ref_pts, moved, 1, nn_intersection_r) > 0                   # :99-101
    first, lengths = _descending(in_both & not_flat)                                         # :104-111
    cand = torch.gather(velo_pts, 1, first[..., None].expand(-1, -1, 3))
    n1 = int(lengths.min())
    dense = api.neighbor_count(velo_pts, cand[:, :n1].contiguous(), max_nn, nn_r) >= min_nn  # :115-122
    second, lengths2 = _descending(dense)
    has_kpts = lengths2 > 0
    n2 = int(lengths2.min())
    if n2 == 0:                                                                              # :132-145
        second, cand = second[has_kpts], cand[has_kpts]
        velo_pts, velo_feat, ref_feat, ref_pts, gt_tform = (x[has_kpts] for x in (velo_pts, velo_feat, ref_feat, ref_pts, gt_tform))
        R_gt, t_gt = gt_tform[:, :3, :3], gt_tform[:, :3, 3]
        n2 = int(lengths2[has_kpts].min())
    n = min(n2, num_samples)
    velo_kp = torch.gather(cand, 1, second[:, :n, None].expand(-1, -1, 3)).contiguous()     # :147-149
    F_velo = ume_moments_autograd(velo_pts, velo_kp, velo_feat, max_nn, nn_r, normalized=normalized_ume)
    hom = torch.cat([velo_kp, torch.ones_like(velo_kp[..., :1])], dim=-1) @ gt_tform.transpose(-1, -2)   # :164-166
    ref_kp = (hom[..., :3] / hom[..., 3:]).contiguous()
    F_ref = ume_moments_autograd(ref_pts, ref_kp, ref_feat, max_nn, nn_r, normalized=normalized_ume)
    # :179-186 share of a keypoint's (zero-padded) neighbour list that lands within r of the matched list
    bs = velo_kp.shape[0]
    nn_v = api.ball_query(velo_kp, velo_pts, K=max_nn, radius=nn_r, return_nn=True).knn
    nn_r_ = api.ball_query(ref_kp, ref_pts, K=max_nn, radius=nn_r, return_nn=True).knn
    nn_v = (nn_v @ R_gt[:, None].transpose(-1, -2) + t_gt[:, None, None]).flatten(0, 1).contiguous()
    hit = api.neighbor_count(nn_r_.flatten(0, 1).contiguous(), nn_v, 1, nn_intersection_r) > 0
    ratio = hit.view(bs, n, -1).float().mean(dim=-1)
    return F_velo, F_ref, velo_kp, ref_kp, ratio, has_kpts


class _UmeCdist(torch.autograd.Function):
    """D = ume_cdist(ume1, ume2) (utils/loc_utils.py:8-15) with a hand-written backward: the orthonormal bases
    and the Gram-form distance kernel forward, `ume_cdist_backward_f32` (two GEMM-shaped passes over the 4x4
    Gram blocks + a per-matrix triangular solve) backward.  Inputs must have full column rank (the loss
    filters the others out, loss.py:76-88)."""

    @staticmethod
    def forward(ctx, ume1, ume2):
        u1, u2 = ume1.detach().contiguous(), ume2.detach().contiguous()
        Q1, Q2 = api.ume_descriptors(u1), api.ume_descriptors(u2)
        D = api.descriptor_cdist(Q1, Q2, want_D=True)[0]
        ctx.save_for_backward(u1, u2, Q1, Q2, D)
        return D

    @staticmethod
    def backward(ctx, gD):
        u1, u2, Q1, Q2, D = ctx.saved_tensors
        return api.ume_cdist_backward(u1, u2, Q1, Q2, D, gD.detach().contiguous())


class _RigidFromUme(torch.autograd.Function):
    """T = batch_estimate_transform_ume_old(G, H)[0] (utils/loc_utils.py:292-335) with a hand-written backward
    (`ume_rigid_solve_backward_f32`)."""

    @staticmethod
    def forward(ctx, G, H):
        g, h = G.detach().contiguous(), H.detach().contiguous()
        ctx.save_for_backward(g, h)
        return api.rigid_solve(g[None], h[None])[0]

    @staticmethod
    def backward(ctx, gT):
        g, h = ctx.saved_tensors
        return api.rigid_solve_backward(g, h, gT.detach().contiguous())


def ume_cdist_autograd(ume1, ume2):
    """Differentiable utils/loc_utils.py:8-15: (bs,n1,C,4) x (bs,n2,C,4) -> D (bs,n1,n2)."""
    return _UmeCdist.apply(ume1, ume2)


def rigid_from_ume_autograd(G, H):
    """Differentiable (R,t) part of utils/loc_utils.py:292-335: G, H (B',C,4) -> T (B',4,4) with T[:3,:3] = R^T,
    T[:3,3] = b2."""
    return _RigidFromUme.apply(G, H)


class UMEContrastiveLoss(torch.nn.Module):
    """loss.py:49-118, same constructor and forward signature / return tuple."""

    def __init__(self, num_samples=1024, max_nn=5000, min_nn=1000, nn_r=10, tau=0.1, tau_neg=0.1, hd_labels_flag=False,
                 flat_labels=[], nn_intersection_r=0.6, svd_thr=1e-5):
        super().__init__()
        self.n_samples, self.max_nn, self.min_nn, self.nn_r = num_samples, max_nn, min_nn, nn_r
        self.tau, self.tau_neg, self.hd_labels_flag, self.flat_labels = tau, tau_neg, hd_labels_flag, flat_labels
        self.nn_intersection_r, self.svd_thr = nn_intersection_r, svd_thr

    def forward(self, velo_pts, velo_seg, velo_feat, ref_pts, ref_feat, gt_tform):
        velo_ume, ref_ume, kp_v, kp_r, ratio, has_kpts = generate_ume_from_keypoints2(
            velo_pts, velo_seg, velo_feat, ref_pts, ref_feat, gt_tform, num_samples=self.n_samples, max_nn=self.max_nn,
            min_nn=self.min_nn, nn_r=self.nn_r, flat_labels=self.flat_labels, normalized_ume=True,
            nn_intersection_r=self.nn_intersection_r)
        with torch.no_grad():                                                                # :76-80 rank filter
            full = lambda u: (torch.linalg.svdvals(u) > self.svd_thr).sum(dim=-1) == 4
            ok = full(velo_ume) & full(ref_ume)
        drop = torch.zeros(velo_ume.shape[1], dtype=torch.bool, device=velo_ume.device)
        drop[torch.where(~ok)[1]] = True                                                     # :87-88: any batch entry
        velo_ume, ref_ume, ratio = velo_ume[:, ~drop], ref_ume[:, ~drop], ratio[:, ~drop]
        D = ume_cdist_autograd(velo_ume, ref_ume)
        root = np.sqrt(velo_ume.shape[-1])
        sim = (root - 2 * D) / root                                                          # :95
        eye = torch.eye(D.shape[-1], dtype=torch.bool, device=D.device)[None].expand_as(D)
        tau = torch.where(eye, torch.full_like(sim, self.tau), torch.full_like(sim, self.tau_neg))
        e = torch.exp(sim / tau)
        loss = -torch.log(torch.diagonal(e / e.sum(dim=-1, keepdim=True), dim1=-1, dim2=-2)).mean()
        return loss, kp_v, kp_r, velo_ume, ref_ume, ratio, has_kpts


class CubeRegistrationLoss(torch.nn.Module):
    """loss.py:121-190, same constructor and forward signature / return tuple."""

    def __init__(self, rtume_max_nn, rtume_r_nn, cube_scale=1.0, nn_inter_ratio_thr=0.75):
        super().__init__()
        corners = torch.tensor([[sx, sy, sz] for sz in (1, -1) for sy in (1, -1) for sx in (-1, 1)], dtype=torch.float32)
        self.points_cube = corners * cube_scale                                             # loss.py:126-134 order
        self.nn_inter_ratio_thr = nn_inter_ratio_thr

    def forward(self, src_pts, src_ume, tgt_pts, tgt_ume, gt_tform, matched_nn_intersection_ratio, valid_batch_entries):
        gt = gt_tform[valid_batch_entries]
        bs, nh = src_ume.shape[:2]
        cube = self.points_cube.to(src_ume.device)
        T = rigid_from_ume_autograd(src_ume.reshape(-1, *src_ume.shape[2:]), tgt_ume.reshape(-1, *tgt_ume.shape[2:]))
        T = T.view(bs, nh, 4, 4)
        R, t = T[..., :3, :3], T[..., :3, 3]
        est = cube @ R.transpose(-1, -2) + t.unsqueeze(-2)                                   # (bs,nh,8,3)
        want = (cube @ gt[:, :3, :3].transpose(-1, -2) + gt[:, None, :3, 3])[:, None]
        per_hyp = (want - est).norm(dim=-1).mean(dim=-1)
        use = matched_nn_intersection_ratio >= self.nn_inter_ratio_thr
        if use.sum() == 0:
            use = matched_nn_intersection_ratio >= matched_nn_intersection_ratio.median(dim=-1, keepdim=True)[0]
        loss = per_hyp[use].mean()
        with torch.no_grad():
            rre = api.relative_rotation_error(R.reshape(-1, 3, 3), gt[:, None, :3, :3].expand(-1, nh, -1, -1).reshape(-1, 3, 3)).view(bs, nh)
            rte = (t - gt[:, None, :3, 3]).norm(dim=-1)
        return loss, rre, rte
