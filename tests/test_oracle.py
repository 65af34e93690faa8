"""CPU: pins the oracle (oracle/) against the committed golden vectors — outputs of the
UNMODIFIED reference functions, see tests/golden/make_golden.py — and against analytic
known-answer tests that follow from the reference's math."""
import numpy as np
import pytest

from oracle import pytorch3d_ops as p3d
from oracle import ume_oracle as orc
from umeregrobust_b200 import synth


def rel_err(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def moment_err(F, F_ref, kappa):
    """max over keypoints of (relative error of the C x 4 matrix) / kappa — see
    oracle.ume_oracle.normaliser_condition."""
    num = np.abs(F - F_ref).max(axis=(-1, -2))
    den = np.abs(F_ref).max(axis=(-1, -2))
    return float((num / den / kappa).max())


# ------------------------------------------------------------------ pytorch3d restatements
def test_ball_query_np_and_c_agree():
    rng = np.random.default_rng(0)
    p2 = rng.uniform(-4, 4, size=(2, 700, 3)).astype(np.float32)
    p1 = p2[:, rng.choice(700, 40, replace=False)]
    a = p3d.ball_query_np(p1, p2, 24, 1.5)
    b = p3d.ball_query_c(p1, p2, 24, 1.5)
    assert np.array_equal(a.idx, b.idx)
    assert np.array_equal(a.dists, b.dists)
    assert np.array_equal(a.knn, b.knn)


def test_ball_query_semantics():
    # first K in ROW order, strict '<', -1 / 0 padding, zero-padded nn
    p2 = np.array([[[0, 0, 0], [3, 0, 0], [1, 0, 0], [0.5, 0, 0], [0, 2, 0], [0, 0, 1.9999]]], np.float32)
    p1 = np.array([[[0, 0, 0], [10, 10, 10]]], np.float32)
    r = p3d.ball_query_c(p1, p2, 3, 2.0)
    assert r.idx[0, 0].tolist() == [0, 2, 3]          # row order, not nearest-first; idx 4 is at d = r (excluded)
    assert r.idx[0, 1].tolist() == [-1, -1, -1]
    assert r.dists[0, 1].tolist() == [0, 0, 0]
    assert np.all(r.knn[0, 1] == 0)
    r = p3d.ball_query_c(p1, p2, 8, 2.0)
    assert r.idx[0, 0].tolist() == [0, 2, 3, 5, -1, -1, -1, -1]   # d == r is NOT inside (strict)


def test_knn_np_and_c_agree_and_ties():
    rng = np.random.default_rng(1)
    p2 = rng.integers(-3, 3, size=(1, 300, 3)).astype(np.float32)     # many exact ties
    p1 = rng.integers(-3, 3, size=(1, 50, 3)).astype(np.float32)
    a = p3d.knn_points_np(p1, p2, 4)
    b = p3d.knn_points_c(p1, p2, 4)
    assert np.array_equal(a.idx, b.idx)
    assert np.array_equal(a.dists, b.dists)
    x = rng.normal(size=(1, 300, 5)).astype(np.float32)
    g = p3d.knn_gather_np(x, a.idx)
    assert g.shape == (1, 50, 4, 5) and np.array_equal(g[0, 7, 2], x[0, a.idx[0, 7, 2]])


def test_ball_query_fma_flips_are_rare():
    rng = np.random.default_rng(2)
    p2 = rng.uniform(-30, 30, size=(1, 20000, 3)).astype(np.float32)
    p1 = p2[:, :64]
    a = p3d.ball_query_c(p1, p2, 4000, 5.0, fma=False)
    b = p3d.ball_query_c(p1, p2, 4000, 5.0, fma=True)
    same = (a.idx == b.idx).all(-1).mean()
    assert same > 0.9


# ------------------------------------------------------------------ golden vectors (reference outputs)
@pytest.mark.parametrize("name", ["hotpath_noisy", "hotpath_exact"])
def test_hot_path_matches_reference_golden(golden, name):
    g = golden(name)
    K, radius = int(g["K"]), float(g["radius"])
    F_src, idx = orc.ume_moments(g["src_pts"][None], g["src_kp"][None], g["src_feat"][None], K, radius,
                                 return_idx=True)
    assert np.array_equal(idx[0], g["bq_idx_src"][0].astype(np.int64))
    F_tgt, idx_t = orc.ume_moments(g["tgt_pts"][None], g["tgt_kp"][None], g["tgt_feat"][None], K, radius,
                                   return_idx=True)
    # the reference's fp32 result vs the fp32 and fp64 restatements, tolerance scaled by the
    # conditioning of the normaliser sum (signed features cancel)
    ks = orc.normaliser_condition(g["src_feat"][None], idx)
    kt = orc.normaliser_condition(g["tgt_feat"][None], idx_t)
    assert moment_err(F_src, g["F_src"], ks) < 2e-6
    assert moment_err(F_tgt, g["F_tgt"], kt) < 2e-6
    F64 = orc.ume_moments(g["src_pts"][None], g["src_kp"][None], g["src_feat"][None], K, radius, dtype=np.float64)
    assert moment_err(g["F_src"], F64, ks) < 2e-6
    # distances / arg-min / rigid solve on the REFERENCE's F so that stages are pinned one by one
    D = orc.ume_cdist(g["F_src"], g["F_tgt"])
    assert np.abs(D - g["D"]).max() < 2e-3            # fp32 mm-form cdist noise near 0
    D64 = orc.ume_cdist(g["F_src"], g["F_tgt"], dtype=np.float64)
    far = g["D"] > 0.05
    assert np.abs(D64 - g["D"])[far].max() < 5e-5
    m64 = orc.match_argmin(D64)
    gap = np.sort(D64, -1)
    clear = (gap[..., 1] - gap[..., 0]) > 1e-3
    assert np.array_equal(m64[clear], g["match"][clear])
    G = g["F_src"][0][g["match"][0, :, 0]]
    H = g["F_tgt"][0][g["match"][0, :, 1]]
    T, Dp = orc.rigid_from_ume(G, H)
    T64, Dp64 = orc.rigid_from_ume(G, H, dtype=np.float64)
    assert orc.rotation_angle_rad(T64[:, :3, :3], g["T"][:, :3, :3]).max() < 2e-3
    assert np.median(orc.rotation_angle_rad(T64[:, :3, :3], g["T"][:, :3, :3])) < 1e-4
    assert np.median(np.abs(T64[:, :3, 3] - g["T"][:, :3, 3]).max(-1)) < 5e-3
    assert np.abs(Dp64 - g["Dpair"]).max() < 2e-3
    rre = orc.relative_rotation_error(np.broadcast_to(g["gt"][:3, :3], (len(T), 3, 3)), g["T"][:, :3, :3])
    assert np.abs(rre - g["rre"]).max() < 5e-2          # fp32 acos near 1 quantises to ~0.04 deg


def test_config1_whole_cloud_golden(golden):
    g = golden("config1_whole_cloud")
    G = orc.moments_from_neighbors(g["src_pts"][None], g["src_feat"][None], normalise=False)
    H = orc.moments_from_neighbors(g["tgt_pts"][None], g["tgt_feat"][None], normalise=False)
    assert rel_err(G, g["G"]) < 1e-5 and rel_err(H, g["H"]) < 1e-5
    T, _ = orc.rigid_from_ume(g["G"], g["H"], dtype=np.float64)
    assert orc.rotation_angle_rad(T[:, :3, :3], g["T"][:, :3, :3]).max() < 1e-4
    assert np.abs(T[:, :3, 3] - g["T"][:, :3, 3]).max() < 1e-3
    # exact recovery of the ground truth (RRE 0.0, RTE ~1e-7..1e-5 m)
    assert orc.rotation_angle_rad(T[0, :3, :3], g["gt"][:3, :3]) < 1e-4
    assert np.abs(T[0, :3, 3] - g["gt"][:3, 3]).max() < 1e-3


def test_rigid_random_golden(golden):
    g = golden("rigid_random")
    for C in (8, 32, 64):
        T64, D64 = orc.rigid_from_ume(g[f"G{C}"], g[f"H{C}"], dtype=np.float64)
        assert orc.rotation_angle_rad(T64[:, :3, :3], g[f"T{C}"][:, :3, :3]).max() < 1e-4
        assert np.abs(T64[:, :3, 3] - g[f"T{C}"][:, :3, 3]).max() < 2e-3
        assert np.abs(D64 - g[f"D{C}"]).max() < 2e-3
        assert orc.rotation_angle_rad(T64[:, :3, :3], g[f"Tgt{C}"][:, :3, :3]).max() < 5e-2   # H carries 1e-3 noise


def test_kp_layer_golden(golden):
    g = golden("kp_layer")
    for tag, diag in (("diag", True), ("full", False)):
        T, D, G, H = orc.ume_kp_layer_forward(g["src_pts"][None], g["src_feat"][None], g["src_kp"][None],
                                              g["tgt_pts"][None], g["tgt_feat"][None], g["tgt_kp"][None],
                                              int(g["ume_knn"]), float(g["ume_desc_rad"]), diag_only=diag,
                                              dtype=np.float64)
        assert rel_err(G[0], g["G_" + tag]) < 2e-5
        assert rel_err(H[0], g["H_" + tag]) < 2e-5
        assert T.shape == g["T_" + tag].shape and D.shape == g["D_" + tag].shape
        assert np.abs(D - g["D_" + tag]).max() < 2e-3
        ang = orc.rotation_angle_rad(T[..., :3, :3], g["T_" + tag][..., :3, :3])
        assert np.median(ang) < 1e-3


# ------------------------------------------------------------------ analytic known answers
def test_cdist_invariances():
    rng = np.random.default_rng(5)
    F1 = rng.normal(size=(1, 20, 32, 4))
    F2 = rng.normal(size=(1, 30, 32, 4))
    D = orc.ume_cdist(F1, F2, dtype=np.float64)
    assert D.min() >= 0 and D.max() <= 2.0 + 1e-9
    assert np.abs(np.diag(orc.ume_cdist(F1, F1, dtype=np.float64)[0])).max() < 1e-6
    A = rng.normal(size=(4, 4)) + 3 * np.eye(4)                      # any invertible 4x4: same column space
    assert np.abs(orc.ume_cdist(F1 @ A, F2, dtype=np.float64) - D).max() < 1e-9
    assert np.abs(orc.ume_cdist_gram(F1, F2) - D).max() < 1e-7       # D^2 = 4 - |Q1^T Q2|_F^2


def test_exact_recovery_from_local_ume():
    p = synth.make_pair(3, N=6000, C=8, n_kp=32, exact_copy=True, generator="disc")
    out = orc.register_pair_hypotheses(p["src_pts"][None], p["src_feat"][None], p["src_kp"][None],
                                       p["tgt_pts"][None], p["tgt_feat"][None], p["tgt_kp"][None],
                                       K=6000, radius=6.0, dtype=np.float64)
    # identical neighbourhoods + identical features: every keypoint matches itself, every
    # hypothesis is the ground truth.  (fp32 inputs: the moved cloud is rounded to fp32.)
    assert np.array_equal(out["match"][0, :, 1], np.arange(32))
    ang = orc.rotation_angle_rad(out["T"][0, :, :3, :3], p["gt"][:3, :3].astype(np.float64))
    assert ang.max() < 1e-4
    assert np.abs(out["T"][0, :, :3, 3] - p["gt"][:3, 3]).max() < 5e-3


def test_sparse_quantize_known_answer():
    # rows 0 and 2 share voxel (0,0,0), rows 1 and 4 share (-1,0,3); first occurrences survive in row order
    c = np.array([[0.1, 0.2, 0.29], [-0.1, 0.0, 0.95], [0.29, 0.0, 0.0], [0.3, 0.0, 0.0], [-0.29, 0.1, 1.1]], np.float32)
    vox, idx = orc.sparse_quantize(c, 0.3)
    assert idx.tolist() == [0, 1, 3]
    assert vox.tolist() == [[0, 0, 0], [-1, 0, 3], [1, 0, 0]]
    rng = np.random.default_rng(0)
    pts = rng.uniform(-50, 50, (20000, 3)).astype(np.float32)
    vox, idx = orc.sparse_quantize(pts, 1.0)
    assert np.all(np.diff(idx) > 0) and len(np.unique(vox, axis=0)) == len(vox)
    # idempotent: the survivors are all distinct voxels
    assert len(orc.sparse_quantize(pts[idx], 1.0)[1]) == len(idx)


def test_moments_backward_against_reference_autograd(golden):
    # the reference's autograd gradient of sum(F_velo * w1) + sum(F_ref * w2) through its materialised
    # gather (tests/golden/make_golden_training.py) vs the oracle's closed form
    g = golden("training")
    K, r = int(g["kw_max_nn"]), float(g["kw_nn_r"])
    for side, seed in (("velo", 1), ("ref", 2)):
        F = g["raw_F_" + side]
        w = np.random.default_rng(seed).normal(size=F.shape).astype(np.float32)
        pts = g["velo_pts"] if side == "velo" else g["ref_pts"]
        feat = g["velo_feat"] if side == "velo" else g["ref_feat"]
        got = orc.ume_moments_backward(pts, g["raw_kp_" + side], w, K, r)
        ref = g["raw_grad_%s_feat" % side]
        assert np.abs(got - ref).max() < 1e-4 * np.abs(ref).max()
        # the reference's raw (un-normalised) moments from the oracle's neighbour sets
        idx = orc.ume_moments(pts, g["raw_kp_" + side], feat, K, r, return_idx=True)[1]
        fp = np.concatenate([feat, np.zeros_like(feat[:, :1])], 1).astype(np.float64)
        pp = np.concatenate([pts, np.zeros_like(pts[:, :1])], 1).astype(np.float64)
        safe = np.where(idx < 0, pts.shape[1], idx)
        for b in range(F.shape[0]):
            nf, npt = fp[b][safe[b]], pp[b][safe[b]]
            Fr = np.concatenate([nf.sum(1)[..., None], np.einsum("nkc,nkd->ncd", nf, npt)], -1)
            assert np.abs(Fr - F[b]).max() < 2e-5 * np.abs(F[b]).max()
