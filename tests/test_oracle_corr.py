"""CPU: pins the oracle's hypothesis-selection restatement against the golden vectors
produced by the reference's own FeatureCorrelator / feature_spatial_var / pc_corr_cost_pytorch3d."""
import numpy as np

from oracle import ume_oracle as orc
from oracle import pytorch3d_ops as p3d


def test_correlator_golden(golden):
    g = golden("correlator")
    sv = orc.feature_spatial_var(g["src_pts"][None], g["src_feat"][None], knn=50)
    tv = orc.feature_spatial_var(g["tgt_pts"][None], g["tgt_feat"][None], knn=50)
    assert np.abs(sv[0] - g["src_var"]).max() < 2e-6 and np.abs(tv[0] - g["tgt_var"]).max() < 2e-6
    best, scores = orc.feature_corr_hypothesis_test(g["src_pts"][None], g["tgt_pts"][None], g["src_feat"][None],
                                                    g["tgt_feat"][None], g["T_kp"], sigma=float(g["sigma"]),
                                                    corr_num_nn=int(g["corr_num_nn"]))
    assert np.abs(scores - g["scores"]).max() < 1e-5 * np.abs(g["scores"]).max()
    assert np.array_equal(best, g["best_T"])
    assert np.array_equal(best, g["T_kp"][0])                       # the ground truth wins
    s64 = orc.feature_corr_hypothesis_test(g["src_pts"][None], g["tgt_pts"][None], g["src_feat"][None],
                                           g["tgt_feat"][None], g["T_kp"], sigma=1.5, corr_num_nn=20, dtype=np.float64)[1]
    assert np.abs(s64 - g["scores"]).max() < 1e-4 * np.abs(g["scores"]).max()


def test_knn_general_k_np_vs_c():
    rng = np.random.default_rng(8)
    p = rng.uniform(-5, 5, size=(2, 400, 3)).astype(np.float32)
    q = rng.uniform(-6, 6, size=(2, 60, 3)).astype(np.float32)
    a, b = p3d.knn_points_np(q, p, 20), p3d.knn_points_c(q, p, 20)
    assert np.array_equal(a.idx, b.idx) and np.array_equal(a.dists, b.dists)
    assert (np.diff(b.dists, axis=-1) >= 0).all()
