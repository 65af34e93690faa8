"""Generate tests/golden/*.npz by running the UNMODIFIED reference functions (imported from
/root/reference under sys.modules stubs, see ref_import.py) on seeded synthetic inputs.

Run in the build container only:   python tests/golden/make_golden.py
The GPU box has no /root/reference; it only ever reads the committed .npz files.

What is reference output and what is not:
  * F_src/F_tgt (evaluate.my_ume_generation), D (utils.loc_utils.ume_cdist), match (the arg-min
    lines of evaluate.py:224-225), T/Dpair (batch_estimate_transform_ume_old), rre
    (utils.eval_utils.relative_rotation_error), kp-layer outputs (ume_kp_layer.forward) are
    produced by the reference's own code, torch 2.11 CPU fp32.
  * the neighbour indices underneath come from the oracle's restatement of pytorch3d.ball_query
    (pytorch3d is not installable offline) — they are saved as `bq_idx_*` and flagged as such.
"""
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, HERE)

from oracle.ref_import import import_reference  # noqa: E402
from umeregrobust_b200 import synth  # noqa: E402
from oracle import pytorch3d_ops as p3d  # noqa: E402


def t(x):
    return torch.from_numpy(np.ascontiguousarray(x))


def small_pair(seed, N, C, n_kp, exact_copy=False, spread=12.0):
    rng = np.random.default_rng(seed)
    src = np.stack([rng.uniform(-spread, spread, N), rng.uniform(-spread, spread, N),
                    rng.uniform(-1.5, 1.5, N)], 1).astype(np.float32)
    src += np.array([30.0, -20.0, 0.0], dtype=np.float32)           # away from the origin, as in a scan
    gt = synth.random_rigid(rng, t_range=(2.0, 6.0))
    feat = synth._normalize_rows(rng.normal(size=(N, C))).astype(np.float32)
    if exact_copy:
        tgt = (src.astype(np.float64) @ gt[:3, :3].T + gt[:3, 3]).astype(np.float32)
        tfeat = feat.copy()
        ks = rng.choice(N, n_kp, replace=False)
        kt = ks.copy()
    else:
        perm = rng.permutation(N)
        tgt = ((src[perm] + rng.normal(scale=0.02, size=(N, 3))).astype(np.float64) @ gt[:3, :3].T
               + gt[:3, 3]).astype(np.float32)
        tfeat = synth._normalize_rows(feat[perm] + rng.normal(scale=0.05, size=(N, C))).astype(np.float32)
        ks = rng.choice(N, n_kp, replace=False)
        inv = np.argsort(perm)
        kt = inv[ks]                                                  # same physical keypoints
        kt = kt[rng.permutation(n_kp)]
    return dict(src_pts=src, src_feat=feat, src_kp=src[ks].copy(), tgt_pts=tgt, tgt_feat=tfeat,
                tgt_kp=tgt[kt].copy(), gt=gt.astype(np.float32))


def golden_hot_path(ref_eval_mod, ref_loc, ref_evalutils, name, seed, N, n_kp, K, radius, exact_copy):
    """evaluate.py:206-257 on one pair, C = 32 (evaluate.py hard-codes 32)."""
    p = small_pair(seed, N, 32, n_kp, exact_copy=exact_copy)
    args = SimpleNamespace(ume_max_nn=K, ume_r_nn=radius)
    with torch.no_grad():
        F_src = ref_eval_mod.my_ume_generation(t(p["src_pts"])[None], t(p["src_kp"])[None],
                                               t(p["src_feat"])[None], args)
        F_tgt = ref_eval_mod.my_ume_generation(t(p["tgt_pts"])[None], t(p["tgt_kp"])[None],
                                               t(p["tgt_feat"])[None], args)
        D = ref_loc.ume_cdist(F_src, F_tgt)
        m = D.min(dim=-1)[1]                                            # evaluate.py:224
        m = torch.cat([torch.arange(D.shape[1])[None, :, None], m[..., None]], dim=-1)
        G = torch.gather(F_src, 1, m[..., 0][..., None, None].expand(-1, -1, 32, 4))   # :231
        H = torch.gather(F_tgt, 1, m[..., 1][..., None, None].expand(-1, -1, 32, 4))   # :230
        T, Dpair = ref_loc.batch_estimate_transform_ume_old(G.reshape(-1, 32, 4), H.reshape(-1, 32, 4))
        R_gt = t(p["gt"])[None, :3, :3].expand(T.shape[0], -1, -1)
        rre = ref_evalutils.relative_rotation_error(T[:, :3, :3], R_gt)
    bq_s = p3d.ball_query_c(p["src_kp"][None], p["src_pts"][None], K, radius, return_nn=False)
    bq_t = p3d.ball_query_c(p["tgt_kp"][None], p["tgt_pts"][None], K, radius, return_nn=False)
    out = dict(p, K=np.int64(K), radius=np.float32(radius), F_src=F_src.numpy(), F_tgt=F_tgt.numpy(),
               D=D.numpy(), match=m.numpy(), T=T.numpy(), Dpair=Dpair.numpy(), rre=rre.numpy(),
               bq_idx_src=bq_s.idx.astype(np.int32), bq_idx_tgt=bq_t.idx.astype(np.int32))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    return out


def golden_config1(ref_loc, ref_evalutils):
    """BASELINE config #1: 2 clouds x 1024 pts, random rigid, 8-dim feats, whole-cloud UME via
    create_local_ume_matrix (utils/loc_utils.py:434-445) + the rigid solve."""
    p = small_pair(101, 1024, 8, 4, exact_copy=True)
    with torch.no_grad():
        G = ref_loc.create_local_ume_matrix(t(p["src_pts"])[None, None], t(p["src_feat"])[None, None])
        H = ref_loc.create_local_ume_matrix(t(p["tgt_pts"])[None, None], t(p["tgt_feat"])[None, None])
        T, Dpair = ref_loc.batch_estimate_transform_ume_old(G.reshape(-1, 8, 4), H.reshape(-1, 8, 4))
        rre = ref_evalutils.relative_rotation_error(T[:, :3, :3], t(p["gt"])[None, :3, :3])
    out = dict(src_pts=p["src_pts"], src_feat=p["src_feat"], tgt_pts=p["tgt_pts"], tgt_feat=p["tgt_feat"],
               gt=p["gt"], G=G.numpy()[0], H=H.numpy()[0], T=T.numpy(), Dpair=Dpair.numpy(), rre=rre.numpy())
    np.savez_compressed(os.path.join(HERE, "config1_whole_cloud.npz"), **out)
    return out


def golden_kp_layer(ref_loc):
    """ume_kp_layer.forward (utils/loc_utils.py:380-431), diag_only True and False, C = 16."""
    p = small_pair(202, 2048, 16, 24, exact_copy=False)
    res = {}
    for diag in (True, False):
        layer = ref_loc.ume_kp_layer(ume_knn=48, ume_desc_rad=3.0, diag_only=diag)
        with torch.no_grad():
            T, D, G_kp, H_kp = layer(t(p["src_pts"])[None], t(p["src_feat"])[None], t(p["src_kp"])[None],
                                     t(p["tgt_pts"])[None], t(p["tgt_feat"])[None], t(p["tgt_kp"])[None])
        tag = "diag" if diag else "full"
        res.update({"T_" + tag: T.numpy(), "D_" + tag: D.numpy(), "G_" + tag: G_kp.numpy(),
                    "H_" + tag: H_kp.numpy()})
    out = dict(p, ume_knn=np.int64(48), ume_desc_rad=np.float32(3.0), **res)
    np.savez_compressed(os.path.join(HERE, "kp_layer.npz"), **out)
    return out


def golden_kp_layer_nrand(ref_loc):
    """ume_kp_layer.forward with n_rand (utils/loc_utils.py:406-410: sums of random triplets of the
    diagonal pairs; host RNG np.random.choice, seeded here so that the draw is reproducible)."""
    p = small_pair(203, 2048, 16, 24, exact_copy=False)
    layer = ref_loc.ume_kp_layer(ume_knn=48, ume_desc_rad=3.0, diag_only=True, n_rand=40)
    np.random.seed(7)
    with torch.no_grad():
        T, D, G_kp, H_kp = layer(t(p["src_pts"])[None], t(p["src_feat"])[None], t(p["src_kp"])[None],
                                 t(p["tgt_pts"])[None], t(p["tgt_feat"])[None], t(p["tgt_kp"])[None])
    out = dict(p, ume_knn=np.int64(48), ume_desc_rad=np.float32(3.0), n_rand=np.int64(40), np_seed=np.int64(7),
               T=T.numpy(), D=D.numpy(), G=G_kp.numpy(), H=H_kp.numpy())
    np.savez_compressed(os.path.join(HERE, "kp_layer_nrand.npz"), **out)
    return out


def golden_rigid_random(ref_loc):
    """batch_estimate_transform_ume_old on random well-conditioned (G,H) pairs for C in {8,32,64}:
    H built from G by a known rigid 'D' matrix so the answer is also known analytically."""
    out = {}
    rng = np.random.default_rng(303)
    for C in (8, 32, 64):
        nb = 64
        w = rng.uniform(0.2, 1.0, size=(nb, C, 1)) * np.sign(rng.normal(size=(nb, C, 1)))
        x = rng.normal(scale=3.0, size=(nb, C, 3)) + rng.uniform(-40, 40, size=(nb, 1, 3))
        G = np.concatenate([w, w * x], -1)
        Ts = np.stack([synth.random_rigid(rng, max_tilt_deg=180.0) for _ in range(nb)])
        y = x @ np.swapaxes(Ts[:, :3, :3], 1, 2) + Ts[:, None, :3, 3]
        H = np.concatenate([w, w * y], -1) + rng.normal(scale=1e-3, size=(nb, C, 4))
        with torch.no_grad():
            T, D = ref_loc.batch_estimate_transform_ume_old(t(G.astype(np.float32)), t(H.astype(np.float32)))
        out.update({f"G{C}": G.astype(np.float32), f"H{C}": H.astype(np.float32), f"T{C}": T.numpy(),
                    f"D{C}": D.numpy(), f"Tgt{C}": Ts.astype(np.float32)})
    np.savez_compressed(os.path.join(HERE, "rigid_random.npz"), **out)
    return out


def golden_correlator(ref_loc):
    """FeatureCorrelator.feature_corr_hypothesis_test (utils/loc_utils.py:634-681) plus its parts
    (feature_spatial_var :579-585, pc_corr_cost_pytorch3d :621-631) on a small pair: 24 hypotheses =
    the ground truth, perturbed versions of it and random rigid motions."""
    rng = np.random.default_rng(404)
    Ns, Nt, C = 1500, 1400, 32
    src = np.stack([rng.uniform(-15, 15, Ns), rng.uniform(-15, 15, Ns), rng.uniform(-1, 1, Ns)], 1).astype(np.float32)
    gt = synth.random_rigid(rng, t_range=(2.0, 6.0))
    sub = rng.choice(Ns, Nt, replace=False)
    tgt = ((src[sub] + rng.normal(scale=0.03, size=(Nt, 3))).astype(np.float64) @ gt[:3, :3].T + gt[:3, 3]).astype(np.float32)
    # smooth feature field + noise, so that spatial variance is informative
    W = rng.normal(size=(3, C)) * 0.15
    sfeat = synth._normalize_rows(np.sin(src @ W) + 0.2 * rng.normal(size=(Ns, C))).astype(np.float32)
    tfeat = synth._normalize_rows(sfeat[sub] + 0.05 * rng.normal(size=(Nt, C))).astype(np.float32)
    hyps = [gt]
    for i in range(11):
        d = synth.random_rigid(rng, t_range=(0.0, 0.5 * (i + 1)), max_tilt_deg=2.0, yaw_deg=rng.uniform(-3, 3) * (i + 1))
        hyps.append(d @ gt)
    for i in range(12):
        hyps.append(synth.random_rigid(rng, t_range=(0.0, 10.0)))
    T_kp = np.stack(hyps).astype(np.float32)
    corr = ref_loc.FeatureCorrelator(sigma=1.5, batch=8, n_hypotheses=10)
    with torch.no_grad():
        best_T = corr.feature_corr_hypothesis_test(t(src)[None], t(tgt)[None], t(sfeat)[None], t(tfeat)[None], t(T_kp))
        sw = ref_loc.feature_spatial_var(t(src)[None], t(sfeat)[None], knn=50)
        tw = ref_loc.feature_spatial_var(t(tgt)[None], t(tfeat)[None], knn=50)
        m = torch.mean(torch.concat((t(sfeat)[None], t(tfeat)[None]), dim=1), dim=1)
        wsf = (t(sfeat)[None] - m) * sw.unsqueeze(-1)
        wtf = (t(tfeat)[None] - m) * tw.unsqueeze(-1)
        scores = ref_loc.pc_corr_cost_pytorch3d(t(T_kp)[:, :3, :3], t(T_kp)[:, :3, 3], t(src), t(tgt), 20, wsf[0], wtf[0], 1.5,
                                                None, use_norm=False, src_norm=None, tgt_norm=None, dev="cpu")
    out = dict(src_pts=src, tgt_pts=tgt, src_feat=sfeat, tgt_feat=tfeat, T_kp=T_kp, gt=gt.astype(np.float32),
               best_T=best_T.numpy(), scores=scores.numpy(), src_var=sw.numpy()[0], tgt_var=tw.numpy()[0],
               sigma=np.float32(1.5), corr_num_nn=np.int64(20))
    np.savez_compressed(os.path.join(HERE, "correlator.npz"), **out)
    return out


def main():
    torch.manual_seed(0)
    np.random.seed(0)
    torch.set_num_threads(1)                                      # reproducible reduction order
    ref_evaluate, ref_loc, ref_evalutils = import_reference()
    a = golden_hot_path(ref_evaluate, ref_loc, ref_evalutils, "hotpath_noisy", 11, 4096, 96, 256, 3.3, False)
    b = golden_hot_path(ref_evaluate, ref_loc, ref_evalutils, "hotpath_exact", 12, 3000, 64, 40, 2.5, True)
    c = golden_config1(ref_loc, ref_evalutils)
    golden_kp_layer(ref_loc)
    golden_rigid_random(ref_loc)
    print("hotpath_noisy: median RRE deg", float(np.median(a["rre"])))
    print("hotpath_exact: median RRE deg", float(np.median(b["rre"])),
          "diag match frac", float((b["match"][0, :, 1] == np.arange(b["match"].shape[1])).mean()))
    print("config1: RRE deg", c["rre"], "t err", np.abs(c["T"][0, :3, 3] - c["gt"][:3, 3]).max())
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "kp_layer_nrand":      # add one fixture without touching the others
        torch.manual_seed(0)
        torch.set_num_threads(1)
        golden_kp_layer_nrand(import_reference()[1])
    else:
        main()
