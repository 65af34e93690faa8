"""Golden vectors of the training-time UME generation and its losses, produced by the
UNMODIFIED reference (`utils.loc_utils.generate_ume_from_keypoints2`, `loss.UMEContrastiveLoss`,
`loss.CubeRegistrationLoss`) on CPU torch under the stubs of ref_import.py, including the gradient
of the summed losses with respect to the per-point features (the reference gets it from autograd
through its materialised (bs, n, max_nn, C) gather).

Run in the build container only:   python tests/golden/make_golden_training.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, HERE)

from oracle.ref_import import import_reference, REF_ROOT  # noqa: E402
from umeregrobust_b200 import synth  # noqa: E402


def scene(seed, bs=2, N=2000, Nr=1700, C=32):
    rng = np.random.default_rng(seed)
    velo, ref, vf, rf, seg, gts = [], [], [], [], [], []
    for _ in range(bs):
        p = np.stack([rng.uniform(-9, 9, N), rng.uniform(-9, 9, N), rng.uniform(-1, 1, N)], 1).astype(np.float32)
        p += np.array([12.0, -7.0, 0.5], np.float32)
        gt = synth.random_rigid(rng, t_range=(1.0, 3.0)).astype(np.float32)
        f = synth._normalize_rows(rng.normal(size=(N, C))).astype(np.float32)
        sel = rng.permutation(N)[:Nr]
        q = ((p[sel] + rng.normal(scale=0.05, size=(Nr, 3))) @ gt[:3, :3].T.astype(np.float64) + gt[:3, 3]).astype(np.float32)
        g = synth._normalize_rows(f[sel] + rng.normal(scale=0.05, size=(Nr, C))).astype(np.float32)
        s = rng.integers(0, 12, size=(N, 1)).astype(np.int64)          # label 9 = "flat"
        velo.append(p); ref.append(q); vf.append(f); rf.append(g); seg.append(s); gts.append(gt)
    return (np.stack(velo), np.stack(seg), np.stack(vf), np.stack(ref), np.stack(rf), np.stack(gts))


def main():
    torch.manual_seed(0)
    np.random.seed(0)
    torch.set_num_threads(1)
    _, ref_loc, _ = import_reference()
    old = os.getcwd()
    os.chdir(REF_ROOT)
    sys.path.insert(0, REF_ROOT)
    import loss as ref_loss                                            # the reference's loss.py
    os.chdir(old)
    kw = dict(nn_r=3.0, max_nn=200, min_nn=40, num_samples=24, flat_labels=[9], nn_intersection_r=0.6)
    velo, seg, vf, ref, rf, gt = scene(5)
    tv = [torch.from_numpy(x) for x in (velo, seg, vf, ref, rf, gt)]
    vft = tv[2].clone().requires_grad_(True)
    rft = tv[4].clone().requires_grad_(True)
    out = {"velo_pts": velo, "velo_seg": seg, "velo_feat": vf, "ref_pts": ref, "ref_feat": rf, "gt_tform": gt}
    out.update({"kw_" + k: np.asarray(v) for k, v in kw.items()})
    for norm in (False, True):
        F_v, F_r, kp_v, kp_r, ratio, cond = ref_loc.generate_ume_from_keypoints2(
            tv[0], tv[1], vft, tv[3], rft, tv[5], normalized_ume=norm, **kw)
        tag = "norm_" if norm else "raw_"
        out.update({tag + "F_velo": F_v.detach().numpy(), tag + "F_ref": F_r.detach().numpy(),
                    tag + "kp_velo": kp_v.numpy(), tag + "kp_ref": kp_r.numpy(), tag + "ratio": ratio.numpy(),
                    tag + "cond": cond.numpy()})
        w1 = torch.from_numpy(np.random.default_rng(1).normal(size=tuple(F_v.shape)).astype(np.float32))
        w2 = torch.from_numpy(np.random.default_rng(2).normal(size=tuple(F_r.shape)).astype(np.float32))
        L = (F_v * w1).sum() + (F_r * w2).sum()
        gv, gr = torch.autograd.grad(L, [vft, rft])
        if not norm:                                                   # (seeds of w1 / w2: 1 and 2)
            out.update({tag + "grad_velo_feat": gv.numpy(), tag + "grad_ref_feat": gr.numpy()})
    # the two UME losses of train_coloring.py:48-58 and their gradient with respect to the features
    ume_loss_fn = ref_loss.UMEContrastiveLoss(num_samples=kw["num_samples"], max_nn=kw["max_nn"], min_nn=kw["min_nn"],
                                              nn_r=kw["nn_r"], tau=0.1, tau_neg=0.1, flat_labels=[9],
                                              nn_intersection_r=kw["nn_intersection_r"])
    reg_loss_fn = ref_loss.CubeRegistrationLoss(rtume_max_nn=kw["max_nn"], rtume_r_nn=kw["nn_r"], cube_scale=1.0,
                                                nn_inter_ratio_thr=0.5)
    ume_loss, kp_v, kp_r, U_v, U_r, ratio, valid = ume_loss_fn(tv[0], tv[1], vft, tv[3], rft, tv[5])
    reg_loss, rre, rte = reg_loss_fn(tv[0], U_v, tv[3], U_r, tv[5], ratio, valid)
    gv, gr = torch.autograd.grad(ume_loss + reg_loss, [vft, rft])
    out.update({"ume_loss": np.asarray(ume_loss.item(), np.float32), "reg_loss": np.asarray(reg_loss.item(), np.float32),
                "loss_rre": rre.numpy(), "loss_rte": rte.numpy(), "loss_grad_velo_feat": gv.numpy(),
                "loss_grad_ref_feat": gr.numpy(), "loss_ume_velo": U_v.detach().numpy(), "loss_ume_ref": U_r.detach().numpy()})
    path = os.path.join(HERE, "training.npz")
    np.savez_compressed(path, **out)
    print("keypoints:", out["raw_kp_velo"].shape, "ratio mean", float(out["raw_ratio"].mean()),
          "ume_loss", float(ume_loss), "reg_loss", float(reg_loss), os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
