"""CPU, build container only: the oracle against the UNMODIFIED reference run live (imported from
/root/reference under the stubs of oracle/ref_import.py) on fresh seeded inputs — more pins
than the frozen golden files.  Skipped wherever the reference tree is absent (the GPU box)."""
import os
import sys
from types import SimpleNamespace

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import ref_import  # noqa: E402

from oracle import ume_oracle as orc  # noqa: E402
from umeregrobust_b200 import synth  # noqa: E402

pytestmark = pytest.mark.skipif(not ref_import.reference_available(), reason="reference tree not present")


@pytest.fixture(scope="module")
def ref():
    torch.set_num_threads(1)
    return ref_import.import_reference()


def t(x):
    return torch.from_numpy(np.ascontiguousarray(x))


def cloud(seed, B, N, C):
    rng = np.random.default_rng(seed)
    p = (np.stack([rng.uniform(-10, 10, (B, N)), rng.uniform(-10, 10, (B, N)), rng.uniform(-1, 1, (B, N))], -1)
         + np.array([25.0, -12.0, 0.3])).astype(np.float32)
    f = synth._normalize_rows(rng.normal(size=(B, N, C))).astype(np.float32)
    return p, f, rng


@pytest.mark.parametrize("seed,N,n,K,r", [(1, 1500, 40, 60, 2.5), (2, 900, 25, 2000, 3.0), (3, 2500, 30, 8, 1.5)])
def test_moments_cdist_rigid_chain(ref, seed, N, n, K, r):
    ev, loc, evu = ref
    pts, feat, rng = cloud(seed, 2, N, 32)                      # my_ume_generation hard-codes 32 channels (:55)
    kp = pts[:, rng.choice(N, n, replace=False)].copy()
    args = SimpleNamespace(ume_max_nn=K, ume_r_nn=r)
    F_ref = ev.my_ume_generation(t(pts), t(kp), t(feat), args).numpy()
    F, idx = orc.ume_moments(pts, kp, feat, K, r, return_idx=True)
    kappa = orc.normaliser_condition(feat, idx)
    err = (np.abs(F - F_ref).max(axis=(-1, -2)) / np.abs(F_ref).max(axis=(-1, -2)) / kappa).max()
    assert err < 1e-5
    pts2, feat2, _ = cloud(seed + 100, 2, N, 32)
    F2_ref = ev.my_ume_generation(t(pts2), t(kp), t(feat2), args).numpy()
    D_ref = loc.ume_cdist(t(F_ref), t(F2_ref)).numpy()
    D = orc.ume_cdist(F_ref, F2_ref)
    assert np.abs(D - D_ref).max() < 5e-3                       # both are fp32 mm-form cdist: sqrt noise near 0
    D64 = orc.ume_cdist_gram(F_ref, F2_ref)
    assert np.abs(D64 - D_ref).max() < 5e-3
    T_ref, Dp_ref = loc.batch_estimate_transform_ume_old(t(F_ref[0]), t(F2_ref[0]))
    T, Dp = orc.rigid_from_ume(F_ref[0], F2_ref[0])
    ang = orc.rotation_angle_rad(T[:, :3, :3].astype(np.float64), T_ref.numpy()[:, :3, :3].astype(np.float64))
    assert np.median(ang) < 1e-4
    assert np.abs(Dp - Dp_ref.numpy()).max() < 5e-3


def test_training_generation_and_gradient(ref):
    _, loc, _ = ref
    rng = np.random.default_rng(7)
    N, Nr, C = 1200, 1000, 16
    p = (np.stack([rng.uniform(-7, 7, N), rng.uniform(-7, 7, N), rng.uniform(-1, 1, N)], 1) + np.array([9.0, 4.0, 0.0])).astype(np.float32)
    gt = synth.random_rigid(rng, t_range=(1.0, 2.0)).astype(np.float32)
    sel = rng.permutation(N)[:Nr]
    q = ((p[sel] + rng.normal(scale=0.03, size=(Nr, 3))) @ gt[:3, :3].T.astype(np.float64) + gt[:3, 3]).astype(np.float32)
    f = synth._normalize_rows(rng.normal(size=(N, C))).astype(np.float32)
    g = synth._normalize_rows(f[sel] + rng.normal(scale=0.05, size=(Nr, C))).astype(np.float32)
    seg = rng.integers(0, 12, size=(1, N, 1)).astype(np.int64)
    vf = t(f[None]).clone().requires_grad_(True)
    kw = dict(nn_r=2.5, max_nn=150, min_nn=30, num_samples=16, flat_labels=[9], nn_intersection_r=0.6)
    F_v, F_r, kp_v, kp_r, ratio, cond = loc.generate_ume_from_keypoints2(t(p[None]), t(seg), vf, t(q[None]), t(g[None]),
                                                                          t(gt[None]), normalized_ume=False, **kw)
    w = torch.from_numpy(rng.normal(size=tuple(F_v.shape)).astype(np.float32))
    (gv,) = torch.autograd.grad((F_v * w).sum(), [vf])
    got = orc.ume_moments_backward(p[None], kp_v.numpy(), w.numpy(), kw["max_nn"], kw["nn_r"])
    assert np.abs(got - gv.numpy()).max() < 1e-4 * np.abs(gv.numpy()).max()


def test_sparse_quantize_semantics_note():
    # MinkowskiEngine is not installable here, so ME.utils.sparse_quantize cannot be run; the
    # restatement is pinned by its known-answer test (tests/test_oracle.py) only.  This test documents that.
    c = np.array([[0.0, 0.0, 0.0], [0.29, 0.0, 0.0], [0.31, 0.0, 0.0]], np.float32)
    assert orc.sparse_quantize(c, 0.3)[1].tolist() == [0, 2]
