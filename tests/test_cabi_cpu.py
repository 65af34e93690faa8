"""CPU: the C-ABI library builds, loads and exports every symbol include/umereg_b200.h declares;
argument validation that does not need a device; no compute calls."""
import ctypes
import os
import re

import pytest

from umeregrobust_b200 import _lib

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def handle():
    _lib.build()
    return _lib.lib()


def declared_symbols():
    text = open(os.path.join(REPO, "include", "umereg_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ume_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(handle):
    names = declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(handle, n), n
    assert sorted(_lib.EXPORTED_SYMBOLS) == names


def test_version_and_status_strings(handle):
    assert handle.ume_abi_version() == 1
    assert handle.ume_status_string(0) == b"ok"
    assert b"workspace" in handle.ume_status_string(-2)
    assert handle.ume_launch_count() >= 0


def test_workspace_queries(handle):
    assert handle.ume_moments_workspace_bytes(0, 100, 10, 32, 5) == 0
    w1 = handle.ume_moments_workspace_bytes(1, 120000, 1024, 32, 750)
    w2 = handle.ume_moments_workspace_bytes(2, 120000, 1024, 32, 750)
    assert w1 >= 120000 * 16 and w2 > w1
    assert handle.ume_ball_query_workspace_bytes(1, 10, 1000, 5) >= 1000 * 16
    assert handle.ume_cdist_workspace_bytes(1, 10, 10, 32, 0) == 0


def test_argument_validation_without_a_device(handle):
    # validation happens before any CUDA call, so it can be exercised on a CPU-only box
    null = ctypes.c_void_p(None)
    rc = handle.ume_moments_f32(null, null, null, 1, 10, 2, 32, 5, 1.0, 0, null, null, null, null, 0, null)
    assert rc == -1 and b"null" in handle.ume_last_error()
    one = ctypes.c_void_p(256)
    rc = handle.ume_moments_f32(one, one, one, 1, 10, 2, 300, 5, 1.0, 0, one, null, null, one, 1 << 30, null)
    assert rc == -3 and b"C = 300" in handle.ume_last_error()
    rc = handle.ume_moments_f32(one, one, one, 1, 10, 2, 32, 5, 1.0, 0, one, null, null, one, 16, null)
    assert rc == -2
    rc = handle.ume_ball_query_f32(one, one, 1, 4, 10, 100000, 1.0, 0, null, null, null, null, one, 1 << 30, null)
    assert rc == -3
    rc = handle.ume_cdist_f32(one, one, 1, 4, 4, 30, 0, null, null, null, null, 0, null)
    assert rc == -3
    rc = handle.ume_rigid_solve_f32(one, one, null, null, one, null, 1, 4, 4, 4, 32, one, null)
    assert rc == -1
    # empty problems are no-ops, not errors
    assert handle.ume_moments_f32(null, null, null, 0, 0, 0, 32, 5, 1.0, 0, null, null, null, null, 0, null) == 0
    assert handle.ume_orthonormalize_f32(null, 0, 32, null, null, null) == 0


def test_wrappers_refuse_cpu_tensors():
    import torch
    import umeregrobust_b200 as ume
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ume.ume_cdist(torch.zeros(1, 2, 32, 4), torch.zeros(1, 2, 32, 4))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ume.my_ume_generation(torch.zeros(1, 5, 3), torch.zeros(1, 2, 3), torch.zeros(1, 5, 32),
                              type("A", (), dict(ume_max_nn=4, ume_r_nn=1.0))())
    with pytest.raises(ValueError):
        ume.batch_estimate_transform_ume_old(torch.zeros(2, 32, 4), torch.zeros(3, 32, 4))


def test_patch_reference_rebinds_names():
    import types
    import umeregrobust_b200 as ume
    from umeregrobust_b200 import training
    fake_eval = types.ModuleType("evaluate")
    fake_loc = types.ModuleType("utils.loc_utils")
    fake_loss = types.ModuleType("loss")
    names = ("ball_query", "knn_points", "knn_gather", "ume_cdist", "batch_estimate_transform_ume_old", "ume_kp_layer")
    for m in (fake_eval, fake_loc):
        for n in names:
            setattr(m, n, object())
    fake_eval.my_ume_generation = object()
    for n in ("generate_ume_from_keypoints2", "UMEContrastiveLoss", "CubeRegistrationLoss", "ume_cdist"):
        setattr(fake_loss, n, object())
    fake_loc.generate_ume_from_keypoints2 = object()
    loc_before = {n: getattr(fake_loc, n) for n in names}
    try:
        # default: `evaluate` only, and the distance test in pytorch3d's CUDA arithmetic (FMA)
        done = ume.patch_reference(fake_eval, fake_loc)
        assert fake_eval.my_ume_generation is ume.my_ume_generation and fake_eval.ball_query is ume.ball_query
        assert all(getattr(fake_loc, n) is v for n, v in loc_before.items())      # the losses import from here: untouched
        assert ("evaluate", "my_ume_generation") in done and len(done) == 7
        assert ume.config["fma_dist"] is True
        # opt-in: the inference kernels inside utils.loc_utils too
        done = ume.patch_reference(fake_eval, fake_loc, fma_dist=False, patch_loc_utils=True)
        assert fake_loc.ume_cdist is ume.ume_cdist and ("utils.loc_utils", "ume_cdist") in done
        assert ume.config["fma_dist"] is False
        # training: differentiable mirrors into utils.loc_utils / loss, nothing without autograd
        for n, v in loc_before.items():
            setattr(fake_loc, n, v)
        done = ume.patch_reference(fake_eval, fake_loc, training=True, loss_module=fake_loss)
        assert fake_loss.UMEContrastiveLoss is training.UMEContrastiveLoss
        assert fake_loss.CubeRegistrationLoss is training.CubeRegistrationLoss
        assert fake_loc.generate_ume_from_keypoints2 is training.generate_ume_from_keypoints2
        assert fake_loc.ume_cdist is loc_before["ume_cdist"] and not isinstance(fake_loss.ume_cdist, types.FunctionType)
        with pytest.raises(ValueError):
            ume.patch_reference(fake_eval, fake_loc, patch_loc_utils=True, training=True)
        with pytest.raises(RuntimeError):
            ume.patch_reference(None, None)
    finally:
        ume.config["fma_dist"] = False


def test_new_entry_points_validate_without_a_device(handle):
    null = ctypes.c_void_p(None)
    one = ctypes.c_void_p(256)
    # backward: C outside the warp kernel's channel counts
    rc = handle.ume_moments_backward_f32(one, one, one, 1, 10, 2, 12, 5, 1.0, 0, one, one, 1 << 30, null)
    assert rc == -3 and b"C = 12" in handle.ume_last_error()
    rc = handle.ume_moments_backward_f32(one, one, null, 1, 10, 2, 32, 5, 1.0, 0, one, one, 1 << 30, null)
    assert rc == -1
    rc = handle.ume_neighbor_count_f32(one, one, 1, 10, 2, 5, 1.0, 0, null, one, 1 << 30, null)
    assert rc == -1
    rc = handle.ume_neighbor_count_f32(one, one, 1, 10, 2, 5, 1.0, 0, one, one, 16, null)
    assert rc == -2
    # raw moments need the warp kernel's channel counts
    rc = handle.ume_moments_f32(one, one, one, 1, 10, 2, 12, 5, 1.0, _lib.UME_FLAG_RAW_MOMENTS, one, null, null, one, 1 << 30, null)
    assert rc == -3
    rc = handle.ume_voxel_unique_f32(one, 10, 0.0, one, null, one, one, 1 << 30, null)
    assert rc == -1 and b"voxel size" in handle.ume_last_error()
    assert handle.ume_voxel_unique_workspace_bytes(0) == 0 and handle.ume_voxel_unique_workspace_bytes(1000) >= 2048 * 12


def test_training_helpers_on_cpu():
    # host-side logic of the training mirror that needs no device: the "selected rows in descending
    # order, zero padded" idiom of utils/loc_utils.py:104-111
    torch = pytest.importorskip("torch")
    from umeregrobust_b200 import training
    cond = torch.tensor([[True, False, True, True, False], [False, False, False, True, False]])
    idx, n = training._descending(cond)
    assert idx.tolist() == [[3, 2, 0, 0, 0], [3, 0, 0, 0, 0]] and n.tolist() == [3, 1]
    # the reference idiom, literally
    mask = -1 * torch.ones_like(cond, dtype=torch.long)
    w = torch.where(cond)
    mask[w] = w[1]
    mask = mask.sort(dim=1, descending=True)[0]
    lengths = (mask > -1).sum(dim=-1)
    mask[mask == -1] = 0
    assert torch.equal(mask, idx) and torch.equal(lengths, n)
    # the differentiable solve / distance are CUDA kernels forward and backward: CPU tensors are refused, not
    # silently routed through torch
    G = torch.rand(3, 8, 4)
    with pytest.raises(RuntimeError):
        training.rigid_from_ume_autograd(G, G)
    with pytest.raises(RuntimeError):
        training.ume_cdist_autograd(G[None], G[None])
