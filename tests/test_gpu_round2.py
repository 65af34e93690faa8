"""GPU tests added in round 2 (VERDICT r1 "Next round" items 1, 3, 4, 8 and the advisor's findings):
the drop-in proven against the reference's OWN evaluate.py source, the warp-per-keypoint kernel
against the oracle at a launch size that actually selects it, the device-side match sub-sampling,
the streamed engine path, per-graph arenas, the metric kernel."""
import os
import textwrap
from types import SimpleNamespace

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import pytorch3d_ops as p3d
from oracle import ume_oracle as orc
from umeregrobust_b200 import synth

pytestmark = pytest.mark.gpu

K_NN, RADIUS = 750, 5.0
KEYS = ("src_pts", "src_feat", "src_kp", "tgt_pts", "tgt_feat", "tgt_kp")


@pytest.fixture(scope="module")
def ume():
    import umeregrobust_b200 as u
    from umeregrobust_b200 import _lib
    _lib.lib()
    yield u
    u.config.update(fma_dist=False, cell_div2=False, cdist_impl=None, cta_moments=False, warp_moments=False)


def dev(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def host(t):
    return t.detach().cpu().numpy()


# ----------------------------------------------------------------------------- the drop-in, for real
def _reference_eval_module():
    from oracle import ref_import
    if not ref_import.reference_available():
        pytest.skip("no reference sources (neither /root/reference nor the staged baseline/_ref)")
    ev, loc, evu = ref_import.import_reference()
    return ev, loc, evu, ref_import.REF_ROOT


@pytest.mark.parametrize("fma", [False, True])
def test_reference_eval_loop_source_runs_on_the_patched_names(ume, golden, fma):
    """VERDICT r1 missing #6: the reference's own evaluate.py:206-257 (UME generation, ume_cdist, arg-min,
    gathers, batch_estimate_transform_ume_old — the source lines themselves, exec'ed unchanged inside the
    reference's `evaluate` module namespace) with `patch_reference()` applied, on the GPU, against the
    golden output of the unpatched reference (tests/golden/hotpath_noisy.npz)."""
    ev, loc, evu, root = _reference_eval_module()
    saved = {k: getattr(ev, k) for k in ("ball_query", "knn_points", "knn_gather", "my_ume_generation", "ume_cdist",
                                          "batch_estimate_transform_ume_old", "ume_kp_layer", "FeatureCorrelator")}
    loc_before = {k: getattr(loc, k) for k in ("ume_cdist", "batch_estimate_transform_ume_old")}
    try:
        done = ume.patch_reference(ev, fma_dist=fma)
        assert ("evaluate", "my_ume_generation") in done and ("evaluate", "ball_query") in done
        assert ev.my_ume_generation is ume.my_ume_generation and ev.ume_cdist is ume.ume_cdist
        assert ume.config["fma_dist"] is fma
        # utils.loc_utils is left alone by default (the training losses need autograd there)
        assert all(getattr(loc, k) is v for k, v in loc_before.items())
        g = golden("hotpath_noisy")
        lines = open(os.path.join(root, "evaluate.py")).read().splitlines()
        assert "ume_src = my_ume_generation(src_pts, src_pts_ds, src_feat, args)" in lines[205]
        assert "rtume_tform = T.view(" in lines[253]
        code = textwrap.dedent("\n".join(lines[205:257]))                     # evaluate.py:206-257
        ns = dict(vars(ev))                                                  # the module's globals, patched names included
        inject = dict(src_pts=dev(g["src_pts"][None]), src_pts_ds=dev(g["src_kp"][None]), src_feat=dev(g["src_feat"][None]),
                      tgt_pts=dev(g["tgt_pts"][None]), tgt_pts_ds=dev(g["tgt_kp"][None]), tgt_feat=dev(g["tgt_feat"][None]),
                      args=SimpleNamespace(ume_max_nn=int(g["K"]), ume_r_nn=float(g["radius"]), hungarian_matching_flag=False,
                                           filter_by_ume_dist_cond=False, batch_size=1, device="cuda"))
        ns.update(inject)
        exec(compile(code, "evaluate.py[206:257]", "exec"), ns)
        T = host(ns["rtume_tform"])[0]
        m = host(ns["m"])[0]
        D = host(ns["D"])[0]
        F_src = host(ns["ume_src"])                                            # gathered by m[...,0] = identity
        G_in, H_in = host(ns["G"]), host(ns["H"])                              # what batch_estimate_transform_ume_old was given
    finally:
        for k, v in saved.items():
            setattr(ev, k, v)
        ume.config["fma_dist"] = False
    assert T.shape == g["T"].shape and m.shape == g["match"][0].shape and m.dtype == np.int64
    assert np.abs(D - g["D"][0]).max() < 3e-3                                  # the reference's own mm-cdist noise
    D64 = orc.ume_cdist(g["F_src"], g["F_tgt"], dtype=np.float64)[0]
    srt = np.sort(D64, -1)
    clear = (srt[:, 1] - srt[:, 0]) > 1e-4
    flips_clear = int((m[clear, 1] != g["match"][0, clear, 1]).sum())
    flips_unclear = int((m[~clear, 1] != g["match"][0, ~clear, 1]).sum())
    print("arg-min vs the reference golden: %d flips on %d clear rows, %d flips on %d near-tie rows (gap <= 1e-4)"
          % (flips_clear, int(clear.sum()), flips_unclear, int((~clear).sum())))
    assert flips_clear == 0
    assert flips_unclear <= max(1, int(0.5 * (~clear).sum()))
    kappa = orc.normaliser_condition(g["src_feat"][None], g["bq_idx_src"].astype(np.int64))
    err = np.abs(F_src - g["F_src"]).max(axis=(-1, -2)) / np.abs(g["F_src"]).max(axis=(-1, -2)) / kappa
    # fma=True may move a neighbour that sits within one ulp of r^2 in or out (pytorch3d CUDA vs CPU arithmetic)
    assert np.median(err) < 3e-6 and (fma or err.max() < 3e-6)
    # the solve, judged on ITS inputs (this path's own fp32 moment matrices, gathered by evaluate.py:230-231):
    # fp64 oracle on the same matrices; tolerance 1e-4 rad / 1e-4 m or the reference's own fp32 distance to
    # fp64 on the same hypothesis where that is larger (absolute coordinates: SURVEY §7 "Tolerance")
    T64, _ = orc.rigid_from_ume(G_in.astype(np.float64), H_in.astype(np.float64), dtype=np.float64)
    same = m[:, 1] == g["match"][0, :, 1]
    Tr64, _ = orc.rigid_from_ume(g["F_src"][0][g["match"][0, :, 0]], g["F_tgt"][0][g["match"][0, :, 1]], dtype=np.float64)
    ang = orc.rotation_angle_rad(T[:, :3, :3], T64[:, :3, :3])
    terr = np.abs(T[:, :3, 3] - T64[:, :3, 3]).max(-1)
    ang_ref = orc.rotation_angle_rad(g["T"][:, :3, :3], Tr64[:, :3, :3])
    terr_ref = np.abs(g["T"][:, :3, 3] - Tr64[:, :3, 3]).max(-1)
    assert (ang[same] <= np.maximum(1e-4, 3 * ang_ref[same])).all(), (ang.max(), ang_ref.max())
    assert (terr[same] <= np.maximum(1e-4, 3 * terr_ref[same])).all(), (terr.max(), terr_ref.max())
    assert np.median(ang) <= max(1e-4, 2 * np.median(ang_ref)) and np.median(terr) <= max(1e-4, 2 * np.median(terr_ref))
    # and against the reference's golden transforms themselves where the match is the same
    dT = np.abs(T[same] - g["T"][same]).max(axis=(-1, -2))
    assert np.median(dT) < 1e-3, np.median(dT)


def test_inference_kernels_refuse_inputs_that_require_grad(ume):
    F = torch.randn(1, 8, 32, 4, device="cuda", requires_grad=True)
    with pytest.raises(RuntimeError, match="not differentiable"):
        ume.ume_cdist(F, F)
    with torch.no_grad():
        ume.ume_cdist(F, F)


# ----------------------------------------------------------------------------- ume_kp_layer(n_rand=...)
def test_kp_layer_n_rand_golden(ume, golden):
    """utils/loc_utils.py:406-410 with the host RNG seeded like the golden run: same triplets, so
    T / D are comparable hypothesis by hypothesis."""
    g = golden("kp_layer_nrand")
    args = [dev(g[k][None]) for k in KEYS]
    layer = ume.ume_kp_layer(int(g["ume_knn"]), float(g["ume_desc_rad"]), diag_only=True, n_rand=int(g["n_rand"]))
    np.random.seed(int(g["np_seed"]))
    T, D, G_kp, H_kp = layer(*args)
    assert tuple(T.shape) == g["T"].shape and tuple(D.shape) == g["D"].shape
    assert tuple(G_kp.shape) == g["G"].shape and tuple(H_kp.shape) == g["H"].shape
    assert np.abs(host(G_kp) - g["G"]).max() / np.abs(g["G"]).max() < 1e-4
    # fp64 restatement of the same triplet sums
    np.random.seed(int(g["np_seed"]))
    tri = np.random.choice(np.arange(g["G"].shape[0]), (int(g["n_rand"]), 3))
    G64 = g["G"].astype(np.float64)[tri[:, 0]] + g["G"].astype(np.float64)[tri[:, 1]] + g["G"].astype(np.float64)[tri[:, 2]]
    H64 = g["H"].astype(np.float64)[tri[:, 0]] + g["H"].astype(np.float64)[tri[:, 1]] + g["H"].astype(np.float64)[tri[:, 2]]
    T64, D64 = orc.rigid_from_ume(G64, H64, dtype=np.float64)
    ang = orc.rotation_angle_rad(host(T)[0][:, :3, :3], T64[:, :3, :3])
    ang_ref = orc.rotation_angle_rad(g["T"][0][:, :3, :3], T64[:, :3, :3])
    assert (ang <= np.maximum(2e-4, 3 * ang_ref)).all(), (ang.max(), ang_ref.max())
    assert np.abs(host(D)[0] - g["D"][0]).max() < 3e-3


# ----------------------------------------------------------------------------- a10: the metric
def test_relative_rotation_error_kernel(ume, golden):
    rng = np.random.default_rng(5)
    Ra = np.stack([synth.random_rigid(rng, max_tilt_deg=180.0)[:3, :3] for _ in range(300)]).astype(np.float32)
    Rb = np.stack([synth.random_rigid(rng, max_tilt_deg=180.0)[:3, :3] for _ in range(300)]).astype(np.float32)
    Rb[:20] = Ra[:20]                                                       # zero angle: the clamp at trace = 3
    got = host(ume.relative_rotation_error(dev(Ra), dev(Rb)))
    want = orc.relative_rotation_error(Ra.astype(np.float64), Rb.astype(np.float64))
    far = want > 1.0
    assert np.abs(got[far] - want[far]).max() < 2e-3                        # degrees; fp32 acos
    assert np.abs(got[~far] - want[~far]).max() < 0.06                      # acos loses digits near 0
    # the reference's own fp32 result (torch CPU) on the golden hypotheses, and in-place reads of (n,4,4) transforms
    g = golden("hotpath_noisy")
    T = dev(g["T"])
    R_gt = dev(np.broadcast_to(g["gt"][:3, :3], (g["T"].shape[0], 3, 3)).copy())
    got = host(ume.relative_rotation_error(T[:, :3, :3], R_gt))
    assert np.abs(got - g["rre"]).max() < 0.06
    big = g["rre"] > 1.0
    assert np.abs(got[big] - g["rre"][big]).max(initial=0) < 2e-3
    with pytest.raises(RuntimeError):
        ume.relative_rotation_error(torch.zeros(2, 3, 3), torch.zeros(2, 3, 3))


# ----------------------------------------------------------------------------- warp kernel at its launch size
@pytest.mark.parametrize("name,N,C,n_kp,model", [("kitti", 120000, 32, 1024, "KITTI"), ("nuscenes", 35000, 32, 1024, "NUSCENES"),
                                                 ("rotkitti", 120000, 64, 2048, "KITTI")])
def test_warp_kernel_against_oracle_at_benchmark_launch_size(ume, name, N, C, n_kp, model):
    """a launch of >= 3072 keypoints (4 pairs) so that the DEFAULT dispatch picks
    the warp-per-keypoint kernel — the one the benchmark times — and that kernel is what meets the
    oracle: neighbour counts exact, F < 3e-6 kappa, matches and transforms vs fp64."""
    B = 4 if n_kp == 1024 else 2
    b = synth.make_batch(B, seed0=61, n_base=2, N=N, C=C, n_kp=n_kp, model=getattr(synth, model))
    assert B * n_kp >= 3072
    d = {k: dev(b[k]) for k in KEYS}
    out = ume.register_hypotheses(*[d[k] for k in KEYS], K_NN, RADIUS, want_D=True)
    _, cnt = ume.ume_moments(d["src_pts"], d["src_kp"], d["src_feat"], K_NN, RADIUS, return_count=True)
    # the CTA kernel on the same launch: same neighbours, sums within rounding of each other
    ume.config["cta_moments"] = True
    F_cta = host(ume.ume_moments(d["src_pts"], d["src_kp"], d["src_feat"], K_NN, RADIUS))
    ume.config["cta_moments"] = False
    F = host(out["F_src"])
    D = host(out["D"])
    am = host(out["match"])[..., 1]
    assert np.array_equal(am, np.argmin(D, -1))
    for p in (0, B - 1):                                                   # the oracle takes seconds per pair
        F64, idx = orc.ume_moments(b["src_pts"][p:p + 1], b["src_kp"][p:p + 1], b["src_feat"][p:p + 1], K_NN, RADIUS,
                                   dtype=np.float64, return_idx=True)
        assert np.array_equal(host(cnt)[p], (idx >= 0).sum(-1)[0])
        kappa = orc.normaliser_condition(b["src_feat"][p:p + 1], idx)
        err = np.abs(F[p:p + 1] - F64).max(axis=(-1, -2)) / np.abs(F64).max(axis=(-1, -2)) / kappa
        assert err.max() < 3e-6, err.max()
        err_cta = np.abs(F_cta[p:p + 1] - F64).max(axis=(-1, -2)) / np.abs(F64).max(axis=(-1, -2)) / kappa
        assert err_cta.max() < 3e-6
        G64 = orc.ume_moments(b["tgt_pts"][p:p + 1], b["tgt_kp"][p:p + 1], b["tgt_feat"][p:p + 1], K_NN, RADIUS, dtype=np.float64)
        D64 = orc.ume_cdist_gram(F64, G64)
        assert np.abs(D[p:p + 1] - D64)[D64 > 0.05].max(initial=0) < 1e-4
        srt = np.sort(D64, -1)
        clear = (srt[..., 1] - srt[..., 0]) > 1e-4
        assert clear.mean() > 0.9
        assert np.array_equal(am[p:p + 1][clear], np.argmin(D64, -1)[clear])
        same = am[p] == np.argmin(D64, -1)[0]
        T64, _ = orc.rigid_from_ume(F64[0][np.nonzero(same)[0]], G64[0][am[p][same]], dtype=np.float64, with_distance=False)
        T = host(out["T"])[p][same]
        ang = orc.rotation_angle_rad(T[:, :3, :3], T64[:, :3, :3])
        terr = np.abs(T[:, :3, 3] - T64[:, :3, 3]).max(-1)
        assert np.median(ang) < 1e-4 and np.median(terr) < 1e-4, (np.median(ang), np.median(terr))


def test_rotkitti_ground_truth_recovery_through_the_streamed_engine(ume):
    """BASELINE config #5 in miniature: pairs whose ground truth is cycled from the reference's RotKITTI
    transforms (28.8-180 deg), streamed from pinned host memory through `register_stream` in packed
    micro-batches (>= 3072 keypoints each: the warp kernel); the streamed results equal the
    device-resident ones bit for bit, and every pair has hypotheses that recover its large rotation."""
    import bench
    from umeregrobust_b200.engine import RegistrationEngine, PackedPairs
    wl = dict(N=120000, n_kp=2048, C=64, model="KITTI", gt="rotkitti")
    mbs = [bench.make_pairs(wl, 2, seed0=7 + 10 * j) for j in range(2)]
    gts = bench.rotkitti_gt()
    for j, mb in enumerate(mbs):
        assert np.abs(mb["gt"] - gts[[(7 + 10 * j) % 600, (8 + 10 * j) % 600]]).max() < 1e-3
    eng = RegistrationEngine(K=K_NN, radius=RADIUS)
    packed = [PackedPairs.from_arrays(mb) for mb in mbs]
    res = eng.register_stream(packed * 2)                                  # 4 micro-batches, the pool cycled
    torch.cuda.synchronize()
    res = {k: v.clone() for k, v in res.items()}
    assert tuple(res["T"].shape) == (8, 2048, 4, 4) and tuple(res["match"].shape) == (8, 2048, 2)
    for j, mb in enumerate(mbs):
        d = {k: dev(mb[k]) for k in KEYS}
        out = RegistrationEngine(K=K_NN, radius=RADIUS).register(d)
        for rep in (0, 1):
            lo = 2 * j + 4 * rep
            assert np.array_equal(res["T"][lo:lo + 2].numpy(), host(out["T"]))
            assert np.array_equal(res["match"][lo:lo + 2].numpy(), host(out["match"]))
            assert np.array_equal(res["dmin"][lo:lo + 2].numpy(), host(out["dmin"]))
        chk = bench.gt_check(torch, out["T"], dev(mb["gt"]))
        assert chk["pairs_with_a_correct_hypothesis"] == 2, chk
        assert chk["median_share_of_correct_hypotheses"] > 0.3, chk          # ~0.7 measured with the fp32 oracle
        # the hypothesis with the smallest descriptor distance is a correct one (cheap stand-in for the correlator)
        best = out["dmin"].argmin(dim=1)
        Tb = out["T"][torch.arange(2), best]
        g = dev(mb["gt"])
        rre = host(ume.relative_rotation_error(g[:, :3, :3], Tb[:, :3, :3]))
        assert (rre < 5.0).all(), rre


# ----------------------------------------------------------------------------- f2: device-side sub-sampling
def test_gumbel_topk_with_the_uniforms_fed_in_as_data(ume):
    rng = np.random.default_rng(11)
    B, n, k, tau = 3, 1000, 250, 0.05
    d = rng.uniform(0.0, 1.2, (B, n)).astype(np.float32)
    u = rng.uniform(0.0, 1.0, (B, n)).astype(np.float32)
    idx = host(ume.weighted_match_subsample(dev(d), tau, k, u=dev(u)))
    uc = np.clip(u, 1e-20, 1.0 - 1e-7).astype(np.float64)
    keys = (1.0 - d.astype(np.float64)) / tau - np.log(-np.log(uc))
    for b in range(B):
        assert (np.diff(idx[b]) > 0).all() and idx[b].min() >= 0 and idx[b].max() < n      # ascending, distinct, in range
        want = np.sort(np.argsort(-keys[b], kind="stable")[:k])
        miss = np.setdiff1d(want, idx[b])
        extra = np.setdiff1d(idx[b], want)
        assert len(miss) == len(extra) <= 2
        if len(miss):                                                                       # only fp32-vs-fp64 near-ties at the cut
            cut = np.sort(keys[b])[::-1][k - 1]
            assert np.abs(keys[b][np.concatenate([miss, extra])] - cut).max() < 1e-3
    # the 1-D form and k >= n
    one = host(ume.weighted_match_subsample(dev(d[0]), tau, k, u=dev(u[0])))
    assert np.array_equal(one, idx[0])
    assert np.array_equal(host(ume.weighted_match_subsample(dev(d[0]), tau, 5000, seed=3)), np.arange(n))


def test_gumbel_topk_distribution_matches_successive_sampling(ume):
    """Inclusion frequencies of the device sampler against np.random.choice(p, replace=False) — the
    reference's sampler at evaluate.py:240 — on a small problem (n = 12, k = 4, 20000 draws)."""
    rng = np.random.default_rng(3)
    n, k, tau, draws = 12, 4, 0.25, 20000
    d = rng.uniform(0.2, 1.0, n).astype(np.float32)
    p = np.exp((1 - d.astype(np.float64)) / tau)
    p /= p.sum()
    D = dev(np.tile(d, (draws, 1)))
    idx = host(ume.weighted_match_subsample(D, tau, k, seed=1234))
    freq = np.bincount(idx.ravel(), minlength=n) / draws
    ref = np.zeros(n)
    rs = np.random.RandomState(0)
    for _ in range(draws):
        ref[rs.choice(n, k, replace=False, p=p)] += 1
    ref /= draws
    assert np.abs(freq - ref).max() < 0.02, (freq, ref)
    # a different seed gives different draws; the same seed the same
    assert not np.array_equal(idx, host(ume.weighted_match_subsample(D, tau, k, seed=1235)))
    assert np.array_equal(idx, host(ume.weighted_match_subsample(D, tau, k, seed=1234)))


def test_engine_subsample_solves_only_the_drawn_matches(ume):
    from umeregrobust_b200.engine import RegistrationEngine
    b = synth.make_batch(4, seed0=71, n_base=2, N=35000, C=32, n_kp=1024, model=synth.NUSCENES)
    d = {k: dev(b[k]) for k in KEYS}
    full = {k: host(v).copy() for k, v in RegistrationEngine(K=K_NN, radius=RADIUS).register(d).items() if v is not None}
    eng = RegistrationEngine(K=K_NN, radius=RADIUS, subsample=256, tau=0.05, seed=5)
    out = eng.register(d)
    m = host(out["match"])
    assert m.shape == (4, 256, 2) and tuple(out["T"].shape) == (4, 256, 4, 4)
    for p in range(4):
        sel = m[p, :, 0]
        assert (np.diff(sel) > 0).all()
        assert np.array_equal(m[p, :, 1], full["match"][p, sel, 1])
        assert np.array_equal(host(out["T"])[p], full["T"][p, sel])          # same solve on the drawn matches
        assert np.array_equal(host(out["dmin"])[p], full["dmin"][p, sel])
    # low-distance matches are preferred: the drawn matches' mean distance is below the population's
    assert host(out["dmin"]).mean() < full["dmin"].mean()


# ----------------------------------------------------------------------------- graphs own their arenas
def test_graphs_of_different_shapes_do_not_share_memory(ume):
    """Advisor r1 (medium): graphed(A), graphed(B with another shape), graphed(A) — each graph owns its
    arena and workspace, so replaying A after B was captured still equals eager."""
    from umeregrobust_b200.engine import RegistrationEngine
    bA = synth.make_batch(2, seed0=81, n_base=2, N=30000, C=32, n_kp=256, model=synth.NUSCENES)
    bB = synth.make_batch(3, seed0=82, n_base=2, N=20000, C=32, n_kp=192, model=synth.NUSCENES)
    dA = {k: dev(bA[k]) for k in KEYS}
    dB = {k: dev(bB[k]) for k in KEYS}
    eager = RegistrationEngine(K=K_NN, radius=RADIUS, want_D=True)
    refA = {k: host(v).copy() for k, v in eager.register(dA).items() if v is not None}
    refB = {k: host(v).copy() for k, v in eager.register(dB).items() if v is not None}
    eng = RegistrationEngine(K=K_NN, radius=RADIUS, want_D=True)
    eng.register_graphed(dA)
    outB = eng.register_graphed(dB)
    eng.register(dB)                                                        # eager call with another shape in between
    junk = [torch.empty(3_000_000, device="cuda").normal_() for _ in range(8)]   # anything freed would be reused here
    outA = eng.register_graphed(dA)
    torch.cuda.synchronize()
    for k in ("match", "D", "T", "dmin"):
        assert np.array_equal(host(outA[k]), refA[k]), k
    outB = eng.register_graphed(dB)
    torch.cuda.synchronize()
    for k in ("match", "D", "T", "dmin"):
        assert np.array_equal(host(outB[k]), refB[k]), k
    del junk


def test_result_pack_is_what_the_kernels_write(ume):
    """The step's T / arg-min / dmin live in ONE buffer (the all-gather payload): the views the engine
    returns alias it."""
    from umeregrobust_b200.engine import RegistrationEngine
    b = synth.make_batch(2, seed0=91, n_base=2, N=20000, C=32, n_kp=128, model=synth.NUSCENES)
    d = {k: dev(b[k]) for k in KEYS}
    eng = RegistrationEngine(K=K_NN, radius=RADIUS)
    out = eng.register(d)
    pack = eng._packs[0]
    v = pack.views()
    assert out["T"].data_ptr() == v["T"].data_ptr() and out["dmin"].data_ptr() == v["dmin"].data_ptr()
    assert torch.equal(out["match"][..., 1], v["argmin"])
    assert pack.nbytes >= 2 * 128 * (64 + 8 + 4)
    av = pack.all_views()
    assert torch.equal(av["T"][0], out["T"])


# ----------------------------------------------------------------------------- fp16-split tensor-core distance
@pytest.mark.parametrize("C,B,n1,n2", [(32, 3, 200, 333), (64, 2, 130, 97), (32, 1, 1024, 1024)])
def test_cdist_presplit_operands_equal_the_generic_entry(ume, C, B, n1, n2):
    """The orthonormalisation kernel writes the distance kernel's [hi | lo] fp16 operands itself; the result
    must be bit-identical to the generic entry (fp32 descriptors -> split pre-pass -> same kernel), and
    both within fp32-grade distance of the fp64 oracle (the 3 x fp16 split keeps 22 mantissa bits)."""
    rng = np.random.default_rng(C + n1)
    F1 = rng.normal(size=(B, n1, C, 4)).astype(np.float32) * np.array([1.0, 30.0, 30.0, 3.0], np.float32)
    F2 = rng.normal(size=(B, n2, C, 4)).astype(np.float32) * np.array([1.0, 30.0, 30.0, 3.0], np.float32)
    F2[:, : min(n1, n2) // 2] = F1[:, : min(n1, n2) // 2] + 1e-3 * rng.normal(size=(B, min(n1, n2) // 2, C, 4)).astype(np.float32)
    Qh1, Qt1 = ume.ume_descriptors_split(dev(F1), want_Qt=True)
    Qh2, Qt2 = ume.ume_descriptors_split(dev(F2), want_Qt=True)
    assert torch.equal(Qt1, ume.ume_descriptors(dev(F1)))
    Ds, ams, dms = ume.descriptor_cdist_split(Qh1, Qh2, want_D=True, want_argmin=True)
    Dg, amg, dmg = ume.descriptor_cdist(Qt1, Qt2, want_D=True, want_argmin=True, impl=1)
    assert torch.equal(Ds, Dg) and torch.equal(ams, amg) and torch.equal(dms, dmg)
    D64 = orc.ume_cdist_gram(F1.astype(np.float64), F2.astype(np.float64))
    D = host(Ds)
    assert np.abs(D - D64)[D64 > 0.05].max() < 2e-5
    assert np.abs(D - D64).max() < 3e-3                                   # sqrt near D = 0
    assert np.array_equal(host(ams), np.argmin(D, -1))
    # the operands: hi + lo reproduces 256 q to ~2^-22
    q = host(Qt1).astype(np.float64) * 256.0
    h = host(Qh1).astype(np.float64)
    assert np.abs(h[..., :C] + h[..., C:] - q).max() < 256.0 * 2.0 ** -21


# ----------------------------------------------------------------------------- source + target in one launch
@pytest.mark.parametrize("C,n", [(32, 1600), (64, 1550)])
def test_pair_launch_is_bit_identical_to_two_launches(ume, C, n):
    """`ume_moments_pair_f32` (one grid build, one moment launch for 2B clouds) against two `ume_moments_f32` calls:
    the cell-sorted arrays are a pure function of each cloud, so F and Fc must agree bit for bit; and the registration
    step built on it returns exactly what the two-launch step returns."""
    b = synth.make_batch(2, seed0=5, n_base=2, N=30000, C=C, n_kp=n)
    d = {k: dev(v) for k, v in b.items() if k.endswith(("pts", "feat", "kp"))}
    ume.config["warp_moments"] = True
    try:
        F1, Fc1 = ume.ume_moments(d["src_pts"], d["src_kp"], d["src_feat"], 750, 5.0, return_centered=True)
        F2, Fc2 = ume.ume_moments(d["tgt_pts"], d["tgt_kp"], d["tgt_feat"], 750, 5.0, return_centered=True)
    finally:
        ume.config["warp_moments"] = False
    pair = ume.ume_moments_pair(d["src_pts"], d["src_kp"], d["src_feat"], d["tgt_pts"], d["tgt_kp"], d["tgt_feat"], 750, 5.0,
                                return_centered=True)
    assert pair is not None
    G1, Gc1, G2, Gc2, both = pair
    assert torch.equal(G1, F1) and torch.equal(Gc1, Fc1) and torch.equal(G2, F2) and torch.equal(Gc2, Fc2)
    assert both.shape[0] == 4 and both.data_ptr() == Gc1.data_ptr()
    # ... and with the CTA-per-keypoint kernel (what small launches get): bit-identical to two CTA-kernel launches
    ume.config["cta_moments"] = True
    try:
        H1, Hc1 = ume.ume_moments(d["src_pts"], d["src_kp"], d["src_feat"], 750, 5.0, return_centered=True)
        H2, Hc2 = ume.ume_moments(d["tgt_pts"], d["tgt_kp"], d["tgt_feat"], 750, 5.0, return_centered=True)
        P1, Pc1, P2, Pc2, _ = ume.ume_moments_pair(d["src_pts"], d["src_kp"], d["src_feat"], d["tgt_pts"], d["tgt_kp"], d["tgt_feat"],
                                                   750, 5.0, return_centered=True)
    finally:
        ume.config["cta_moments"] = False
    assert torch.equal(P1, H1) and torch.equal(Pc1, Hc1) and torch.equal(P2, H2) and torch.equal(Pc2, Hc2)
    # different shapes on the two sides: the pair entry declines, the step falls back to two launches
    assert ume.ume_moments_pair(d["src_pts"], d["src_kp"], d["src_feat"], d["tgt_pts"][:, :-8], d["tgt_kp"], d["tgt_feat"][:, :-8],
                                750, 5.0) is None
    out = ume.register_hypotheses(d["src_pts"], d["src_feat"], d["src_kp"], d["tgt_pts"], d["tgt_feat"], d["tgt_kp"], 750, 5.0, want_D=True)
    ume.config["cta_moments"] = True             # the pair entry with the CTA-per-keypoint kernel
    try:
        ref = ume.register_hypotheses(d["src_pts"], d["src_feat"], d["src_kp"], d["tgt_pts"], d["tgt_feat"], d["tgt_kp"], 750, 5.0, want_D=True)
    finally:
        ume.config["cta_moments"] = False
    # (the CTA kernel sums in another order: the distances agree to the Gram form's rounding, the matches wherever
    # the distance gap is clear)
    assert float((out["D"] - ref["D"]).abs().max()) < 2e-3
    same = (out["match"][..., 1] == ref["match"][..., 1]).float().mean()
    assert float(same) > 0.99
