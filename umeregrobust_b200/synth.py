"""Synthetic KITTI-/nuScenes-shape registration pairs.

Data generation only — nothing here is on the measured path.  A pair mirrors what
`batch_collate_fn_dset` hands the reference's eval loop (datasets/kitti/kitti_dataset.py:546-616):
dense (N,3) float32 clouds whose ROW ORDER IS A RANDOM PERMUTATION, per-point L2-normalised
features like the backbone's output (models.py:612-616), keypoints drawn as random rows
(evaluate.py:199-204) and a ground-truth 4x4 with tgt = R src + t (kitti_dataset.py:437).

Clouds come from a spinning-LiDAR model (64 or 32 beams) ray-cast onto a ground plane plus random
axis-aligned boxes, accumulated over a few jittered sensor poses, de-duplicated on a 0.3 m voxel
grid, trimmed to exactly N rows and permuted.
"""
import numpy as np

KITTI = dict(beams=64, elev_deg=(-24.8, 2.0), az_steps=1875, sensor_h=1.73, max_range=80.0,
             fill_range=57.0, fill_exp=0.55)
NUSCENES = dict(beams=32, elev_deg=(-30.0, 10.0), az_steps=1090, sensor_h=1.84, max_range=70.0,
                fill_range=44.0, fill_exp=0.85)


def random_rotation(rng, max_tilt_deg=3.0, yaw_deg=None):
    """Yaw ~ U(-180,180) about z with a small random tilt."""
    yaw = np.deg2rad(rng.uniform(-180.0, 180.0) if yaw_deg is None else yaw_deg)
    tilt = np.deg2rad(rng.uniform(0.0, max_tilt_deg))
    axis_ang = rng.uniform(0.0, 2 * np.pi)
    ax = np.array([np.cos(axis_ang), np.sin(axis_ang), 0.0])
    Kx = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    Rt = np.eye(3) + np.sin(tilt) * Kx + (1 - np.cos(tilt)) * (Kx @ Kx)
    c, s = np.cos(yaw), np.sin(yaw)
    Rz = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1.0]])
    return Rt @ Rz


def random_rigid(rng, t_range=(4.0, 20.0), **kw):
    R = random_rotation(rng, **kw)
    d = rng.normal(size=3)
    d[2] *= 0.1
    d /= np.linalg.norm(d)
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = d * rng.uniform(*t_range)
    return T


class Scene:
    """Ground plane z = -sensor_h plus random axis-aligned boxes within the sensor's range."""

    def __init__(self, rng, n_boxes=40, extent=60.0):
        c = rng.uniform(-extent, extent, size=(n_boxes, 2))
        keep = np.linalg.norm(c, axis=1) > 4.0               # keep the ego position free
        c = c[keep]
        half = np.stack([rng.uniform(1.0, 6.0, len(c)), rng.uniform(1.0, 6.0, len(c))], 1)
        height = rng.uniform(1.5, 8.0, len(c))
        self.lo = np.concatenate([c - half, np.full((len(c), 1), -10.0)], 1)
        self.hi = np.concatenate([c + half, height[:, None] - 1.73], 1)

    def cast(self, origin, dirs, ground_z, max_range):
        """Nearest hit distance along unit `dirs` from `origin` (inf when nothing within range)."""
        with np.errstate(divide="ignore", invalid="ignore"):
            t_ground = (ground_z - origin[2]) / dirs[:, 2]
            t_ground = np.where((dirs[:, 2] < 0) & (t_ground > 0), t_ground, np.inf)
            inv = 1.0 / dirs                                   # (R,3)
            t0 = (self.lo[None] - origin[None, None]) * inv[:, None]   # (R,nb,3)
            t1 = (self.hi[None] - origin[None, None]) * inv[:, None]
        tmin = np.minimum(t0, t1).max(axis=2)
        tmax = np.maximum(t0, t1).min(axis=2)
        hit = (tmax >= np.maximum(tmin, 0.0))
        t_box = np.where(hit, np.maximum(tmin, 0.0), np.inf).min(axis=1)
        t = np.minimum(t_ground, t_box)
        return np.where(t <= max_range, t, np.inf)


def _sweep(scene, rng, model, origin):
    el = np.deg2rad(np.linspace(model["elev_deg"][0], model["elev_deg"][1], model["beams"]))
    az = np.linspace(0, 2 * np.pi, model["az_steps"], endpoint=False) + rng.uniform(0, 2 * np.pi)
    el_g, az_g = np.meshgrid(el, az, indexing="ij")
    el_g = el_g + rng.normal(scale=np.deg2rad(0.05), size=el_g.shape)
    az_g = az_g + rng.normal(scale=np.deg2rad(0.02), size=az_g.shape)
    dirs = np.stack([np.cos(el_g) * np.cos(az_g), np.cos(el_g) * np.sin(az_g), np.sin(el_g)], -1)
    dirs = dirs.reshape(-1, 3)
    out = []
    for s in range(0, len(dirs), 32768):                       # bound the (R,nb,3) temporaries
        d = dirs[s:s + 32768]
        t = scene.cast(origin, d, -model["sensor_h"], model["max_range"])
        ok = np.isfinite(t)
        t = t[ok] + rng.normal(scale=0.02, size=int(ok.sum()))
        out.append(origin[None] + d[ok] * t[:, None])
    return np.concatenate(out, 0)


def voxel_dedupe(pts, voxel=0.3):
    key = np.floor(pts / voxel).astype(np.int64)
    key = (key[:, 0] + 4096) * (8192 * 8192) + (key[:, 1] + 4096) * 8192 + (key[:, 2] + 4096)
    _, first = np.unique(key, return_index=True)
    return pts[np.sort(first)]


def _surface_samples(scene, rng, model, n):
    """Dense samples of the scene's surfaces (stand-in for the reference's surface-completed SEM
    clouds, which are far denser than a raw sweep): ground disc (denser near the sensor) plus the
    vertical faces and tops of the boxes, 2 cm noise."""
    R = model["fill_range"]
    n_g = int(n * 0.8)
    r = R * rng.uniform(0, 1, n_g) ** model["fill_exp"]
    a = rng.uniform(0, 2 * np.pi, n_g)
    ground = np.stack([r * np.cos(a), r * np.sin(a), np.full(n_g, -model["sensor_h"])], 1)
    inside = ((ground[:, None, :2] > scene.lo[None, :, :2]) & (ground[:, None, :2] < scene.hi[None, :, :2])).all(-1).any(-1) \
        if len(scene.lo) else np.zeros(n_g, bool)
    ground = ground[~inside]
    n_b = n - n_g
    nb = len(scene.lo)
    which = rng.integers(0, nb, n_b)
    lo, hi = scene.lo[which].copy(), scene.hi[which]
    lo[:, 2] = -model["sensor_h"]
    u = rng.uniform(size=(n_b, 3))
    p = lo + u * (hi - lo)
    face = rng.integers(0, 5, n_b)                              # 4 walls + roof
    for f, (ax, side) in enumerate([(0, 0), (0, 1), (1, 0), (1, 1), (2, 1)]):
        m = face == f
        p[m, ax] = (hi if side else lo)[m, ax]
    p = p[(np.linalg.norm(p[:, :2], axis=1) < R) & (p[:, 2] < 3.0)]
    out = np.concatenate([ground, p], 0)
    return out + rng.normal(scale=0.02, size=out.shape)


def lidar_cloud(scene, rng, N, model=KITTI, voxel=0.3, n_sweeps=1):
    """Exactly N rows, float32, randomly permuted: `n_sweeps` ray-cast sweeps plus surface
    completion samples, de-duplicated on a `voxel` grid until N voxels are filled."""
    pts = np.zeros((0, 3))
    for s in range(n_sweeps):
        origin = np.array([rng.uniform(-1.5, 1.5), rng.uniform(-1.5, 1.5), 0.0]) * (s > 0)
        pts = voxel_dedupe(np.concatenate([pts, _sweep(scene, rng, model, origin)], 0), voxel)
    for _ in range(12):
        if len(pts) >= N:
            break
        pts = voxel_dedupe(np.concatenate([pts, _surface_samples(scene, rng, model, N)], 0), voxel)
    if len(pts) < N:                                          # top up with jittered duplicates
        extra = pts[rng.integers(0, len(pts), N - len(pts))] + rng.normal(scale=0.05, size=(N - len(pts), 3))
        pts = np.concatenate([pts, extra], 0)
    sel = rng.permutation(len(pts))[:N]
    return pts[sel].astype(np.float32)


def disc_cloud(rng, N, r_max=50.0):
    """The survey's simpler fallback: radius ~ U(0,r_max) (density ~ 1/r), z ~ U(-2,2)."""
    r = rng.uniform(0, r_max, N)
    a = rng.uniform(0, 2 * np.pi, N)
    return np.stack([r * np.cos(a), r * np.sin(a), rng.uniform(-2, 2, N)], 1).astype(np.float32)


def _normalize_rows(x):
    return x / np.maximum(np.linalg.norm(x, axis=-1, keepdims=True), 1e-12)


def _corresponding_rows(src_kp, tgt, gt, rng):
    """Rows of `tgt` nearest to the images of the source keypoints under gt, in shuffled order: the same
    physical locations seen in the other cloud (an independent sampling of the same surfaces)."""
    from scipy.spatial import cKDTree
    q = src_kp.astype(np.float64) @ gt[:3, :3].T + gt[:3, 3]
    _, kt = cKDTree(tgt).query(q, k=1, workers=-1)
    return kt[rng.permutation(len(kt))]


def make_pair(seed, N=120000, C=32, n_kp=1024, model=KITTI, gt=None, exact_copy=False,
              feat_noise=0.05, generator="lidar", feat_model="iid", kp_mode="random", field_scale=0.3,
              field_noise=0.2):
    """One registration pair.  Returns dict of float32 arrays:
    src_pts (N,3), src_feat (N,C), src_kp (n,3), src_kp_idx (n,), tgt_* likewise, gt (4,4).

    exact_copy=True: the target is the SAME points moved by gt with identical features and the
    same keypoint rows (the exact-recovery known-answer case, BASELINE config #1).
    feat_model: "iid" = normalize(randn) per point, carried to the target by the nearest source point
    (descriptors of DIFFERENT samplings of a surface are then unrelated, so only timing and
    kernel-vs-oracle parity are meaningful); "field" = a smooth random field normalize(sin(W x + phi) +
    noise) evaluated at every point of both clouds in the scene frame — what a backbone's output looks
    like: two samplings of the same neighbourhood give nearly the same UME matrix, so matches and
    per-match transforms are meaningful against the ground truth (BASELINE config #5's check).
    kp_mode: "random" rows of each cloud (evaluate.py:199-204) or "corresponding" (the target keypoints
    are the target rows nearest to the moved source keypoints)."""
    from scipy.spatial import cKDTree
    rng = np.random.default_rng(seed)
    scene = Scene(rng) if generator == "lidar" else None
    src = lidar_cloud(scene, rng, N, model) if scene is not None else disc_cloud(rng, N)
    if gt is None:
        gt = random_rigid(rng)
    gt = np.asarray(gt, dtype=np.float64)
    R, t = gt[:3, :3], gt[:3, 3]
    if feat_model == "field":
        W, phi = rng.normal(size=(3, C)) * field_scale, rng.uniform(0, 2 * np.pi, C)

        def field(x):
            return _normalize_rows(np.sin(x.astype(np.float64) @ W + phi) + field_noise * rng.normal(size=(len(x), C))).astype(np.float32)
        src_feat = field(src)
    else:
        src_feat = _normalize_rows(rng.normal(size=(N, C))).astype(np.float32)
    if exact_copy:
        tgt = (src.astype(np.float64) @ R.T + t).astype(np.float32)
        tgt_feat = src_feat.copy()
        kp_idx_s = rng.choice(N, n_kp, replace=False)
        kp_idx_t = kp_idx_s.copy()
    else:
        tgt_local = lidar_cloud(scene, rng, N, model) if scene is not None else disc_cloud(rng, N)
        if feat_model == "field":
            tgt_feat = field(tgt_local)
        else:
            _, nn = cKDTree(src).query(tgt_local, k=1, workers=-1)
            tgt_feat = _normalize_rows(src_feat[nn] + rng.normal(scale=feat_noise, size=(N, C))).astype(np.float32)
        tgt = (tgt_local.astype(np.float64) @ R.T + t).astype(np.float32)
        kp_idx_s = rng.choice(N, n_kp, replace=False)
        kp_idx_t = _corresponding_rows(src[kp_idx_s], tgt, gt, rng) if kp_mode == "corresponding" else rng.choice(N, n_kp, replace=False)
    return dict(src_pts=src, src_feat=src_feat, src_kp=src[kp_idx_s].copy(), src_kp_idx=kp_idx_s,
                tgt_pts=tgt, tgt_feat=tgt_feat, tgt_kp=tgt[kp_idx_t].copy(), tgt_kp_idx=kp_idx_t,
                gt=gt.astype(np.float32))


def rederive_pair(base, seed, n_kp=None, gt_extra=None, kp_mode="random"):
    """A cheap new pair from a generated one: fresh row permutations, fresh keypoints and an extra
    rigid motion of the target (features ride along with their rows).  Used to fill large batches
    without re-running the ray caster for every pair."""
    rng = np.random.default_rng(seed)
    N = base["src_pts"].shape[0]
    n_kp = n_kp or base["src_kp"].shape[0]
    ps, pt = rng.permutation(N), rng.permutation(N)
    extra = random_rigid(rng) if gt_extra is None else np.asarray(gt_extra, dtype=np.float64)
    gt = extra @ base["gt"].astype(np.float64)
    tgt = (base["tgt_pts"][pt].astype(np.float64) @ extra[:3, :3].T + extra[:3, 3]).astype(np.float32)
    src = base["src_pts"][ps]
    ks = rng.choice(N, n_kp, replace=False)
    kt = _corresponding_rows(src[ks], tgt, gt, rng) if kp_mode == "corresponding" else rng.choice(N, n_kp, replace=False)
    return dict(src_pts=src, src_feat=base["src_feat"][ps], src_kp=src[ks].copy(), src_kp_idx=ks,
                tgt_pts=tgt, tgt_feat=base["tgt_feat"][pt], tgt_kp=tgt[kt].copy(), tgt_kp_idx=kt,
                gt=gt.astype(np.float32))


def make_batch(n_pairs, seed0=0, n_base=4, **kw):
    """Stack `n_pairs` pairs into (B, ...) arrays; `n_base` distinct scenes are ray-cast, the rest
    are re-derived from them."""
    bases = [make_pair(seed0 + i, **kw) for i in range(min(n_base, n_pairs))]
    pairs = []
    for p in range(n_pairs):
        pairs.append(bases[p] if p < len(bases) else
                     rederive_pair(bases[p % len(bases)], seed0 + 1000 + p, n_kp=kw.get("n_kp"),
                                   kp_mode=kw.get("kp_mode", "random")))
    return {k: np.stack([q[k] for q in pairs], 0) for k in pairs[0]}
