"""ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Import the UNMODIFIED reference on CPU under sys.modules stubs.

Where the reference comes from, in this order:
  1. $UME_REFERENCE_ROOT,
  2. /root/reference (the build container),
  3. <repo>/baseline/_ref/ — a git-ignored STAGED copy of the handful of reference files the hot
     path lives in (`stage()` below, run by `__graft_entry__.build()` whenever /root/reference is
     present).  It is not part of the repository's history (the reference's sources are never
     committed); it travels to the GPU box with the snapshot exactly like the built `.so` files,
     so that `bench.py --impl reference` can time the reference's OWN functions there.

The stubs stand in for wheels that are not installable offline (MinkowskiEngine, pytorch3d,
open3d, nksr, pycg); the three pytorch3d ops the hot path calls are supplied by the oracle's
restatement (oracle/pytorch3d_ops.py), everything else is the reference's own code.

Only tests/, tests/golden/make_golden*.py and bench.py's cpu_baseline / --impl reference legs may
import this module.
"""
import contextlib
import os
import shutil
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(_HERE)
STAGED_ROOT = os.path.join(REPO, "baseline", "_ref")

# what `import evaluate` + `import utils.loc_utils` + `import loss` touch
STAGED_FILES = [
    "evaluate.py", "models.py", "loss.py",
    "utils/__init__.py", "utils/loc_utils.py", "utils/eval_utils.py", "utils/general_utils.py",
    "datasets/__init__.py", "datasets/kitti/kitti_dataset.py", "datasets/kitti/kitti_config.yaml",
    "datasets/nuscenes/nuscenes_dataset.py",
]


def _has_reference(root):
    return bool(root) and os.path.isfile(os.path.join(root, "evaluate.py")) and \
        os.path.isfile(os.path.join(root, "utils", "loc_utils.py"))


def find_root():
    for root in (os.environ.get("UME_REFERENCE_ROOT"), "/root/reference", STAGED_ROOT):
        if _has_reference(root):
            return root
    return None


REF_ROOT = find_root() or "/root/reference"


def reference_available():
    return _has_reference(REF_ROOT)


def reference_kind():
    """'tree' = the full reference checkout, 'staged' = baseline/_ref, None = absent."""
    if not reference_available():
        return None
    return "staged" if os.path.abspath(REF_ROOT) == os.path.abspath(STAGED_ROOT) else "tree"


def stage(src_root="/root/reference", force=False):
    """Copy STAGED_FILES from the reference checkout into baseline/_ref (git-ignored).  No-op when
    the checkout is absent (the GPU box).  Returns the staged root or None."""
    if not _has_reference(src_root):
        return STAGED_ROOT if _has_reference(STAGED_ROOT) else None
    for rel in STAGED_FILES:
        src = os.path.join(src_root, rel)
        dst = os.path.join(STAGED_ROOT, rel)
        if not os.path.isfile(src):
            continue
        if (not force and os.path.isfile(dst) and os.path.getsize(dst) == os.path.getsize(src)
                and os.path.getmtime(dst) >= os.path.getmtime(src)):
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
    return STAGED_ROOT


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


@contextlib.contextmanager
def _cwd(path):
    old = os.getcwd()
    os.chdir(path)
    try:
        yield
    finally:
        os.chdir(old)


_cached = None


def install_stubs(num_threads=0, fma=False):
    """sys.modules stubs for the wheels the reference imports.  The pytorch3d names are bound to the
    OpenMP C restatement (num_threads = 0: OpenMP's default; fma: nvcc-contracted distance)."""
    import torch
    if REPO not in sys.path:
        sys.path.insert(0, REPO)
    from oracle import pytorch3d_ops as p3d

    def _t_ball_query(p1, p2, lengths1=None, lengths2=None, K=500, radius=0.2, return_nn=True):
        return p3d.ball_query_torch(p1, p2, K=K, radius=radius, return_nn=return_nn, fma=fma, threads=num_threads)

    def _t_knn_points(p1, p2, lengths1=None, lengths2=None, norm=2, K=1, version=-1,
                      return_nn=False, return_sorted=True):
        return p3d.knn_points_torch(p1, p2, K=K, return_nn=return_nn, fma=fma, threads=num_threads)

    def _t_knn_gather(x, idx, lengths=None):
        return p3d.knn_gather_torch(x, idx)

    me = _stub("MinkowskiEngine", MinkowskiNetwork=torch.nn.Module)
    me.MinkowskiFunctional = _stub("MinkowskiEngine.MinkowskiFunctional")
    me.utils = _stub("MinkowskiEngine.utils")
    _stub("pytorch3d")
    _stub("pytorch3d.ops", ball_query=_t_ball_query, knn_points=_t_knn_points,
          knn_gather=_t_knn_gather, sample_farthest_points=None)
    _stub("pytorch3d.structures", Pointclouds=None, padded_to_list=None)
    _stub("open3d")
    _stub("nksr")
    pycg = _stub("pycg")
    pycg.vis = _stub("pycg.vis")
    for name, attrs in (("torch.utils.tensorboard", dict(SummaryWriter=None)), ("tqdm", dict(tqdm=lambda x, *a, **k: x))):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                _stub(name, **attrs)


def import_reference(num_threads=0, fma=False):
    """Returns (evaluate_module, loc_utils_module, eval_utils_module) of the reference."""
    global _cached
    if _cached is not None and _cached[0] == (num_threads, fma):
        return _cached[1]
    if not reference_available():
        raise RuntimeError("reference not found (looked at $UME_REFERENCE_ROOT, /root/reference, %s)" % STAGED_ROOT)
    install_stubs(num_threads=num_threads, fma=fma)
    for name in ("evaluate", "utils", "utils.loc_utils", "utils.eval_utils", "utils.general_utils", "models", "loss",
                 "datasets", "datasets.kitti", "datasets.kitti.kitti_dataset", "datasets.nuscenes",
                 "datasets.nuscenes.nuscenes_dataset"):
        sys.modules.pop(name, None)
    sys.path.insert(0, REF_ROOT)
    try:
        with _cwd(REF_ROOT):
            import evaluate as ref_evaluate
            import utils.loc_utils as ref_loc
            import utils.eval_utils as ref_eval
    finally:
        sys.path.remove(REF_ROOT)
    _cached = ((num_threads, fma), (ref_evaluate, ref_loc, ref_eval))
    return _cached[1]
