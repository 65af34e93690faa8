"""Import the UNMODIFIED reference (/root/reference) on CPU under sys.modules stubs.

Only usable inside the build container (the GPU box has no /root/reference). Used by
make_golden.py to generate the committed golden vectors, and by the optional
`tests/test_oracle_vs_reference.py` (skipped when /root/reference is absent).

The stubs stand in for wheels that are not installable offline (MinkowskiEngine, pytorch3d,
open3d, nksr, pycg); the three pytorch3d ops the hot path calls are supplied by the oracle's
restatement (oracle/pytorch3d_ops.py), everything else is the reference's own code.
"""
import os
import sys
import types
import contextlib

REF_ROOT = os.environ.get("UME_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isfile(os.path.join(REF_ROOT, "evaluate.py"))


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


def import_reference():
    """Returns (evaluate_module, loc_utils_module, eval_utils_module)."""
    global _cached
    if _cached is not None:
        return _cached
    import torch
    here = os.path.dirname(os.path.abspath(__file__))
    repo = os.path.dirname(os.path.dirname(here))
    if repo not in sys.path:
        sys.path.insert(0, repo)
    from oracle import pytorch3d_ops as p3d

    def _t_ball_query(p1, p2, lengths1=None, lengths2=None, K=500, radius=0.2, return_nn=True):
        return p3d.ball_query_torch(p1, p2, K=K, radius=radius, return_nn=return_nn)

    def _t_knn_points(p1, p2, lengths1=None, lengths2=None, norm=2, K=1, version=-1,
                      return_nn=False, return_sorted=True):
        return p3d.knn_points_torch(p1, p2, K=K, return_nn=return_nn)

    def _t_knn_gather(x, idx, lengths=None):
        return p3d.knn_gather_torch(x, idx)

    me = _stub("MinkowskiEngine", MinkowskiNetwork=torch.nn.Module)
    me.MinkowskiFunctional = _stub("MinkowskiEngine.MinkowskiFunctional")
    _stub("pytorch3d")
    _stub("pytorch3d.ops", ball_query=_t_ball_query, knn_points=_t_knn_points,
          knn_gather=_t_knn_gather, sample_farthest_points=None)
    _stub("pytorch3d.structures", Pointclouds=None, padded_to_list=None)
    _stub("open3d")
    _stub("nksr")
    pycg = _stub("pycg")
    pycg.vis = _stub("pycg.vis")
    if "torch.utils.tensorboard" not in sys.modules:
        try:
            import torch.utils.tensorboard  # noqa: F401
        except Exception:
            _stub("torch.utils.tensorboard", SummaryWriter=None)
    sys.path.insert(0, REF_ROOT)
    try:
        with _cwd(REF_ROOT):
            import evaluate as ref_evaluate
            import utils.loc_utils as ref_loc
            import utils.eval_utils as ref_eval
    finally:
        sys.path.remove(REF_ROOT)
    _cached = (ref_evaluate, ref_loc, ref_eval)
    return _cached
