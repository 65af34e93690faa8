"""Drop-in: rebind the names the reference's eval loop resolves at call time.

The reference has no plugin interface; `evaluate.py` looks up `ball_query`, `knn_points`,
`knn_gather`, `my_ume_generation`, `ume_cdist`, `batch_estimate_transform_ume_old` and
`ume_kp_layer` as module globals (evaluate.py:5,15,50; utils/loc_utils.py:4,8,292,357), so
replacing those globals is the whole integration.
"""
import sys

from . import api

_NAMES = ("ball_query", "knn_points", "knn_gather", "my_ume_generation", "ume_cdist",
          "batch_estimate_transform_ume_old", "ume_kp_layer", "ball_query_gather",
          # hypothesis selection
          "FeatureCorrelator", "feature_spatial_var", "cauchy_kernel", "pc_corr_cost_pytorch3d")


def patch_reference(evaluate_module=None, loc_utils_module=None):
    """Rebinds the hot-path names inside the (already imported) reference modules.  Returns the
    list of (module, name) pairs that were replaced."""
    mods = []
    ev = evaluate_module or sys.modules.get("evaluate")
    lu = loc_utils_module or sys.modules.get("utils.loc_utils")
    if ev is not None:
        mods.append(ev)
    if lu is not None:
        mods.append(lu)
    if not mods:
        raise RuntimeError("patch_reference: import the reference's `evaluate` / `utils.loc_utils` first")
    done = []
    for m in mods:
        for name in _NAMES:
            if hasattr(m, name):
                setattr(m, name, getattr(api, name))
                done.append((m.__name__, name))
    return done
