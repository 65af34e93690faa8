"""Drop-in: rebind the names the reference's eval loop resolves at call time.

The reference has no plugin interface; `evaluate.py` looks up `ball_query`, `knn_points`,
`knn_gather`, `my_ume_generation`, `ume_cdist`, `batch_estimate_transform_ume_old`,
`ume_kp_layer` and `FeatureCorrelator` as ITS OWN module globals (evaluate.py:5,15,50), so
replacing those globals is the whole integration of the inference path.

Arithmetic: the reference's eval loop runs pytorch3d's CUDA `ball_query`, whose `dist2 += diff *
diff` nvcc contracts into fused multiply-adds; `patch_reference()` therefore switches the distance
test of this library to the fused form (`config["fma_dist"] = True`, bit-identical to the oracle's
`fma=True` restatement).  Pass `fma_dist=False` for pytorch3d's CPU arithmetic.

`utils.loc_utils` and `loss` are NOT touched by default: the training losses import
`ume_cdist` / `batch_estimate_transform_ume_old` from `utils.loc_utils` and need autograd, which the
inference kernels do not provide (they raise on inputs that require grad).  `training=True` installs
the differentiable mirrors of `umeregrobust_b200.training` there instead.
"""
import sys

from . import api

_EVAL_NAMES = ("ball_query", "knn_points", "knn_gather", "my_ume_generation", "ume_cdist",
               "batch_estimate_transform_ume_old", "ume_kp_layer",
               # hypothesis selection
               "FeatureCorrelator")
_LOC_UTILS_NAMES = ("ball_query", "knn_points", "knn_gather", "ume_cdist", "batch_estimate_transform_ume_old",
                    "ume_kp_layer", "ball_query_gather", "FeatureCorrelator", "feature_spatial_var", "cauchy_kernel",
                    "pc_corr_cost_pytorch3d")


def _rebind(module, names, source, done):
    for name in names:
        if hasattr(module, name) and hasattr(source, name):
            setattr(module, name, getattr(source, name))
            done.append((module.__name__, name))


def patch_reference(evaluate_module=None, loc_utils_module=None, fma_dist=True, patch_loc_utils=False, training=False,
                    loss_module=None):
    """Rebinds the hot-path names inside the (already imported) reference modules and returns the
    list of (module, name) pairs that were replaced.

      evaluate_module  the reference's `evaluate` (default: sys.modules["evaluate"]) — always patched;
      patch_loc_utils  also rebind the INFERENCE kernels inside `utils.loc_utils` (for callers that
                       use `ume_kp_layer` / `FeatureCorrelator` from there without autograd);
      training         install the differentiable mirrors (`umeregrobust_b200.training`) into
                       `utils.loc_utils` and `loss` (`generate_ume_from_keypoints2`,
                       `UMEContrastiveLoss`, `CubeRegistrationLoss`; `ume_cdist` and
                       `batch_estimate_transform_ume_old` keep torch's autograd there);
      fma_dist         distance test as pytorch3d's CUDA build evaluates it (default) or as its CPU
                       build does."""
    ev = evaluate_module or sys.modules.get("evaluate")
    lu = loc_utils_module or sys.modules.get("utils.loc_utils")
    if ev is None and not ((patch_loc_utils or training) and lu is not None):
        raise RuntimeError("patch_reference: import the reference's `evaluate` (or pass the module) first")
    if patch_loc_utils and training:
        raise ValueError("patch_reference: patch_loc_utils installs kernels without autograd into utils.loc_utils; "
                         "it cannot be combined with training=True")
    api.config["fma_dist"] = bool(fma_dist)
    done = []
    if ev is not None:
        _rebind(ev, _EVAL_NAMES, api, done)
    if patch_loc_utils:
        if lu is None:
            raise RuntimeError("patch_reference: utils.loc_utils is not imported")
        _rebind(lu, _LOC_UTILS_NAMES, api, done)
    if training:
        from . import training as tr
        if lu is not None:
            _rebind(lu, ("generate_ume_from_keypoints2",), tr, done)
        ls = loss_module or sys.modules.get("loss")
        if ls is not None:
            _rebind(ls, ("generate_ume_from_keypoints2", "UMEContrastiveLoss", "CubeRegistrationLoss"), tr, done)
    return done
