"""ctypes binding of libumereg_b200.so (C ABI: include/umereg_b200.h) and its in-tree build.

There is deliberately no CPU or pure-torch fallback: if the CUDA library is missing or a call
fails, the error is raised to the caller.
"""
import ctypes
import glob
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(_HERE), "include")
LIB_PATH = os.environ.get("UME_LIB_PATH") or os.path.join(CSRC, "libumereg_b200.so")   # override: kernel-variant experiments

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC"]

UME_FLAG_FMA_DIST = 1
UME_FLAG_CELL_DIV2 = 2
UME_FLAG_CTA_MOMENTS = 4
UME_FLAG_RAW_MOMENTS = 8
UME_FLAG_WARP_MOMENTS = 16

_lock = threading.Lock()
_lib = None


def _sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = _sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(INCLUDE, "*.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, jobs=8):
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> csrc/libumereg_b200.so (in-tree, so it
    travels with the repo snapshot).  Cross-compiles without a GPU."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    objs, procs = [], []
    os.makedirs(os.path.join(CSRC, "build"), exist_ok=True)
    for src in _sources():
        obj = os.path.join(CSRC, "build", os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        cmd = [nvcc] + NVCC_FLAGS + ["-I", INCLUDE, "-c", src, "-o", obj]
        if verbose:
            print(" ".join(cmd))
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
        if len(procs) >= jobs:
            _drain(procs)
    _drain(procs)
    cmd = [nvcc, "-shared", "-o", LIB_PATH] + objs
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return LIB_PATH


def _drain(procs):
    while procs:
        src, p = procs.pop(0)
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s" % (src, out.decode(errors="replace")))


def _bind(lib):
    c = ctypes
    vp, i32, i64, f32, u32, sz = c.c_void_p, c.c_int, c.c_int64, c.c_float, c.c_uint, c.c_size_t
    sig = {
        "ume_abi_version": (i32, []),
        "ume_last_error": (c.c_char_p, []),
        "ume_status_string": (c.c_char_p, [i32]),
        "ume_launch_count": (c.c_uint64, []),
        "ume_profile_enable": (None, [i32]),
        "ume_profile_reset": (None, []),
        "ume_profile_read": (i32, [i32, c.POINTER(c.c_double), c.POINTER(c.c_uint64)]),
        "ume_ball_query_workspace_bytes": (sz, [i32, i32, i32, i32]),
        "ume_ball_query_f32": (i32, [vp, vp, i32, i32, i32, i32, f32, u32, vp, vp, vp, vp, vp, sz, vp]),
        "ume_moments_workspace_bytes": (sz, [i32, i32, i32, i32, i32]),
        "ume_moments_f32": (i32, [vp, vp, vp, i32, i32, i32, i32, i32, f32, u32, vp, vp, vp, vp, sz, vp]),
        "ume_moments_pair_f32": (i32, [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, f32, u32, vp, vp, vp, vp, sz, vp]),
        "ume_moments_backward_f32": (i32, [vp, vp, vp, i32, i32, i32, i32, i32, f32, u32, vp, vp, sz, vp]),
        "ume_neighbor_count_f32": (i32, [vp, vp, i32, i32, i32, i32, f32, u32, vp, vp, sz, vp]),
        "ume_orthonormalize_f32": (i32, [vp, i64, i32, vp, vp, vp]),
        "ume_orthonormalize_split_f32": (i32, [vp, i64, i32, vp, vp, vp, vp]),
        "ume_cdist_split_f16": (i32, [vp, vp, i32, i32, i32, i32, vp, vp, vp, vp]),
        "ume_cdist_workspace_bytes": (sz, [i32, i32, i32, i32, i32]),
        "ume_cdist_f32": (i32, [vp, vp, i32, i32, i32, i32, i32, vp, vp, vp, vp, sz, vp]),
        "ume_pair_dist_f32": (i32, [vp, vp, i64, i32, f32, vp, vp]),
        "ume_rigid_solve_f32": (i32, [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp, vp]),
        "ume_rotation_error_deg_f32": (i32, [vp, vp, i64, i32, i32, vp, vp]),
        "ume_rigid_solve_backward_f32": (i32, [vp, vp, vp, i64, i32, vp, vp, vp]),
        "ume_cdist_backward_workspace_bytes": (sz, [i32, i32, i32, i32]),
        "ume_cdist_backward_f32": (i32, [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, vp, vp, vp, sz, vp]),
        "ume_gumbel_topk_f32": (i32, [vp, vp, i32, i32, i32, f32, c.c_uint64, vp, vp]),
        "ume_knn1_workspace_bytes": (sz, [i32, i32, i32]),
        "ume_knn1_gather_f32": (i32, [vp, vp, vp, i32, i32, i32, i32, u32, vp, vp, vp, vp, sz, vp]),
        "ume_knn_workspace_bytes": (sz, [i32, i32, i32]),
        "ume_knn_f32": (i32, [vp, vp, i32, i32, i32, i32, u32, vp, vp, vp, sz, vp]),
        "ume_feature_spatial_var_workspace_bytes": (sz, [i32, i32]),
        "ume_feature_spatial_var_f32": (i32, [vp, vp, i32, i32, i32, i32, u32, vp, vp, sz, vp]),
        "ume_weight_features_f32": (i32, [vp, vp, vp, i64, i32, vp, vp]),
        "ume_linear_sum_assignment_host_f32": (i32, [vp, i32, i32, vp, vp]),
        "ume_voxel_unique_workspace_bytes": (sz, [i32]),
        "ume_voxel_unique_f32": (i32, [vp, i32, f32, vp, vp, vp, vp, sz, vp]),
        "ume_corr_scores_workspace_bytes": (sz, [i32, i32, i32]),
        "ume_corr_scores_f32": (i32, [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, f32, u32, vp, vp, vp, sz, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)          # AttributeError here = header and library out of sync
        fn.restype = res
        fn.argtypes = args
    return sig


EXPORTED_SYMBOLS = ["ume_abi_version", "ume_last_error", "ume_status_string", "ume_launch_count",
                    "ume_profile_enable", "ume_profile_reset", "ume_profile_read",
                    "ume_ball_query_workspace_bytes", "ume_ball_query_f32", "ume_moments_workspace_bytes",
                    "ume_moments_f32", "ume_orthonormalize_f32", "ume_cdist_workspace_bytes", "ume_cdist_f32",
                    "ume_pair_dist_f32", "ume_rigid_solve_f32", "ume_knn1_workspace_bytes",
                    "ume_knn1_gather_f32", "ume_knn_workspace_bytes", "ume_knn_f32",
                    "ume_feature_spatial_var_workspace_bytes", "ume_feature_spatial_var_f32", "ume_weight_features_f32",
                    "ume_corr_scores_workspace_bytes", "ume_corr_scores_f32",
                    "ume_voxel_unique_workspace_bytes", "ume_voxel_unique_f32",
                    "ume_moments_backward_f32", "ume_neighbor_count_f32", "ume_linear_sum_assignment_host_f32",
                    "ume_rotation_error_deg_f32", "ume_gumbel_topk_f32", "ume_orthonormalize_split_f32", "ume_cdist_split_f16",
                    "ume_rigid_solve_backward_f32", "ume_cdist_backward_workspace_bytes", "ume_cdist_backward_f32",
                    "ume_moments_pair_f32"]


def lib():
    """The loaded library.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(
                        "umeregrobust_b200: %s is missing - build it with `python -c 'import __graft_entry__ as g; "
                        "g.build()'` (there is no CPU / torch fallback for this path)" % LIB_PATH)
                handle = ctypes.CDLL(LIB_PATH)
                _bind(handle)
                if handle.ume_abi_version() != 1:
                    raise RuntimeError("umeregrobust_b200: ABI version mismatch")
                _lib = handle
    return _lib


def check(status, what=""):
    if status != 0:
        L = lib()
        msg = L.ume_last_error().decode(errors="replace")
        kind = L.ume_status_string(status).decode()
        if status == -1:
            raise ValueError("%s: %s (%s)" % (what or "umereg_b200", msg, kind))
        raise RuntimeError("%s: %s (%s)" % (what or "umereg_b200", msg, kind))


def launch_count():
    return int(lib().ume_launch_count())


PROF_SLOTS = {"grid": 0, "moments": 1, "ortho": 2, "cdist": 3, "rigid": 4, "ball_query": 5, "knn": 6, "corr": 7}


def profile_enable(on=True):
    lib().ume_profile_enable(1 if on else 0)


def profile_reset():
    lib().ume_profile_reset()


def profile_read():
    """{stage: (total_ms, brackets)} accumulated since the last reset (synchronises the events)."""
    out = {}
    for name, slot in PROF_SLOTS.items():
        ms, n = ctypes.c_double(0.0), ctypes.c_uint64(0)
        check(lib().ume_profile_read(slot, ctypes.byref(ms), ctypes.byref(n)), "profile_read")
        out[name] = (ms.value, int(n.value))
    return out
