"""ctypes binding of libdsnerf.so (include/dsnerf.h).  The library is the only
compute path: if it is missing or the device is not a B200-class GPU, loading
fails loudly -- there is no CPU or eager fallback."""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DSNERF_LIB") or os.path.join(_HERE, "libdsnerf.so")  # DSNERF_LIB: A/B experiments with a differently built library

SAMPLE_UNIFORM = 0
SAMPLE_GG = 1
MLP_FP32_SIMT = 2
EARLY_STOP = 4
NUM_WEIGHT_TENSORS = 33

ENTRY_POINTS = [
    "dsnerf_abi_version", "dsnerf_create", "dsnerf_destroy", "dsnerf_last_error", "dsnerf_set_weights",
    "dsnerf_set_mesh", "dsnerf_set_frame", "dsnerf_render", "dsnerf_render_train", "dsnerf_render_host", "dsnerf_render_host_async", "dsnerf_wait", "dsnerf_render_gather", "dsnerf_render_z",
    "dsnerf_resample", "dsnerf_composite", "dsnerf_composite_noise", "dsnerf_warp_points", "dsnerf_query_density", "dsnerf_eval_points", "dsnerf_ppts_to_pts", "dsnerf_camera_rays",
    "dsnerf_last_transparent_mask", "dsnerf_tensor_path_active", "dsnerf_mlp_kernel_variant",
    "dsnerf_get_stats", "dsnerf_profile", "dsnerf_profile_read", "dsnerf_debug_tc_timing", "dsnerf_debug_table", "dsnerf_debug_sm_clock", "dsnerf_debug_active",
]


class Stats(ctypes.Structure):
    _fields_ = [
        ("rays", ctypes.c_int64), ("samples", ctypes.c_int64), ("evaluated_samples", ctypes.c_int64),
        ("nn_candidates", ctypes.c_int64), ("algorithmic_flop", ctypes.c_double),
        ("kernel_launches", ctypes.c_int32), ("reserved", ctypes.c_int32),
    ]


class DsnerfError(RuntimeError):
    pass


_lib = None


def load():
    """Load libdsnerf.so and declare every prototype of include/dsnerf.h."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DsnerfError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  There is no fallback path."
        )
    L = ctypes.CDLL(LIB_PATH)
    vp, fp, i64, ci, cu = ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_uint
    L.dsnerf_abi_version.restype = ci
    L.dsnerf_create.argtypes = [ctypes.POINTER(vp), ci]
    L.dsnerf_destroy.argtypes = [vp]
    L.dsnerf_destroy.restype = None
    L.dsnerf_last_error.argtypes = [vp]
    L.dsnerf_last_error.restype = ctypes.c_char_p
    L.dsnerf_set_weights.argtypes = [vp, ctypes.POINTER(ctypes.c_void_p), ci]
    L.dsnerf_set_mesh.argtypes = [vp, vp, ci, fp, ci]
    L.dsnerf_set_frame.argtypes = [vp, fp, fp, ci, ci, fp, fp, fp, vp]
    L.dsnerf_render.argtypes = [vp, fp, fp, fp, fp, i64, ci, cu, fp, fp, fp, fp, fp, fp, vp]
    L.dsnerf_render_host.argtypes = L.dsnerf_render.argtypes
    L.dsnerf_render_host_async.argtypes = L.dsnerf_render.argtypes + [ctypes.POINTER(ctypes.c_int)]
    L.dsnerf_wait.argtypes = [vp, ci]
    L.dsnerf_render_gather.argtypes = [vp, fp, fp, fp, fp, i64, ci, cu, fp, ctypes.POINTER(ctypes.c_void_p), ci, fp, vp]
    L.dsnerf_render_train.argtypes = [vp, fp, fp, fp, fp, i64, ci, cu, fp, fp, fp, fp, fp, fp, fp, fp, vp]
    L.dsnerf_render_z.argtypes = [vp, fp, fp, fp, i64, ci, cu, fp, fp, fp, fp, fp, vp]
    L.dsnerf_resample.argtypes = [vp, fp, fp, i64, ci, ci, fp, vp]
    L.dsnerf_composite.argtypes = [vp, fp, fp, fp, i64, ci, fp, fp, fp, fp, fp, vp]
    L.dsnerf_composite_noise.argtypes = [vp, fp, fp, fp, fp, i64, ci, fp, fp, fp, fp, fp, vp]
    L.dsnerf_warp_points.argtypes = [vp, fp, i64, fp, vp, vp, vp]
    L.dsnerf_query_density.argtypes = [vp, fp, vp, i64, fp, cu, vp]
    L.dsnerf_eval_points.argtypes = [vp, fp, fp, fp, i64, fp, fp, cu, vp]
    L.dsnerf_ppts_to_pts.argtypes = [vp, fp, fp, fp, i64, fp, vp]
    L.dsnerf_camera_rays.argtypes = [vp, ci, ci, vp, vp, vp, fp, fp, fp, fp, fp, vp, vp]
    L.dsnerf_last_transparent_mask.argtypes = [vp, i64, ci, vp, vp]
    L.dsnerf_tensor_path_active.argtypes = [vp]
    L.dsnerf_mlp_kernel_variant.argtypes = [vp]
    L.dsnerf_get_stats.argtypes = [vp, ctypes.POINTER(Stats)]
    L.dsnerf_profile.argtypes = [vp, ci]
    L.dsnerf_profile_read.argtypes = [vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int64), ci]
    L.dsnerf_debug_tc_timing.argtypes = [vp, ctypes.POINTER(ctypes.c_longlong)]
    L.dsnerf_debug_table.argtypes = [vp, ci, ctypes.POINTER(ctypes.c_int)]
    L.dsnerf_debug_sm_clock.argtypes = [vp, vp, vp]
    L.dsnerf_debug_active.argtypes = [vp, i64, vp, vp, ctypes.POINTER(ctypes.c_int64)]
    for name in ENTRY_POINTS:
        fn = getattr(L, name)
        if name not in ("dsnerf_destroy", "dsnerf_last_error"):
            fn.restype = ci
    if L.dsnerf_abi_version() != 1:
        raise DsnerfError("libdsnerf.so ABI version mismatch")
    _lib = L
    return L


class Context:
    """One dsnerf_ctx (one GPU).  Thin, typed wrapper; tensors cross as raw pointers."""

    def __init__(self, device=0):
        self.L = load()
        h = ctypes.c_void_p()
        rc = self.L.dsnerf_create(ctypes.byref(h), int(device))
        if rc != 0:
            raise DsnerfError(
                f"dsnerf_create(device={device}) failed with {rc}: a CUDA sm_100 device is required (no fallback)"
            )
        self.h = h
        self.device = int(device)

    def close(self):
        if getattr(self, "h", None):
            self.L.dsnerf_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc):
        if rc != 0:
            msg = self.L.dsnerf_last_error(self.h)
            raise DsnerfError(f"libdsnerf error {rc}: {msg.decode() if msg else ''}")

    def stats(self):
        s = Stats()
        self.check(self.L.dsnerf_get_stats(self.h, ctypes.byref(s)))
        d = {k: getattr(s, k) for k, _ in Stats._fields_ if k != "reserved"}
        d["searched_samples"] = s.reserved  # only counted with profile bit 2
        return d

    def profile(self, enable):
        self.check(self.L.dsnerf_profile(self.h, int(enable)))

    def profile_read(self, reset=True):
        ms, n = ctypes.c_double(), ctypes.c_int64()
        self.check(self.L.dsnerf_profile_read(self.h, ctypes.byref(ms), ctypes.byref(n), int(reset)))
        return ms.value, n.value
