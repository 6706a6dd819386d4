"""Shared parity criteria for CPU and GPU tests."""
from __future__ import annotations

import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# north_star tolerance: 1e-4 max-abs on RGB and depth.
TOL = 1e-4
# The reference's shading normal is d(density)/d(xyz) through 7x256 ReLUs
# (model/spacenet.py:251): it is DISCONTINUOUS wherever a pre-activation crosses
# zero.  A sample whose smallest relative pre-activation margin is below
# KINK_MARGIN is within fp32 rounding distance of such a jump, so any two
# correct fp32 implementations (even the reference on CPU vs GPU) may disagree on
# its normal by O(0.1) and on its colour by O(1e-3).  Rays containing such a
# sample are exempt from the 1e-4 RGB bound, must stay within KINK_RGB_BOUND, and
# the number of rays above 1e-4 is capped at max(3, 0.2 %) and recorded (measured:
# 2 of 12 288 rays of the 512x512x64 frame, worst 2.2e-4; profiles/r02_parity.json).  Depth/acc never depend on the normal
# and are held to 1e-4 everywhere.
KINK_MARGIN = 2e-6
KINK_RGB_BOUND = 5e-3


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def bits_equal(a, b):
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32).reshape(a.shape)
    return int(np.sum(a.view(np.int32) != b.view(np.int32)))


def record(name, stats):
    """Append the achieved error statistics of a parity test to gpurun_out/r02_parity.json (copied to profiles/ by hand
    after a GPU run: the numbers a test merely prints are lost)."""
    import json

    try:
        import torch

        if not torch.cuda.is_available():  # CPU runs pin the oracle; only the GPU path's achieved errors are recorded
            return
    except ImportError:
        return
    root = os.environ.get("GRAFT_REPO_ROOT") or os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    d = os.path.join(root, "gpurun_out")
    try:
        os.makedirs(d, exist_ok=True)
        path = os.path.join(d, "r02_parity.json")
        data = {}
        if os.path.exists(path):
            with open(path) as f:
                data = json.load(f)
        data[name] = stats
        with open(path, "w") as f:
            json.dump(data, f, indent=1, sort_keys=True)
    except OSError:
        pass


def check_rays(out, ref, kink_ray=None, tol=TOL, what="", strict=False):
    """Assert the per-ray outputs match; returns a dict of error statistics.  ``strict``: every ray within ``tol`` on rgb, no
    ReLU-kink exemption (used wherever the committed golden is known to be met on every ray)."""
    if strict:
        n_kink = int(np.sum(kink_ray)) if kink_ray is not None else 0
        kink_ray = None
    col = np.abs(np.asarray(out["color"]) - np.asarray(ref["color"])).reshape(len(ref["color"]), -1).max(1)
    dep = np.abs(np.asarray(out["depth_map"]).ravel() - np.asarray(ref["depth_map"]).ravel())
    acc = np.abs(np.asarray(out["acc_map"]).ravel() - np.asarray(ref["acc_map"]).ravel())
    assert dep.max() <= tol, f"{what} depth err {dep.max():.3e}"
    assert acc.max() <= tol, f"{what} acc err {acc.max():.3e}"
    bad = col > tol
    stats = {"rgb_max": float(col.max()), "depth_max": float(dep.max()), "acc_max": float(acc.max()),
             "rays_over_tol": int(bad.sum())}
    if kink_ray is None:
        assert not bad.any(), f"{what} rgb err {col.max():.3e} on {bad.sum()} rays"
    else:
        unexplained = bad & ~kink_ray
        assert not unexplained.any(), f"{what} rgb err {col[unexplained].max():.3e} on {unexplained.sum()} rays with no ReLU kink"
        assert col.max() <= KINK_RGB_BOUND, f"{what} rgb err {col.max():.3e} exceeds the kink bound"
        assert bad.sum() <= max(3, 0.002 * len(col)), f"{what} too many kink rays over tolerance: {bad.sum()}"
        stats["rgb_max_nokink"] = float(col[~kink_ray].max()) if (~kink_ray).any() else 0.0
    if strict:
        stats["kink_rays"] = n_kink
        stats["strict"] = True
    elif kink_ray is not None:
        stats["kink_rays"] = int(np.sum(kink_ray))
    d0 = np.asarray(ref["disp_map"]).ravel()
    d1 = np.asarray(out["disp_map"]).ravel()
    assert np.array_equal(np.isnan(d0), np.isnan(d1)), f"{what} disp NaN pattern (acc==0 rays) differs"
    if what:
        record(what, stats)
    return stats
