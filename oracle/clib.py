"""ctypes loader for oracle/libdsoracle.so (oracle/geom.c).  Test infrastructure."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "libdsoracle.so")
    src = os.path.join(_HERE, "geom.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libdsoracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        fp = ctypes.POINTER(ctypes.c_float)
        ip = ctypes.POINTER(ctypes.c_int32)
        L.dso_linspace01.argtypes = [ctypes.c_int, fp]
        L.dso_centroids.argtypes = [fp, ip, ctypes.c_int, fp]
        L.dso_nearest.argtypes = [fp, ctypes.c_int64, fp, ctypes.c_int, ip, fp]
        L.dso_gg_bounds.argtypes = [fp, fp, ctypes.c_int64, fp, ctypes.c_int, ctypes.c_float, fp, fp, fp, fp]
        L.dso_norm3.argtypes = [fp, ctypes.c_int64, fp]
        L.dso_project.argtypes = [fp, fp, ctypes.c_int64, fp, fp]
        L.dso_map2can.argtypes = [fp, fp, fp, ctypes.c_int64, fp]
        L.dso_num_threads.restype = ctypes.c_int
        _LIB = L
    return _LIB


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))


def linspace01(n):
    out = np.empty(n, dtype=np.float32)
    lib().dso_linspace01(n, out.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
    return out


def centroids(verts, faces):
    v, vp = _f(verts)
    f, fp_ = _i(faces)
    out = np.empty((f.shape[0], 3), dtype=np.float32)
    lib().dso_centroids(vp, fp_, f.shape[0], out.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
    return out


def nearest(pts, cent, want_d2=False):
    p, pp = _f(pts.reshape(-1, 3))
    c, cp = _f(cent.reshape(-1, 3))
    idx = np.empty(p.shape[0], dtype=np.int32)
    d2 = np.empty(p.shape[0], dtype=np.float32) if want_d2 else None
    lib().dso_nearest(
        pp, p.shape[0], cp, c.shape[0], idx.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
        d2.ctypes.data_as(ctypes.POINTER(ctypes.c_float)) if want_d2 else None,
    )
    return (idx, d2) if want_d2 else idx


def gg_bounds(o0, ray_d, xyz, near, far, gamma=0.05):
    o, op = _f(o0)
    d, dp = _f(ray_d.reshape(-1, 3))
    x, xp = _f(xyz.reshape(-1, 3))
    n, np_ = _f(near.reshape(-1))
    f, fp_ = _f(far.reshape(-1))
    no = np.empty_like(n)
    fo = np.empty_like(f)
    gamma2 = np.float32(gamma ** 2)  # python double 0.05**2 rounded to fp32, as torch does for `tmp < gamma**2`
    lib().dso_gg_bounds(
        op, dp, d.shape[0], xp, x.shape[0], gamma2, np_, fp_,
        no.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), fo.ctypes.data_as(ctypes.POINTER(ctypes.c_float)),
    )
    return no, fo


def norm3(x):
    a, ap = _f(x.reshape(-1, 3))
    out = np.empty(a.shape[0], dtype=np.float32)
    lib().dso_norm3(ap, a.shape[0], out.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
    return out


def project(pts, tri):
    p, pp = _f(pts.reshape(-1, 3))
    t, tp = _f(tri.reshape(-1, 9))
    uv = np.empty((p.shape[0], 2), dtype=np.float32)
    h = np.empty(p.shape[0], dtype=np.float32)
    fp = ctypes.POINTER(ctypes.c_float)
    lib().dso_project(pp, tp, p.shape[0], uv.ctypes.data_as(fp), h.ctypes.data_as(fp))
    return uv, h


def map2can(uv, h, tri):
    u, up = _f(uv.reshape(-1, 2))
    hh, hp = _f(h.reshape(-1))
    t, tp = _f(tri.reshape(-1, 9))
    out = np.empty((u.shape[0], 3), dtype=np.float32)
    lib().dso_map2can(up, hp, tp, u.shape[0], out.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
    return out
