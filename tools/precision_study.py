#!/usr/bin/env python
"""CPU precision study (oracle emulation, no GPU): max |d rgb| / |d depth| of reduced-precision GEMM operand schemes against
the fp32 oracle on 512 rays x 64 samples of the 128x128 scene.  Used to decide how many tensor-core MMAs per forward k-step
the 1e-4 parity bound allows.  Run:  python tools/precision_study.py [scheme ...]"""
import functools
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import common as C  # noqa: E402
from dual_space_nerf_b200 import net as N  # noqa: E402
from dual_space_nerf_b200 import scene as S  # noqa: E402
from oracle import oracle as O  # noqa: E402

f32 = np.float32


def h16(x):
    return x.astype(np.float16).astype(f32)


def e4m3(x):
    """round to nearest fp8 e4m3 (3 mantissa bits, min normal 2^-6, subnormal step 2^-9, max 448)"""
    x = np.asarray(x, f32)
    ax = np.abs(x)
    e = np.floor(np.log2(np.maximum(ax, 2.0 ** -6)))
    step = np.exp2(e - 3).astype(f32)
    step = np.where(ax < 2.0 ** -6, f32(2.0 ** -9), step)
    return (np.clip(np.round(x / step) * step, -448, 448)).astype(f32)


def p2scale(m, target=256.0):
    m = np.maximum(m, 1e-30)
    return np.exp2(np.floor(np.log2(target / m))).astype(f32)


def fp8_tensor(x):
    s = p2scale(np.abs(x).max())
    return e4m3(x * s) / s


def fp8_mx(x, axis):
    """MX-style block scaling: one power-of-two scale per 32 consecutive elements along the contraction axis"""
    x = np.moveaxis(x, axis, -1)
    K = x.shape[-1]
    pad = (-K) % 32
    xp = np.pad(x, [(0, 0)] * (x.ndim - 1) + [(0, pad)])
    b = xp.reshape(xp.shape[:-1] + (-1, 32))
    s = p2scale(np.abs(b).max(-1, keepdims=True))
    q = (e4m3(b * s) / s).reshape(xp.shape)[..., :K]
    return np.moveaxis(q, -1, axis)


FWD = lambda name: not name.startswith("b.") and name not in ("rgb.0", "rgb.1", "dens")


def make_rounder(scheme):
    def r(name, x, w):
        x = x.astype(f32)
        w = w.astype(f32)
        if name.startswith("b.") or name == "rgb.0":       # as shipped: backward chain and rgb head single-pass fp16
            return h16(x) @ h16(w)
        if not FWD(name):
            return x @ w
        xh, wh = h16(x), h16(w)
        xl, wl = x - xh, w - wh
        if scheme == "fp16x3":
            return xh @ wh + xh @ h16(wl) + h16(xl) @ wh
        if scheme == "fp16x1":
            return xh @ wh
        if scheme == "fp8corr_tensor":                      # corrections with both operands in per-tensor-scaled e4m3
            return xh @ wh + fp8_tensor(xh) @ fp8_tensor(wl) + fp8_tensor(xl) @ fp8_tensor(wh)
        if scheme == "fp8corr_mx":                          # corrections in MX-scaled e4m3 (32-element blocks along K)
            return xh @ wh + fp8_mx(xh, 1) @ fp8_mx(wl, 0) + fp8_mx(xl, 1) @ fp8_mx(wh, 0)
        if scheme == "fp8corr_mx_merged":                   # one fp8 GEMM with K doubled: [xh | xl] . [wl ; wh]
            return xh @ wh + np.concatenate([fp8_mx(xh, 1), fp8_mx(xl, 1)], 1) @ np.concatenate([fp8_mx(wl, 0), fp8_mx(wh, 0)], 0)
        if scheme == "hl8_lh16":                            # x_hi.w_lo in MX fp8, x_lo.w_hi in fp16 (2.5 pass-equivalents)
            return xh @ wh + fp8_mx(xh, 1) @ fp8_mx(wl, 0) + h16(xl) @ wh
        if scheme == "hl16_lh8":                            # x_hi.w_lo in fp16, x_lo.w_hi in MX fp8
            return xh @ wh + xh @ h16(wl) + fp8_mx(xl, 1) @ fp8_mx(wh, 0)
        raise SystemExit(f"unknown scheme {scheme}")
    return r


def main():
    schemes = sys.argv[1:] or ["fp16x3", "fp16x1", "fp8corr_tensor", "fp8corr_mx"]
    g = C.golden("render_128x128x64.npz")
    rays = g["rays"]
    sc = S.make_scene(128, 128)
    sd = N.synthetic_net(0).state_dict()
    args = (sc["ray_o"][rays], sc["ray_d"][rays], sc["near"][rays], sc["far"][rays], sc["posed"], sc["poses"], sc["frame"])
    orig = O.spacenet_forward
    st = {}
    ref = O.Oracle(sd, sc["canonical"], sc["faces"], 64).render(*args, Th=sc["Th"], stages=st)
    kink = (st["kink_margin"] < C.KINK_MARGIN).reshape(-1, 64).any(1)
    print(f"{len(rays)} rays, {int((~st['mask']).sum())} evaluated samples, {int(kink.sum())} kink rays (excluded from the rgb column)")
    for s in schemes:
        O.spacenet_forward = functools.partial(orig, rounder=make_rounder(s))
        try:
            out = O.Oracle(sd, sc["canonical"], sc["faces"], 64).render(*args, Th=sc["Th"])
        finally:
            O.spacenet_forward = orig
        drgb = np.abs(out["color"] - ref["color"]).max(1)
        print(f"{s:20s} max|d rgb| {drgb[~kink].max():.2e} (all rays {drgb.max():.2e})  max|d depth| "
              f"{np.abs(out['depth_map'] - ref['depth_map']).max():.2e}  max|d acc| {np.abs(out['acc_map'] - ref['acc_map']).max():.2e}")


if __name__ == "__main__":
    main()
