"""Generate tests/golden/lbs_ppts_to_pts.npz by running the UNMODIFIED reference function
utils/blend_utils.py:ppts_to_pts (the second definition, :72-81, is the one Python binds) on CPU.

Run HERE (needs /root/reference):  python tests/make_golden_lbs.py
"""
import importlib.util
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def make_inputs(P=768, seed=0):
    """Skinning-like inputs: sparse blend weights over 24 joints, rigid joint transforms."""
    rng = np.random.RandomState(seed)
    bw = np.zeros((24, P), np.float32)
    for p in range(P):
        js = rng.choice(24, 4, replace=False)
        w = rng.rand(4).astype(np.float32) ** 2
        bw[js, p] = w / w.sum()
    A = np.zeros((24, 4, 4), np.float32)
    for j in range(24):
        ax = rng.randn(3)
        ax /= np.linalg.norm(ax)
        ang = rng.uniform(-1.2, 1.2)
        K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
        A[j, :3, :3] = np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * (K @ K)
        A[j, :3, 3] = rng.uniform(-0.5, 0.5, 3)
        A[j, 3, 3] = 1
    pts = rng.uniform(-1, 1, (P, 3)).astype(np.float32)
    return pts, bw, A


def main():
    spec = importlib.util.spec_from_file_location("ref_blend_utils", "/root/reference/utils/blend_utils.py")
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    pts, bw, A = make_inputs()
    out = m.ppts_to_pts(torch.from_numpy(pts)[None], torch.from_numpy(bw)[None], torch.from_numpy(A)[None])[0].numpy()
    np.savez_compressed(os.path.join(HERE, "golden", "lbs_ppts_to_pts.npz"), pts=pts, bw=bw, A=A, out=out)
    print("ppts_to_pts golden:", out.shape, float(np.abs(out).max()))


if __name__ == "__main__":
    main()
