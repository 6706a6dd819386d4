"""Deterministic synthetic scene for parity tests and the benchmark.

The reference ships no data (SMPL model and ZJU-Mocap/H36M are licensed), so
every test and bench input is generated here, following SURVEY.md Appendix A:

* a UV-sphere mesh with exactly SMPL's counts (V = 6890, F = 13776), scaled to
  a body-sized ellipsoid for the canonical ("X-pose") mesh and sheared +
  translated for the posed mesh;
* a pin-hole camera whose rays follow ``utils/rays_utils.py:16-30`` (camera
  centre as origin, unnormalised direction ``R^T K^-1 [i, j, 1]``);
* near/far from a slab test against the padded body bounding box
  (``utils/rays_utils.py:63-97`` semantics) where the ray hits the box and the
  constants (2.6, 3.6) elsewhere, so all H*W rays are rendered;
* axis-angle ``poses`` (24,3), ``frame`` index and ``Th`` translation as the
  reference's ``batch`` dict carries them
  (``dataloader/zju_mocap_dataset.py:160-185``).

Only numpy is used; weights are produced by :mod:`dual_space_nerf_b200.net`.
"""
from __future__ import annotations

import numpy as np

N_RINGS = 82
N_SEG = 84
V = 2 + N_RINGS * N_SEG  # 6890
F = 2 * N_SEG + 2 * (N_RINGS - 1) * N_SEG  # 13776

CANON_SCALE = np.array([0.25, 0.15, 0.85], dtype=np.float32)
POSE_SHIFT = np.array([0.2, -0.1, 1.0], dtype=np.float32)
CAM_CENTRE = np.array([0.2, -3.1, 1.0], dtype=np.float64)
MISS_NEAR_FAR = (2.6, 3.6)
LIGHT_CENTER_313 = np.array([0.219, -0.178, 1.146], dtype=np.float32)  # configs/zju_mocap/313.yml:52


def uv_sphere():
    """Unit UV sphere: (V,3) float32 vertices and (F,3) int64 faces (outward winding)."""
    theta = np.linspace(0.0, np.pi, N_SEG)[1:-1]  # 82 rings
    phi = np.linspace(0.0, 2.0 * np.pi, N_SEG, endpoint=False)
    assert theta.shape[0] == N_RINGS
    verts = np.zeros((V, 3), dtype=np.float64)
    verts[0] = (0.0, 0.0, 1.0)
    st, ct = np.sin(theta)[:, None], np.cos(theta)[:, None]
    ring = np.stack(
        [st * np.cos(phi)[None], st * np.sin(phi)[None], np.broadcast_to(ct, (N_RINGS, N_SEG))], -1
    )
    verts[1:-1] = ring.reshape(-1, 3)
    verts[-1] = (0.0, 0.0, -1.0)

    faces = []
    s = np.arange(N_SEG)
    sn = (s + 1) % N_SEG
    faces.append(np.stack([np.zeros_like(s), 1 + s, 1 + sn], -1))
    for r in range(N_RINGS - 1):
        a = 1 + N_SEG * r + s
        b = 1 + N_SEG * r + sn
        c = a + N_SEG
        d = b + N_SEG
        quad = np.stack([np.stack([a, c, b], -1), np.stack([b, c, d], -1)], 1).reshape(-1, 3)
        faces.append(quad)
    last = 1 + (N_RINGS - 1) * N_SEG
    faces.append(np.stack([np.full_like(s, V - 1), last + sn, last + s], -1))
    faces = np.concatenate(faces, 0).astype(np.int64)
    assert faces.shape == (F, 3)
    return verts.astype(np.float32), faces


def body_meshes():
    """Canonical and posed vertices (V,3) float32 and faces (F,3) int64."""
    verts, faces = uv_sphere()
    canonical = (verts * CANON_SCALE).astype(np.float32)
    posed = canonical.copy()
    posed[:, 0] = posed[:, 0] + np.float32(0.15) * np.sin(np.float32(2.0) * canonical[:, 2])
    posed = (posed + POSE_SHIFT).astype(np.float32)
    return canonical, posed, faces


def camera(H, W):
    K = np.array([[1.2 * W, 0, W / 2.0], [0, 1.2 * W, H / 2.0], [0, 0, 1.0]], dtype=np.float64)
    R = np.array([[1.0, 0, 0], [0, 0, -1.0], [0, 1.0, 0]], dtype=np.float64)
    T = -R @ CAM_CENTRE
    return K, R, T


def pinhole_rays(H, W, K, R, T):
    """Camera-centre origin and unnormalised directions, (H*W,3) float32 each."""
    origin = -(R.T @ T)
    jj, ii = np.meshgrid(np.arange(H, dtype=np.float64), np.arange(W, dtype=np.float64), indexing="ij")
    pix = np.stack([ii, jj, np.ones_like(ii)], -1).reshape(-1, 3)
    cam = pix @ np.linalg.inv(K).T
    world = (cam - T[None]) @ R
    d = world - origin[None]
    o = np.broadcast_to(origin[None], d.shape)
    return np.ascontiguousarray(o, dtype=np.float32), np.ascontiguousarray(d, dtype=np.float32)


def box_near_far(bounds, ray_o, ray_d):
    """Slab test of rays against an axis-aligned box padded by 0.01.

    Returns near, far (float32, in units of ray_d like the reference, which
    divides the hit distance by |ray_d|) and the hit mask.
    """
    lo = bounds[0].astype(np.float64) - 0.01
    hi = bounds[1].astype(np.float64) + 0.01
    o = ray_o.astype(np.float64)
    d = ray_d.astype(np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        t0 = (lo[None] - o) / d
        t1 = (hi[None] - o) / d
    tmin = np.minimum(t0, t1)
    tmax = np.maximum(t0, t1)
    t_in = np.nanmax(tmin, axis=1)
    t_out = np.nanmin(tmax, axis=1)
    hit = (t_in < t_out) & (t_out > 0)
    near = np.where(hit, t_in, MISS_NEAR_FAR[0]).astype(np.float32)
    far = np.where(hit, t_out, MISS_NEAR_FAR[1]).astype(np.float32)
    return near, far, hit


def fake_smpl(faces, rng):
    """Stand-in for SMPL_NEUTRAL.pkl with the three keys the reference reads
    (utils/smpl_utils.py:3-14, can_render.py:384-395)."""
    kintree = np.stack([np.arange(-1, 23), np.arange(24)]).astype(np.int64)
    w = rng.rand(V, 24).astype(np.float32)
    w = w / w.sum(1, keepdims=True)
    return {"kintree_table": kintree, "weights": w, "f": faces.astype(np.uint32)}


def make_scene(H=64, W=64, pose_seed=0):
    """Everything the reference's ``batch`` dict carries, as numpy arrays.

    RNG order: ``np.random.seed(pose_seed)`` -> blend
    weights -> poses.
    """
    rng = np.random.RandomState(pose_seed)
    canonical, posed, faces = body_meshes()
    smpl = fake_smpl(faces, rng)
    poses = (rng.randn(24, 3) * 0.2).astype(np.float32)
    K, R, T = camera(H, W)
    ray_o, ray_d = pinhole_rays(H, W, K, R, T)
    bounds = np.stack([posed.min(0), posed.max(0)]).astype(np.float32)
    bounds[0, 2] -= 0.05
    bounds[1, 2] += 0.05
    near, far, hit = box_near_far(bounds, ray_o, ray_d)
    return {
        "H": H,
        "W": W,
        "canonical": canonical,
        "posed": posed,
        "faces": faces,
        "smpl": smpl,
        "poses": poses,
        "frame": 5,
        "Th": POSE_SHIFT.reshape(1, 3).copy(),
        "K": K,
        "R": R,
        "T": T,
        "ray_o": ray_o,
        "ray_d": ray_d,
        "near": near,
        "far": far,
        "hit_box": hit,
        "mask_at_box": np.ones(H * W, dtype=bool),
        "bounds": bounds,
    }


def to_batch(scene, torch, device="cpu", rays=None):
    """Pack a scene into the reference's batch-dict layout (leading B=1).

    ``rays`` optionally selects a subset of ray indices.
    """
    sel = slice(None) if rays is None else rays
    # copies: the reference overwrites near/far in place (utils/pts_utils.py:52-53)
    t = lambda a: torch.from_numpy(np.array(a, copy=True)).to(device)
    H, W = scene["H"], scene["W"]
    return {
        "ray_o": t(scene["ray_o"][sel])[None],
        "ray_d": t(scene["ray_d"][sel])[None],
        "near": t(scene["near"][sel])[None],
        "far": t(scene["far"][sel])[None],
        "xyz": t(scene["posed"])[None],
        "poses": t(scene["poses"])[None],
        "Th": t(scene["Th"])[None],
        "frame": torch.tensor([scene["frame"]], dtype=torch.long, device=device),
        "img": torch.zeros(1, H, W, 3),
        "mask_at_box": t(scene["mask_at_box"])[None],
    }
