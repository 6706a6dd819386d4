"""CPU oracle (numpy + oracle/geom.c) for the Dual-Space-NeRF render path.

TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Importers allowed: tests/,
__graft_entry__.smoke(), bench.py's cpu_baseline / `--impl reference` legs.

Every function restates one reference function (file:line cited) in fp32, one
numpy op per torch op so that roundings happen in the same places.  The
discrete stages (nearest triangle, GG bounds) go through oracle/geom.c.
Differences to the reference that are deliberate and exact:

* the density gradient is computed analytically (backward through the ReLU
  masks and the positional encoding) instead of ``torch.autograd.grad``
  (model/spacenet.py:301-311);
* ``ray_d_can`` (can_render.py:367-376) is not computed: SpaceNet.use_dir is
  False (model/spacenet.py:21), nothing reads it.

Pinning: checked against the live reference (tests/refharness.py) by
tests/test_oracle.py and against tests/golden/*.npz generated from it
(tests/make_golden.py).  Config 3 (hierarchical) has no reference implementation
(SURVEY.md 0 #4): `sample_pdf` below is our own spec -> PARITY UNPINNED there.
"""
from __future__ import annotations

import numpy as np

from . import clib

f32 = np.float32
GAMMA = 0.05
MASK_FLOOR, MASK_CEIL, MASK_MAX_DIST = -4.0, 5.0, 0.1
PE_L = 10


# --------------------------------------------------------------------------- sampling
def uniform_sampling(ray_o, ray_d, n_pts, near, far, t_rand=None):
    """utils/pts_utils.py:3-16.  ``t_rand`` (R,N) in [0,1) is the training-mode draw of ``torch.rand`` (:12): the
    stratified jitter of :6-13 is applied when it is given (perturb > 0 and net.training), eval mode otherwise."""
    t = clib.linspace01(n_pts)
    z = near[:, None] * (f32(1.0) - t)[None] + far[:, None] * t[None]
    if t_rand is not None:
        z = z.astype(f32)
        mids = f32(0.5) * (z[:, 1:] + z[:, :-1])
        upper = np.concatenate([mids, z[:, -1:]], -1)
        lower = np.concatenate([z[:, :1], mids], -1)
        z = lower + (upper - lower) * np.asarray(t_rand, f32)
    pts = ray_o[:, None, :] + ray_d[:, None, :] * z[..., None]
    return pts.astype(f32), z.astype(f32)


def gg_sampling(ray_o, ray_d, n_pts, near, far, xyz, t_rand=None):
    """utils/pts_utils.py:18-58 (geometry-guided near/far then uniform samples)."""
    near2, far2 = clib.gg_bounds(ray_o[0], ray_d, xyz, near, far, GAMMA)
    pts, z = uniform_sampling(ray_o, ray_d, n_pts, near2, far2, t_rand)
    return pts, z, near2, far2


# --------------------------------------------------------------------------- warp
def _sum3(x):
    return (x[:, 0] + x[:, 1]) + x[:, 2]


def _norm3(x):
    """torch.norm(dim=-1) of 3-vectors: fma-accumulated (oracle/geom.c:dso_norm3)."""
    return clib.norm3(np.ascontiguousarray(x, f32))


def project_point2mesh(pts, tri):
    """utils/geo_utils.py:181-200 + get_barycentric_coordinates :96-113 (oracle/geom.c:dso_project)."""
    return clib.project(pts, tri)


def transparent_mask(uv, h):
    """utils/render_utils.py:103-109."""
    c = (uv > f32(MASK_CEIL)) | (uv < f32(MASK_FLOOR))
    return c[:, 0] | c[:, 1] | (np.abs(h) > f32(MASK_MAX_DIST))


def barycentric_map2can(uv, h, tri):
    """utils/geo_utils.py:138-156 (oracle/geom.c:dso_map2can)."""
    return clib.map2can(uv, h, tri)


def world_to_canonical(pts, posed, canonical, faces):
    """Renderer.w2l_without_lbs, can_render.py:333-379 (without the dead ray_d_can half)."""
    cent = clib.centroids(posed, faces)
    idx = clib.nearest(pts, cent)
    tri_w = posed[faces[idx]]
    uv, h = project_point2mesh(pts, tri_w)
    mask = transparent_mask(uv, h)
    cano = barycentric_map2can(uv, h, canonical[faces[idx]])
    return cano.astype(f32), mask, idx, uv, h


# --------------------------------------------------------------------------- network
def positional_encoding(x, L=PE_L):
    """model/dimension_kernel.py:5-35: [x, sin(2^0 x), cos(2^0 x), ..., sin(2^(L-1) x), cos(2^(L-1) x)]."""
    outs = [x]
    for k in range(L):
        fx = x * f32(2.0 ** k)
        outs.append(np.sin(fx))
        outs.append(np.cos(fx))
    return np.concatenate(outs, -1).astype(f32)


def rod2quat(rot_vecs):
    """model/spacenet.py:314-331."""
    rv = rot_vecs.astype(f32)
    a = rv + f32(1e-16)
    angle = np.sqrt(_sum3(a * a))[:, None]
    d = rv / angle
    c = np.cos(angle / f32(2))
    s = np.sin(angle / f32(2))
    return np.concatenate([d * s, c - f32(1.0)], 1).astype(f32)


class Weights:
    """numpy view of DualSpaceNeRF.state_dict()."""

    def __init__(self, state_dict):
        g = lambda k: np.ascontiguousarray(np.asarray(state_dict[k].detach().cpu().numpy() if hasattr(state_dict[k], "detach") else state_dict[k], dtype=f32))
        self.embedding = g("nerf.embedding.weight")
        self.s1 = [(g(f"nerf.stage1.{i}.weight"), g(f"nerf.stage1.{i}.bias")) for i in (0, 2, 4, 6)]
        self.s2 = [(g(f"nerf.stage2.{i}.weight"), g(f"nerf.stage2.{i}.bias")) for i in (0, 2, 4)]
        self.dens = (g("nerf.density_net.0.weight"), g("nerf.density_net.0.bias"))
        self.rgb = [(g(f"nerf.rgb_net.{i}.weight"), g(f"nerf.rgb_net.{i}.bias")) for i in (1, 3)]
        self.light = [(g(f"lighting_mlp.lights_encoding.{i}.weight"), g(f"lighting_mlp.lights_encoding.{i}.bias")) for i in (0, 2, 4)]
        self.pose = [(g(f"pose_mlp.{i}.weight"), g(f"pose_mlp.{i}.bias")) for i in (0, 2, 4)]


def _lin(x, wb):
    return x @ wb[0].T + wb[1]


def pose_feature(W, poses):
    """model/spacenet.py:223-236: quaternion-ish encoding of joints 1..23 -> pose_mlp."""
    q = rod2quat(poses[1:].reshape(-1, 3)).reshape(1, -1)
    x = np.maximum(_lin(q, W.pose[0]), 0)
    x = np.maximum(_lin(x, W.pose[1]), 0)
    return _lin(x, W.pose[2])[0].astype(f32)


def spacenet_forward(W, xyz_cano, code, pose_feat, want_grad=True, density_only=False, rounder=None, extra=None):
    """SpaceNet.forward (model/spacenet.py:93-148) plus the analytic d(density)/d(xyz)
    that the reference gets from autograd (model/spacenet.py:251,301-311).

    ``rounder(name, x, w)`` optionally emulates reduced-precision GEMM operands
    (used only by the precision study in tests/).  ``extra`` (dict) receives
    ``relu_sig`` (a per-sample hash of all 7x256 ReLU on/off bits) and
    ``kink_margin`` (min over units of |pre-activation| / (sum|w||x| + |b|)): the
    density gradient is discontinuous where a pre-activation crosses zero, so a
    sample whose margin is within rounding distance of 0 has no stable normal.
    """
    mm = (lambda name, x, w: x @ w) if rounder is None else rounder
    P = xyz_cano.shape[0]
    pe = positional_encoding(xyz_cano)
    x0 = np.concatenate([np.broadcast_to(code[None], (P, code.shape[0])), pe, np.broadcast_to(pose_feat[None], (P, 16))], 1).astype(f32)
    acts = []
    margin = np.full(P, np.inf, f32)

    def note(a, x, wb):
        if extra is not None:
            scale = np.abs(x) @ np.abs(wb[0].T) + np.abs(wb[1])
            np.minimum(margin, np.min(np.abs(a) / scale, axis=1), out=margin)

    x = x0
    for i, wb in enumerate(W.s1):
        a = mm(f"s1.{i}", x, wb[0].T) + wb[1]
        note(a, x, wb)
        acts.append(a > 0)
        x = np.maximum(a, 0)
    x = np.concatenate([x, pe], 1)
    for i, wb in enumerate(W.s2):
        a = mm(f"s2.{i}", x, wb[0].T) + wb[1]
        note(a, x, wb)
        acts.append(a > 0)
        x = np.maximum(a, 0)
    feat = x
    if extra is not None:
        rs = np.random.RandomState(1234).randint(1, 2 ** 31 - 1, size=(7, 256)).astype(np.uint64)
        sig = np.zeros(P, np.uint64)
        for l, m in enumerate(acts):
            sig += (m.astype(np.uint64) * rs[l][None]).sum(1)
        extra["relu_sig"] = sig
        extra["kink_margin"] = margin
    density = (mm("dens", feat, W.dens[0].T) + W.dens[1])[:, 0]
    if density_only:
        return density
    r = np.maximum(mm("rgb.0", feat, W.rgb[0][0].T) + W.rgb[0][1], 0)
    essence = mm("rgb.1", r, W.rgb[1][0].T) + W.rgb[1][1]
    grad = None
    if want_grad:
        g = np.broadcast_to(W.dens[0], (P, 256)).astype(f32)  # d sigma / d feat
        g = g * acts[6]
        g = mm("b.s2.2", g, W.s2[2][0]) * acts[5]
        g = mm("b.s2.1", g, W.s2[1][0]) * acts[4]
        g = mm("b.s2.0", g, W.s2[0][0])  # (P, 319): [h | pe]
        gpe = g[:, 256:]
        g = g[:, :256] * acts[3]
        g = mm("b.s1.3", g, W.s1[3][0]) * acts[2]
        g = mm("b.s1.2", g, W.s1[2][0]) * acts[1]
        g = mm("b.s1.1", g, W.s1[1][0]) * acts[0]
        g0 = mm("b.s1.0", g, W.s1[0][0][:, 8:71])  # PE columns only
        gpe = gpe + g0
        grad = gpe[:, 0:3].copy()
        for k in range(PE_L):
            fr = f32(2.0 ** k)
            s = pe[:, 3 + 6 * k: 6 + 6 * k]
            c = pe[:, 6 + 6 * k: 9 + 6 * k]
            grad = grad + (gpe[:, 3 + 6 * k: 6 + 6 * k] * c - gpe[:, 6 + 6 * k: 9 + 6 * k] * s) * fr
        grad = grad.astype(f32)
    return essence.astype(f32), density.astype(f32), grad


def _normalize(x, eps=1e-12):
    n = _norm3(x)
    return x / np.maximum(n, f32(eps))[:, None]


def normal_local2world(grad, xyz_cano, posed, canonical, faces):
    """model/spacenet.py:278-298."""
    cent_c = clib.centroids(canonical, faces)
    idx = clib.nearest(xyz_cano, cent_c)
    tri_c = canonical[faces[idx]]
    tri_w = posed[faces[idx]]
    uv, h = project_point2mesh(xyz_cano, tri_c)
    start = barycentric_map2can(uv, h, tri_w)
    uv, h = project_point2mesh(xyz_cano + grad, tri_c)
    end = barycentric_map2can(uv, h, tri_w)
    return _normalize(end - start).astype(f32), idx


def lighting(W, normal_w, xyz_world, view_dir, essence, rounder=None):
    """LightingMLP.forward, model/spacenet.py:174-188 (L=0 encodings are identities)."""
    mm = (lambda name, x, w: x @ w) if rounder is None else rounder
    v = view_dir / _norm3(view_dir)[:, None]
    x = np.concatenate([normal_w, xyz_world, v], 1).astype(f32)
    x = np.maximum(mm("l.0", x, W.light[0][0].T) + W.light[0][1], 0)
    x = np.maximum(mm("l.1", x, W.light[1][0].T) + W.light[1][1], 0)
    x = mm("l.2", x, W.light[2][0].T) + W.light[2][1]
    x = np.where(x > 0, x, np.expm1(np.minimum(x, 0)))  # ELU(alpha=1)
    return ((x + f32(1.0)) * essence).astype(f32)


# --------------------------------------------------------------------------- compositing
def raw2outputs(rgb, sigma, z_vals, rays_d, noise=None):
    """utils/nerf_net_utils.py:5-56 with white_bkgd=False.  ``noise`` (R,N) = randn * raw_noise_std is the training-mode
    draw of :29-33, added to the density before the ReLU; None = raw_noise_std 0."""
    if noise is not None:
        sigma = (sigma + np.asarray(noise, f32)).astype(f32)
    R, N = z_vals.shape
    dists = z_vals[:, 1:] - z_vals[:, :-1]
    dists = np.concatenate([dists, np.full((R, 1), 1e10, f32)], 1)
    dists = dists * _norm3(rays_d)[:, None]
    with np.errstate(over="ignore"):
        alpha = f32(1.0) - np.exp(-np.maximum(sigma, 0) * dists)
    t = f32(1.0) - alpha + f32(1e-10)
    T = np.ones((R, N), f32)
    for i in range(1, N):
        T[:, i] = T[:, i - 1] * t[:, i - 1]
    weights = (alpha * T).astype(f32)
    rgb_map = np.zeros((R, 3), f32)
    depth = np.zeros(R, f32)
    acc = np.zeros(R, f32)
    for i in range(N):  # torch.sum over a strided dim accumulates in index order
        rgb_map = rgb_map + weights[:, i, None] * rgb[:, i]
    depth = np.sum(weights * z_vals, -1, dtype=f32)
    acc = np.sum(weights, -1, dtype=f32)
    with np.errstate(divide="ignore", invalid="ignore"):
        disp = f32(1.0) / np.maximum(f32(1e-10), depth / acc)
    return {"color": rgb_map, "disp_map": disp.astype(f32), "acc_map": acc, "depth_map": depth, "weights": weights}


# --------------------------------------------------------------------------- hierarchical (own spec)
def sample_pdf(z_vals, weights, n_importance):
    """Deterministic inverse-CDF resampling (NeRF `sample_pdf`, det=True).

    The reference calls an undefined ``Renderer.resampling`` (can_render.py:213),
    so this is OUR specification for config 3, not a restatement:
      bins = midpoints of z_vals; pdf ~ weights[1:-1] + 1e-5; u = linspace(0,1,n);
      new z by linear interpolation of the CDF; merged = sort(concat(z, new)).
    """
    R, N = z_vals.shape
    bins = f32(0.5) * (z_vals[:, 1:] + z_vals[:, :-1])
    w = weights[:, 1:-1] + f32(1e-5)
    pdf = w / np.sum(w, -1, keepdims=True, dtype=f32)
    cdf = np.concatenate([np.zeros((R, 1), f32), np.cumsum(pdf, -1, dtype=f32)], -1)
    u = clib.linspace01(n_importance)
    out = np.empty((R, n_importance), f32)
    for r in range(R):
        inds = np.searchsorted(cdf[r], u, side="right")
        below = np.maximum(inds - 1, 0)
        above = np.minimum(inds, cdf.shape[1] - 1)
        c0, c1 = cdf[r][below], cdf[r][above]
        b0, b1 = bins[r][below], bins[r][above]
        den = c1 - c0
        den = np.where(den < f32(1e-5), f32(1.0), den)
        t = (u - c0) / den
        out[r] = b0 + t * (b1 - b0)
    return np.sort(np.concatenate([z_vals, out], -1), -1).astype(f32)


# --------------------------------------------------------------------------- camera rays
def camera_rays(H, W, K, R, T):
    """utils/rays_utils.py:16-30 get_rays, inference layout: origin (3,) and directions (H*W,3) in float64 (the callers cast
    to float32, rays_utils.py:175-176).  Pixel (i = column, j = row), xy1 = (i, j, 1)."""
    K, R, T = np.asarray(K, np.float64), np.asarray(R, np.float64), np.asarray(T, np.float64).reshape(3)
    origin = -(R.T @ T)
    ii, jj = np.meshgrid(np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64), indexing="xy")
    xy1 = np.stack([ii, jj, np.ones_like(ii)], 2).reshape(-1, 3)
    cam = xy1 @ np.linalg.inv(K).T
    world = (cam - T[None]) @ R
    return origin, world - origin[None]


def box_near_far(bounds, ray_o, ray_d):
    """utils/rays_utils.py:63-97 get_near_far: the six plane hits of every ray with the box padded by 0.01, the rays with
    exactly two hits inside the (eps = 1e-6) box are kept; near / far = the two hit distances divided by |ray_d|.
    Returns near, far for the kept rays (float64) and the mask over all rays."""
    b = np.asarray(bounds, np.float64) + np.array([-0.01, 0.01])[:, None]
    o, d = np.asarray(ray_o, np.float64), np.asarray(ray_d, np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        t = ((b[None] - o[:, None]) / d[:, None]).reshape(-1, 6)
        pts = t[..., None] * d[:, None] + o[:, None]
    eps = 1e-6
    lo, hi = b[0] - eps, b[1] + eps
    inside = np.all((pts >= lo) & (pts <= hi), axis=-1)
    mask = inside.sum(-1) == 2
    pair = pts[mask][inside[mask]].reshape(-1, 2, 3)
    rd32 = np.asarray(ray_d, f32)[mask]  # np.linalg.norm of the float32 directions stays in float32 (rays_utils.py:90)
    nd = np.sqrt(np.sum(rd32 * rd32, axis=1, dtype=f32), dtype=f32).astype(np.float64)
    d0 = np.sqrt(np.sum((pair[:, 0] - o[mask]) ** 2, axis=1)) / nd
    d1 = np.sqrt(np.sum((pair[:, 1] - o[mask]) ** 2, axis=1)) / nd
    return np.minimum(d0, d1), np.maximum(d0, d1), mask


# --------------------------------------------------------------------------- inverse LBS (auxiliary op)
def ppts_to_pts(pts, bw, A):
    """utils/blend_utils.py:72-81 (the binding definition of ppts_to_pts): blend the 24 joint transforms with the per-point
    weights, subtract the blended translation, multiply by the inverse of the blended 3x3.
    pts (P,3), bw (24,P), A (24,4,4) -> (P,3), fp32 throughout (torch.bmm / torch.inverse on CPU)."""
    pts = np.ascontiguousarray(pts, f32)
    Ap = (np.ascontiguousarray(bw, f32).T @ np.ascontiguousarray(A, f32).reshape(24, 16)).reshape(-1, 4, 4)
    d = pts - Ap[:, :3, 3]
    Rinv = np.linalg.inv(Ap[:, :3, :3]).astype(f32)
    return np.sum(Rinv * d[:, None, :], axis=2, dtype=f32).astype(f32)


# --------------------------------------------------------------------------- full path
class Oracle:
    """Renderer.render / render_view restated (can_render.py:137-168, 248-278)."""

    def __init__(self, state_dict, canonical, faces, n_samples, mode="GG", zero_code=False,
                 light_center=None, rot=None, rot_center=None):
        self.W = Weights(state_dict)
        self.canonical = np.ascontiguousarray(canonical, f32)
        self.faces = np.ascontiguousarray(faces, np.int64)
        self.N = n_samples
        self.mode = mode
        self.zero_code = zero_code
        self.light_center = None if light_center is None else np.asarray(light_center, f32)
        self.rot = None if rot is None else np.asarray(rot, f32)
        self.rot_center = None if rot_center is None else np.asarray(rot_center, f32)

    def shade_points(self, pts, z, ray_d, posed, poses, frame, Th=None, stages=None, noise=None):
        """``noise`` given (training mode): a transparent sample then has alpha = 1 - exp(-relu(0 + noise) dist) > 0 and its
        colour counts, so the network runs on EVERY sample as in the reference (can_render.py:113-120 zeroes only the
        density); without noise the masked samples have weight exactly 0 and are skipped."""
        R, N = z.shape
        W = self.W
        posed = np.ascontiguousarray(posed, f32)
        flat = pts.reshape(-1, 3)
        cano, mask, idx, uv, h = world_to_canonical(flat, posed, self.canonical, self.faces)
        code = W.embedding[int(frame)] * (f32(0.0) if self.zero_code else f32(1.0))
        pf = pose_feature(W, poses)
        P = flat.shape[0]
        active = ~mask if noise is None else np.ones(mask.shape, bool)
        essence = np.zeros((P, 3), f32)
        sigma = np.zeros(P, f32)
        color = np.zeros((P, 3), f32)
        grad = np.zeros((P, 3), f32)
        nw = np.zeros((P, 3), f32)
        idx2 = np.full(P, -1, np.int32)
        if active.any():
            a = np.nonzero(active)[0]
            ex = {} if stages is not None else None
            e_a, s_a, g_a = spacenet_forward(W, cano[a], code, pf, extra=ex)
            n_a, i2 = normal_local2world(g_a, cano[a], posed, self.canonical, self.faces)
            xw = flat[a].copy()
            if self.rot is not None and self.rot_center is not None:  # spacenet.py:254-258
                xw[:, :2] = (xw[:, :2] - self.rot_center[:, :2]) @ self.rot + self.rot_center[:, :2]
            if self.light_center is not None:  # spacenet.py:260-263
                xw = xw + (self.light_center - np.mean(np.asarray(Th, f32), axis=0))[None, :3]
            vd = np.repeat(ray_d, N, axis=0)[a]
            c_a = lighting(W, n_a, xw.astype(f32), vd, e_a)
            essence[a], sigma[a], grad[a], nw[a], color[a], idx2[a] = e_a, s_a, g_a, n_a, c_a, i2
        # masked samples: density forced to 0 (can_render.py:118-120) => weight exactly 0 (without noise)
        sigma[mask] = 0
        out = raw2outputs(color.reshape(R, N, 3), sigma.reshape(R, N), z, ray_d, noise)
        out["z_vals"] = z
        if stages is not None:
            sig = np.zeros(P, np.uint64)
            margin = np.full(P, np.inf, f32)
            if active.any():
                sig[a], margin[a] = ex["relu_sig"], ex["kink_margin"]
            stages.update(relu_sig=sig, kink_margin=margin)
            stages.update(idx=idx, uv=uv, h=h, mask=mask, xyz_cano=cano, essence=essence, density=sigma,
                          grad=grad, normal_world=nw, color=color, idx_cano=idx2, pose_feat=pf)
        return out

    def render(self, ray_o, ray_d, near, far, posed, poses, frame, Th=None, stages=None, t_rand=None, noise=None):
        """``t_rand`` / ``noise``: the training-mode draws (see uniform_sampling / raw2outputs); both None = eval mode."""
        ray_o = np.ascontiguousarray(ray_o, f32).reshape(-1, 3)
        ray_d = np.ascontiguousarray(ray_d, f32).reshape(-1, 3)
        near = np.ascontiguousarray(near, f32).reshape(-1)
        far = np.ascontiguousarray(far, f32).reshape(-1)
        posed = np.ascontiguousarray(posed, f32)
        if self.mode == "GG":
            pts, z, n2, f2 = gg_sampling(ray_o, ray_d, self.N, near, far, posed, t_rand)
        else:
            pts, z = uniform_sampling(ray_o, ray_d, self.N, near, far, t_rand)
            n2, f2 = near, far
        if stages is not None:
            stages.update(near_gg=n2, far_gg=f2, pts=pts, z_vals=z)
        return self.shade_points(pts, z, ray_d, posed, poses, frame, Th, stages, noise)

    def render_hierarchical(self, ray_o, ray_d, near, far, posed, poses, frame, n_importance=128, Th=None):
        """Config 3 (own spec): coarse pass -> sample_pdf -> second pass of the SAME net on the
        merged, sorted z (batchify_pts falls back to self.net when fine_net is None, can_render.py:77-78)."""
        coarse = self.render(ray_o, ray_d, near, far, posed, poses, frame, Th)
        z2 = sample_pdf(coarse["z_vals"], coarse["weights"], n_importance)
        ray_o = np.ascontiguousarray(ray_o, f32).reshape(-1, 3)
        ray_d = np.ascontiguousarray(ray_d, f32).reshape(-1, 3)
        pts = (ray_o[:, None, :] + ray_d[:, None, :] * z2[..., None]).astype(f32)
        fine = self.shade_points(pts, z2, ray_d, posed, poses, frame, Th)
        return coarse, fine
