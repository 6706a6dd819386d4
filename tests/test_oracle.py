"""Pin the CPU oracle: against committed golden vectors generated from the
unmodified reference (tests/make_golden.py) and, when /root/reference is
present, against the live reference."""
import os

import numpy as np
import pytest
import torch

import common as C
import refharness as RH
from dual_space_nerf_b200 import scene as S
from oracle import clib
from oracle import oracle as O


def test_linspace_matches_torch():
    for n in (2, 3, 5, 16, 32, 33, 64, 65, 128, 192):
        assert np.array_equal(clib.linspace01(n), torch.linspace(0.0, 1.0, steps=n).numpy()), n


def test_geometry_ops_match_torch_bitwise():
    rng = np.random.RandomState(0)
    tri = rng.randn(4000, 3, 3).astype(np.float32)
    pts = rng.randn(4000, 3).astype(np.float32)
    a, b = torch.from_numpy(tri[:, 1] - tri[:, 0]), torch.from_numpy(tri[:, 2] - tri[:, 0])
    n = torch.cross(a, b, dim=-1)
    assert C.bits_equal(clib.norm3(n.numpy()), torch.norm(n, dim=-1).numpy()) == 0
    verts = rng.randn(500, 3).astype(np.float32)
    faces = rng.randint(0, 500, size=(900, 3))
    ref_c = torch.from_numpy(verts)[torch.from_numpy(faces)].mean(dim=-2).numpy()
    assert C.bits_equal(clib.centroids(verts, faces), ref_c) == 0
    # brute-force NN agrees with a float64 argmin except on fp32 near-ties
    cent = rng.randn(3000, 3).astype(np.float32)
    idx, d2 = clib.nearest(pts, cent, want_d2=True)
    d64 = ((pts[:, None].astype(np.float64) - cent[None]) ** 2).sum(-1)
    assert (idx == d64.argmin(1)).mean() > 0.999
    assert np.allclose(d2, d64.min(1), rtol=1e-6)
    # ties: duplicate centroid -> lowest index
    cent2 = np.concatenate([cent, cent[:10]])
    idx2 = clib.nearest(pts, cent2)
    assert np.array_equal(idx, idx2)


def _oracle_for(sc, sd, n, **kw):
    return O.Oracle(sd, sc["canonical"], sc["faces"], n, **kw)


def test_stages_vs_golden(scene64, state_dict):
    g = C.golden("stages_64x64x32.npz")
    rays = g["rays"]
    sc = scene64
    st = {}
    _oracle_for(sc, state_dict, 32).render(sc["ray_o"][rays], sc["ray_d"][rays], sc["near"][rays], sc["far"][rays],
                                          sc["posed"], sc["poses"], sc["frame"], stages=st)
    for k in ("near_gg", "far_gg", "z_vals", "pts", "uv", "h", "xyz_cano"):
        assert C.bits_equal(st[k], g[k]) == 0, k  # bit-exact geometry
    assert np.array_equal(st["idx"], g["idx"])
    assert np.array_equal(st["mask"], g["mask"])
    act = ~g["mask"]
    assert np.abs(st["pose_feat"] - g["pose_feat"]).max() < 1e-7
    assert np.abs(st["density"][act] - g["density"][act]).max() < 2e-3  # |sigma| ~ 170, fp32 GEMM order
    assert np.abs(st["essence"][act] - g["essence"][act]).max() < 1e-5
    kink = st["kink_margin"][act] < C.KINK_MARGIN
    gn = np.abs(st["grad"][act] - g["grad"][act]).max(1) / np.abs(g["grad"][act]).max(1)
    assert gn[~kink].max() < 1e-4
    assert np.abs(st["normal_world"][act] - g["normal_world"][act]).max(1)[~kink].max() < 1e-3
    assert np.abs(st["color"][act] - g["color"][act]).max(1)[~kink].max() < 1e-5


def _kink_rays(st, n):
    return (st["kink_margin"] < C.KINK_MARGIN).reshape(-1, n).any(1)


def test_render_view_vs_golden(scene64, state_dict):
    g = C.golden("render_64x64x32.npz")
    sc = scene64
    st = {}
    out = _oracle_for(sc, state_dict, 32).render(sc["ray_o"], sc["ray_d"], sc["near"], sc["far"], sc["posed"],
                                                 sc["poses"], sc["frame"], stages=st)
    ref = {"color": g["coarse_color"].reshape(-1, 3), "depth_map": g["coarse_depth"].ravel(),
           "acc_map": g["coarse_acc"].ravel(), "disp_map": g["coarse_disp"].ravel()}
    stats = C.check_rays(out, ref, _kink_rays(st, 32), what="render_view 64x64x32")
    assert stats["rgb_max_nokink"] < 1e-5


def test_uniform_mode_vs_golden(scene64, state_dict):
    g = C.golden("render_uniform.npz")
    rays = g["rays"]
    sc = scene64
    st = {}
    out = _oracle_for(sc, state_dict, 16, mode="uniform").render(
        sc["ray_o"][rays], sc["ray_d"][rays], sc["near"][rays], sc["far"][rays], sc["posed"], sc["poses"], sc["frame"], stages=st)
    C.check_rays(out, g, _kink_rays(st, 16), what="uniform")
    assert C.bits_equal(out["z_vals"], g["z_vals"]) == 0
    assert np.abs(out["weights"] - g["weights"]).max() < 1e-5


def test_novel_pose_vs_golden(state_dict):
    g = C.golden("render_novelpose.npz")
    rays = g["rays"]
    sc = S.make_scene(64, 64, pose_seed=3)
    st = {}
    out = _oracle_for(sc, state_dict, 32, zero_code=True, light_center=S.LIGHT_CENTER_313).render(
        sc["ray_o"][rays], sc["ray_d"][rays], sc["near"][rays], sc["far"][rays], sc["posed"], sc["poses"], sc["frame"],
        Th=sc["Th"], stages=st)
    C.check_rays(out, g, _kink_rays(st, 32), what="novel pose")


def test_rot_and_trained_magnitude_weights_vs_golden(state_dict):
    """Round-2 goldens (tests/make_golden_r2.py): net.set_rot / set_rot_center (model/spacenet.py:254-258) and a network whose
    hidden layers have an rms gain > 1 (gradients grow through the backward chain)."""
    from make_golden_r2 import ROT_ANGLE, ROT_CENTER, angle2rot, big_weight_net

    g = C.golden("render_rot.npz")
    rays = g["rays"]
    sc = S.make_scene(64, 64, pose_seed=2)
    st = {}
    out = _oracle_for(sc, state_dict, 32, light_center=S.LIGHT_CENTER_313, rot=angle2rot(ROT_ANGLE).astype(np.float32),
                      rot_center=ROT_CENTER).render(
        sc["ray_o"][rays], sc["ray_d"][rays], sc["near"][rays], sc["far"][rays], sc["posed"], sc["poses"], sc["frame"],
        Th=sc["Th"], stages=st)
    C.check_rays(out, g, _kink_rays(st, 32), what="rot", strict=True)
    g = C.golden("render_bigw.npz")
    rays = g["rays"]
    sc = S.make_scene(64, 64)
    st = {}
    out = _oracle_for(sc, big_weight_net(0).state_dict(), 32).render(
        sc["ray_o"][rays], sc["ray_d"][rays], sc["near"][rays], sc["far"][rays], sc["posed"], sc["poses"], sc["frame"],
        Th=sc["Th"], stages=st)
    C.check_rays(out, g, _kink_rays(st, 32), what="trained-magnitude weights", strict=True)
    assert np.abs(st["grad"]).max() > 1e3  # the chain does grow (default init: ~1e-2)


def test_128x128x64_vs_golden(state_dict):
    g = C.golden("render_128x128x64.npz")
    rays = g["rays"]
    sc = S.make_scene(128, 128)
    st = {}
    out = _oracle_for(sc, state_dict, 64).render(
        sc["ray_o"][rays], sc["ray_d"][rays], sc["near"][rays], sc["far"][rays], sc["posed"], sc["poses"], sc["frame"], stages=st)
    C.check_rays(out, g, _kink_rays(st, 64), what="128x128x64")
    assert C.bits_equal(out["z_vals"], g["z_vals"]) == 0


@pytest.mark.skipif(not RH.available(), reason="/root/reference not present")
def test_live_reference_weights_and_render(scene64, state_dict):
    RH._install_shims()
    from model.spacenet import DualSpaceNeRF as RefNet

    from dual_space_nerf_b200 import net as N

    torch.manual_seed(0)
    ref = RefNet(None)
    N.synthetic_head_rescale_(ref)
    rsd = ref.state_dict()
    assert list(rsd.keys()) == N.STATE_DICT_ORDER
    for k in state_dict:
        assert torch.equal(rsd[k], state_dict[k]), k
    sc = scene64
    rays = np.nonzero(sc["hit_box"])[0][5::40]
    rig = RH.ReferenceRig(sc, 32, state_dict)
    ref_out = rig.render(rays)
    st = {}
    out = _oracle_for(sc, state_dict, 32).render(sc["ray_o"][rays], sc["ray_d"][rays], sc["near"][rays], sc["far"][rays],
                                                 sc["posed"], sc["poses"], sc["frame"], stages=st)
    C.check_rays(out, ref_out, _kink_rays(st, 32), what="live reference")


def test_sample_pdf_properties():
    rng = np.random.RandomState(0)
    z = np.sort(rng.rand(50, 64).astype(np.float32) + 2.0, -1)
    w = rng.rand(50, 64).astype(np.float32)
    w[:, :20] = 0
    z2 = O.sample_pdf(z, w, 128)
    assert z2.shape == (50, 192)
    assert np.all(np.diff(z2, axis=-1) >= 0)
    assert np.all(z2 >= z[:, :1]) and np.all(z2 <= z[:, -1:])
    for r in range(50):  # the coarse samples survive the merge
        assert np.all(np.isin(z[r], z2[r]))


def test_ppts_to_pts_oracle_vs_reference_golden():
    """oracle.ppts_to_pts against the golden made by the reference's own utils/blend_utils.py:ppts_to_pts (tests/make_golden_lbs.py).
    fp32 blend + 3x3 inverse: tolerance 2e-6 absolute on points of magnitude <= 1.7."""
    g = C.golden("lbs_ppts_to_pts.npz")
    out = O.ppts_to_pts(g["pts"], g["bw"], g["A"])
    assert out.dtype == np.float32 and out.shape == g["out"].shape
    assert np.abs(out - g["out"]).max() < 2e-6


def test_camera_rays_oracle_vs_reference_golden_and_live():
    """oracle.camera_rays / box_near_far (utils/rays_utils.py:16-30, 63-97) against the golden made by the reference's own
    functions (tests/make_golden_camera.py), and against the live reference when /root/reference is mounted."""
    g = C.golden("camera_rays.npz")
    for c in range(2):
        H, W = int(g[f"H{c}"]), int(g[f"W{c}"])
        o, d = O.camera_rays(H, W, g[f"K{c}"], g[f"R{c}"], g[f"T{c}"])
        o32, d32 = o.astype(np.float32), d.astype(np.float32)
        assert np.array_equal(o32, g[f"ray_o{c}"]) and np.array_equal(d32, g[f"ray_d{c}"])
        near, far, mask = O.box_near_far(g[f"bounds{c}"], np.broadcast_to(o32, d32.shape), d32)
        assert np.array_equal(mask, g[f"mask{c}"])
        assert np.array_equal(near.astype(np.float32), g[f"near{c}"]) and np.array_equal(far.astype(np.float32), g[f"far{c}"])
    ref = "/root/reference/utils/rays_utils.py"
    if os.path.exists(ref):
        import importlib.util

        spec = importlib.util.spec_from_file_location("ref_rays_utils", ref)
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        sc = S.make_scene(32, 40)
        ro, rd = m.get_rays(32, 40, sc["K"], sc["R"], sc["T"].reshape(3, 1))
        o, d = O.camera_rays(32, 40, sc["K"], sc["R"], sc["T"])
        assert np.array_equal(rd.reshape(-1, 3).astype(np.float32), d.astype(np.float32))


def test_training_mode_forward_vs_reference_golden(state_dict):
    """Training-mode forward: stratified jitter (utils/pts_utils.py:6-13) and density noise
    (utils/nerf_net_utils.py:29-33) with the draws as inputs, against the UNMODIFIED reference run by
    tests/make_golden_train.py: z bit-exact, per-ray outputs within 1e-5 (fp32 summation order)."""
    g = C.golden("render_train.npz")
    sc = S.make_scene(64, 64)
    rays = g["rays"]
    o = O.Oracle(state_dict, sc["canonical"], sc["faces"], 32)
    args = (sc["ray_o"][rays], sc["ray_d"][rays], sc["near"][rays], sc["far"][rays], sc["posed"], sc["poses"], sc["frame"])
    out = o.render(*args, Th=sc["Th"], t_rand=g["t_rand"], noise=g["noise"])
    assert C.bits_equal(out["z_vals"], g["z_vals"]) == 0
    for k in ("color", "depth_map", "acc_map", "weights"):
        assert np.abs(out[k] - g[k]).max() < 1e-5, k
    jit = o.render(*args, Th=sc["Th"], t_rand=g["t_rand"])  # jitter only: transparent samples are skipped exactly
    for k in ("color", "depth_map", "acc_map", "weights"):
        assert np.abs(jit[k] - g["jitter_only_" + k]).max() < 1e-5, k
    # the noise matters: the two goldens differ by far more than the tolerance
    assert np.abs(g["acc_map"] - g["jitter_only_acc_map"]).max() > 0.05


def test_compositor_and_jitter_properties():
    """Size-independent properties of the oracle's sampling / compositing restatement (utils/pts_utils.py:3-16,
    utils/nerf_net_utils.py:5-56) on random inputs: weights in [0,1] with sum <= 1 (+ rounding), colour inside the convex
    hull scaled by acc, zero density => exactly zero weight and NaN disparity, jittered z sorted and inside [near, far]."""
    rng = np.random.RandomState(5)
    R, N = 257, 37
    z = np.sort(rng.rand(R, N).astype(np.float32) * 2 + 2, axis=1)
    rd = rng.randn(R, 3).astype(np.float32)
    rgb = rng.rand(R, N, 3).astype(np.float32)
    sig = (rng.randn(R, N) * 30).astype(np.float32)
    out = O.raw2outputs(rgb, sig, z, rd)
    w = out["weights"]
    assert w.min() >= 0 and w.max() <= 1 and (w.sum(1) <= 1 + 1e-5).all()
    assert np.all(w[sig <= 0] == 0)                                    # relu(sigma) = 0 => alpha exactly 0
    assert np.all(out["color"] <= out["acc_map"][:, None] * 1.00001 + 1e-6)  # colours in [0,1] => rgb_map <= acc
    assert np.all((out["depth_map"] >= out["acc_map"] * z[:, 0] * 0.99999) & (out["depth_map"] <= out["acc_map"] * z[:, -1] * 1.00001))
    dead = O.raw2outputs(rgb, -np.abs(sig), z, rd)
    assert np.all(dead["weights"] == 0) and np.all(dead["acc_map"] == 0) and np.isnan(dead["disp_map"]).all()
    near = (rng.rand(R).astype(np.float32) + 2)
    far = near + rng.rand(R).astype(np.float32) + np.float32(0.1)
    t = rng.rand(R, N).astype(np.float32)
    o = np.zeros((R, 3), np.float32)
    _, zj = O.uniform_sampling(o, rd, N, near, far, t)
    _, z0 = O.uniform_sampling(o, rd, N, near, far)
    assert np.all(np.diff(zj, axis=1) >= 0) and np.all(zj >= near[:, None]) and np.all(zj <= far[:, None])
    _, zlo = O.uniform_sampling(o, rd, N, near, far, np.zeros_like(t))   # t_rand = 0 picks the lower interval ends
    assert np.array_equal(zlo[:, 0], z0[:, 0]) and np.all(zlo[:, 1:] <= z0[:, 1:])
