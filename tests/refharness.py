"""Run the UNMODIFIED reference (/root/reference) on CPU -- test infrastructure.

The reference needs four things that are absent in this container:
pytorch3d (-> tests/refshim/pytorch3d), yacs (-> SimpleNamespace cfg), the
licensed SMPL pickle (-> synthetic pickle), and a GPU (-> ``.cuda()`` no-op).
Nothing here is imported by the product package; /root/reference does not exist
on the GPU box, so this module is only used to (re)generate tests/golden/ and
by CPU tests that skip when the reference is missing.
"""
from __future__ import annotations

import os
import pickle
import sys
import tempfile
from types import SimpleNamespace

import numpy as np
import torch

REF_ROOT = os.environ.get("DSNERF_REFERENCE", "/root/reference")
_HERE = os.path.dirname(os.path.abspath(__file__))
_REPO = os.path.dirname(_HERE)


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "model")) and os.path.isfile(os.path.join(REF_ROOT, "can_render.py"))


def _install_shims():
    for p in (_REPO, os.path.join(_HERE, "refshim"), REF_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    if not getattr(torch.Tensor.cuda, "_dsnerf_shim", False):
        def _t_cuda(self, *a, **k):
            return self

        def _m_cuda(self, *a, **k):
            return self

        _t_cuda._dsnerf_shim = True
        torch.Tensor.cuda = _t_cuda
        torch.nn.Module.cuda = _m_cuda


def make_cfg(n_samples, smpl_dir, mode="GG"):
    return SimpleNamespace(
        MODEL=SimpleNamespace(
            TYPE="nerf", COARSE_RAY_SAMPLING=n_samples, FINE_RAY_SAMPLING=-1,
            sample_points_mode=mode, perturb=1.0, raw_noise_std=1.0,
        ),
        DATASETS=SimpleNamespace(SMPL_PATH=smpl_dir),
    )


class ReferenceRig:
    """Reference ``Renderer`` + ``DualSpaceNeRF`` built on a synthetic scene."""

    def __init__(self, scene, n_samples, state_dict, mode="GG"):
        _install_shims()
        from can_render import Renderer  # noqa: reference code
        from model.spacenet import DualSpaceNeRF  # noqa: reference code

        self.scene = scene
        self.tmp = tempfile.mkdtemp(prefix="dsnerf_smpl_")
        with open(os.path.join(self.tmp, "SMPL_NEUTRAL.pkl"), "wb") as f:
            pickle.dump(scene["smpl"], f)
        self.cfg = make_cfg(n_samples, self.tmp, mode)
        torch.manual_seed(0)
        self.net = DualSpaceNeRF(self.cfg)
        self.net.load_state_dict(state_dict)
        canon = torch.from_numpy(scene["canonical"])
        self.renderer = Renderer(self.net, None, self.cfg, canon)
        self.renderer.eval()

    def batch(self, rays=None):
        from dual_space_nerf_b200.scene import to_batch

        return to_batch(self.scene, torch, "cpu", rays)

    def render(self, rays=None):
        b = self.batch(rays)
        out = self.renderer.render(b)["coarse"]
        return {k: v.detach().numpy() for k, v in out.items()}

    def render_train(self, rays, t_rand, noise):
        """``Renderer.render`` in TRAINING mode (perturb = raw_noise_std = 1).  The reference draws the jitter with
        ``torch.rand`` (utils/pts_utils.py:12) and the density noise with ``torch.randn`` (utils/nerf_net_utils.py:31)
        from the global generator; both are replaced for the duration of the call by functions that hand out the given
        arrays, so that the draws are inputs of the golden vector instead of a property of torch's CPU generator."""
        b = self.batch(rays)
        real_rand, real_randn = torch.rand, torch.randn
        used = []

        def fake_rand(*shape, **kw):
            shape = tuple(shape[0]) if len(shape) == 1 and not isinstance(shape[0], int) else tuple(shape)
            assert shape[-2:] == t_rand.shape, shape
            used.append("rand")
            return torch.from_numpy(t_rand.copy()).reshape(shape)

        def fake_randn(*shape, **kw):
            shape = tuple(shape[0]) if len(shape) == 1 and not isinstance(shape[0], int) else tuple(shape)
            assert shape == noise.shape, shape
            used.append("randn")
            return torch.from_numpy(noise.copy())

        self.renderer.train()
        torch.rand, torch.randn = fake_rand, fake_randn
        try:
            out = self.renderer.render(b)["coarse"]
        finally:
            torch.rand, torch.randn = real_rand, real_randn
            self.renderer.eval()
        assert used == ["rand", "randn"], used
        return {k: v.detach().numpy() for k, v in out.items()}

    def render_view(self):
        b = self.batch(None)
        out = self.renderer.render_view(b)
        return {k: v.detach().numpy() for k, v in out.items()}

    def stages(self, rays=None):
        """Stage-boundary tensors of one ``render`` call (eval mode)."""
        r = self.renderer
        b = self.batch(rays)
        ray_o, ray_d = b["ray_o"], b["ray_d"]
        near, far = b["near"].clone(), b["far"].clone()
        pts, z = r.get_sampling_points(ray_o, ray_d, near, far, b["xyz"], mode=self.cfg.MODEL.sample_points_mode)
        out = {"near_gg": near.numpy()[0].copy(), "far_gg": far.numpy()[0].copy(),
               "pts": pts.numpy()[0].copy(), "z_vals": z.numpy()[0].copy()}
        sh = pts.shape
        meshes = b["xyz"][:, r.face_idx]
        from utils.render_utils import get_closest_mesh, get_transparent_mask
        from utils.geo_utils import project_point2mesh, barycentric_map2can

        pw = pts.reshape(1, -1, 3)
        closest, idx = get_closest_mesh(pw, meshes)
        uv, h = project_point2mesh(pw.reshape(-1, 3), meshes=closest.reshape(-1, 3, 3))
        mask = get_transparent_mask(uv, h)
        cano = barycentric_map2can(uv, h, r.canonical_model["meshes"][idx.flatten()])
        out.update(idx=idx.flatten().numpy().astype(np.int32), uv=uv.numpy(), h=h.numpy(),
                   mask=mask.numpy(), xyz_cano=cano.numpy(),
                   centroids=meshes.mean(dim=-2)[0].numpy())
        # network on every sample (what render_rays does)
        pts6, rays6, tmask = r.w2l(pts, ray_o, ray_d, b)
        b["transparent_mask"] = tmask.reshape(-1, sh[2])
        b["canonical_model"] = r.canonical_model
        b["face_idx"] = r.face_idx
        frame_idx = b["frame"][..., None, None].repeat(1, sh[1], sh[2]).reshape(-1, sh[2])
        p6 = pts6.reshape(-1, 6).clone()
        r6 = rays6.reshape(-1, 6)
        net = self.net
        xyz_cano = p6[..., 3:]
        xyz_cano.requires_grad = True
        from model.spacenet import batch_rod2quat, gradient, normal_local2world

        pose = batch_rod2quat(b["poses"][0][1:, :].float().reshape(-1, 3)).reshape(1, -1)
        pose_feat1 = net.pose_mlp(pose)
        essence, density, _ = net.nerf(xyz_cano, r6, frame_idx, False, pose_feat1.repeat(xyz_cano.shape[0], 1))
        g = gradient(xyz_cano, density)
        nw = normal_local2world(g, xyz_cano, b)
        out.update(pose_feat=pose_feat1.detach().numpy()[0], essence=essence.detach().numpy(),
                   density=density.detach().numpy()[:, 0], grad=g.detach().numpy(), normal_world=nw.detach().numpy())
        color, density2, _ = net(pts6.reshape(-1, 6), r6, frame_idx, batch_info=b)
        out.update(color=color.detach().numpy())
        return out
