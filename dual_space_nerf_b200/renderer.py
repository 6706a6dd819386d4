"""Drop-in for ``can_render.Renderer`` of zyhbili/Dual-Space-NeRF on one B200.

Same constructor, methods, batch-dict keys and return dictionaries as the
reference class (can_render.py:14-406); the work happens in libdsnerf.so
(hand-written sm_100a kernels) through the C ABI in include/dsnerf.h.  torch is
used only to hold device memory and to provide the CUDA stream.

Not a port: there is no chunking (the reference chunks 3072 rays at a time to
bound its (V x rays x 3) intermediates, can_render.py:257), no autograd graph,
no per-chunk device->host copy; one call renders the whole ray batch.

``train()`` + ``render(batch)`` / ``render_rays`` is the training-mode FORWARD (stratified
jitter and density noise): the random draws are made on the
device with torch (``self.generator``) and handed to the kernels; no autograd
graph is built, so gradients for training stay with the caller's PyTorch model.
``render_view`` and the hierarchical pass are eval-mode only.
"""
from __future__ import annotations

import ctypes
import os
import pickle

import numpy as np
import torch

from . import lib as _lib
from .net import STATE_DICT_ORDER


def _load_bodydata(model_path, model_type="smpl", gender="neutral"):
    """utils/smpl_utils.py:3-14."""
    if os.path.isdir(model_path):
        model_path = os.path.join(model_path, f"{model_type.upper()}_{gender.upper()}.pkl")
    if not os.path.exists(model_path):
        raise FileNotFoundError(f"Path {model_path} does not exist!")
    with open(model_path, "rb") as f:
        return pickle.load(f, encoding="latin1")


def _np(a):
    return a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _f32_host(a):
    if isinstance(a, torch.Tensor):
        a = a.detach().to("cpu", torch.float32).numpy()
    return np.ascontiguousarray(a, dtype=np.float32)


class Renderer:
    def __init__(self, net, fine_net=None, cfg=None, canonical_vertex=None, device=None, faces=None):
        self.net = net
        self.cfg = cfg
        self.fine_net = fine_net
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", 0)) if torch.cuda.is_available() else 0
        self.device = torch.device("cuda", int(device))
        self.ctx = _lib.Context(int(device))  # raises without a B200: no fallback
        self.flags_extra = 0
        # optional early ray termination (DSNERF_EARLY_STOP, include/dsnerf.h): off by default, the default path evaluates every
        # non-transparent sample like the reference
        self.early_stop = False
        self.generator = None  # torch.Generator on the device for the training-mode draws (None = global generator)
        cv = canonical_vertex
        if cv is not None:
            cv = torch.as_tensor(cv, dtype=torch.float32).reshape(-1, 3)
        self.canonical_vertex = cv
        self.load_body_model(gender="neutral", body_model="smpl",
                             model_path=None if faces is not None else cfg.DATASETS.SMPL_PATH, faces=faces)
        self.sample_points_mode = cfg.MODEL.sample_points_mode
        self._weights_version = None
        if hasattr(net, "_point_evaluator"):
            net._point_evaluator = self._net_forward

    # ------------------------------------------------------------------ reference surface
    def train(self):
        self.net.training = True
        if self.fine_net is not None:
            self.fine_net.training = True

    def eval(self):
        self.net.training = False
        if self.fine_net is not None:
            self.fine_net.training = False

    def load_body_model(self, gender, body_model, model_path, faces=None):
        """can_render.py:382-406: faces, blend weights, canonical mesh."""
        if faces is None:
            tmp = _load_bodydata(model_path, body_model, gender)
            faces = np.asarray(tmp["f"]).astype(np.int64)
            self.smpl_blend_weight = torch.as_tensor(np.asarray(tmp["weights"], dtype=np.float32))[None].to(self.device)
            parents = torch.as_tensor(np.asarray(tmp["kintree_table"])[0].astype(np.int64))
            parents[0] = -1
            self.parents = parents
        faces = np.ascontiguousarray(faces, dtype=np.int64)
        self.face_idx = torch.from_numpy(faces).to(self.device)
        x_pose = np.zeros((1, 24, 3), dtype=np.float32)
        x_pose[:, 1, 2] += 0.6
        x_pose[:, 2, 2] -= 0.6
        self.x_pose = torch.from_numpy(x_pose).to(self.device)
        if self.canonical_vertex is not None:
            cv = self.canonical_vertex.to(self.device)
            self.canonical_model = {"vertex": cv, "meshes": cv[self.face_idx]}
            f32 = np.ascontiguousarray(faces, dtype=np.int32)
            v = _f32_host(self.canonical_vertex)
            self._n_verts = v.shape[0]
            self.ctx.check(self.ctx.L.dsnerf_set_mesh(self.ctx.h, f32.ctypes.data_as(ctypes.c_void_p), f32.shape[0],
                                                      v.ctypes.data_as(ctypes.c_void_p), v.shape[0]))

    # ------------------------------------------------------------------ state staging
    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _weights_fingerprint(self):
        """Changes whenever a parameter is replaced or edited in place, with no caller cooperation (mark_weights_dirty):
        ``Tensor._version`` counts in-place writes through the parameter itself (optimizer.step(), ``p.copy_()``,
        load_state_dict); writes through ``p.data`` bypass that counter, so the L2 norm of every tensor (one fused
        ``_foreach_norm``, ~0.2 ms for the 500 k parameters) is part of the fingerprint too."""
        sd = self.net.state_dict(keep_vars=True)
        ps = [sd[k] for k in STATE_DICT_ORDER]
        with torch.no_grad():
            norms = torch.stack(torch._foreach_norm([p.detach() for p in ps], 2)).tolist()
        return (getattr(self.net, "_weights_version", 0),) + tuple((p.data_ptr(), p._version, n) for p, n in zip(ps, norms))

    def _sync_weights(self):
        ver = self._weights_fingerprint()
        if ver == self._weights_version:
            return
        sd = self.net.state_dict()
        arrs = [_f32_host(sd[k]) for k in STATE_DICT_ORDER]
        ptrs = (ctypes.c_void_p * len(arrs))(*[a.ctypes.data_as(ctypes.c_void_p) for a in arrs])
        self.ctx.check(self.ctx.L.dsnerf_set_weights(self.ctx.h, ptrs, len(arrs)))
        self._weights_version = ver

    def _set_frame(self, batch):
        self._sync_weights()
        xyz = _f32_host(batch["xyz"]).reshape(-1, 3)
        if xyz.shape[0] != self._n_verts:
            raise ValueError("batch['xyz'] must hold one posed vertex per canonical vertex (batch size 1)")
        poses = _f32_host(batch["poses"]).reshape(-1, 3)
        if poses.shape[0] != 24:
            raise ValueError("batch['poses'] must be (1, 24, 3)")
        frame = int(torch.as_tensor(batch["frame"]).reshape(-1)[0])
        zero_code = 0 if getattr(self.net.nerf, "w", None) is None else 1
        shift = rot = rc = None
        if getattr(self.net, "light_center", None) is not None:
            th = _f32_host(batch["Th"]).reshape(-1, 3)
            shift = (_f32_host(self.net.light_center).reshape(-1)[:3] - th.mean(0)).astype(np.float32)
        if getattr(self.net, "rot_center", None) is not None and getattr(self.net, "rot", None) is not None:
            rot = _f32_host(self.net.rot).reshape(2, 2)
            rc = _f32_host(self.net.rot_center).reshape(-1)[:2].copy()
        p = lambda a: None if a is None else a.ctypes.data_as(ctypes.c_void_p)
        self.ctx.check(self.ctx.L.dsnerf_set_frame(self.ctx.h, p(xyz), p(poses), frame, zero_code, p(shift), p(rot), p(rc),
                                                   self._stream()))

    def _dev(self, t):
        return torch.as_tensor(t).to(self.device, torch.float32, non_blocking=True).contiguous()

    def _flags(self, mode=None):
        mode = mode or self.cfg.MODEL.sample_points_mode
        if mode not in ("GG", "uniform"):
            raise ValueError(f"unknown sample_points_mode {mode!r}")
        return (_lib.SAMPLE_GG if mode == "GG" else _lib.SAMPLE_UNIFORM) | self.flags_extra | (_lib.EARLY_STOP if self.early_stop else 0)

    def _check_eval(self):
        if getattr(self.net, "training", False):
            raise NotImplementedError("this entry point of dual_space_nerf_b200.Renderer works in eval mode only; call .eval() first "
                                      "(training mode is supported by render())")

    def _training_draws(self, R, n_samples):
        """The two random draws of a training-mode forward, on the device from ``self.generator`` (None = torch's global
        CUDA generator): torch.rand for the stratified jitter when cfg.MODEL.perturb > 0 (utils/pts_utils.py:6-13) and
        torch.randn * raw_noise_std when cfg.MODEL.raw_noise_std > 0 (utils/nerf_net_utils.py:29-33), else None."""
        if not getattr(self.net, "training", False):
            return None, None
        gen = getattr(self, "generator", None)
        jitter = noise = None
        if float(getattr(self.cfg.MODEL, "perturb", 0.0)) > 0.0:
            jitter = torch.rand(R, n_samples, device=self.device, generator=gen)
        std = float(getattr(self.cfg.MODEL, "raw_noise_std", 0.0))
        if std > 0.0:
            noise = torch.randn(R, n_samples, device=self.device, generator=gen) * std
        return jitter, noise

    def _render_rays_device(self, ray_o, ray_d, near, far, n_samples, want_weights, jitter=None, noise=None):
        R = ray_o.shape[0]
        mk = lambda *s: torch.empty(*s, device=self.device, dtype=torch.float32)
        rgb, depth, acc, disp = mk(R, 3), mk(R), mk(R), mk(R)
        weights = mk(R, n_samples) if want_weights else None
        z_vals = mk(R, n_samples) if want_weights else None
        if jitter is not None or noise is not None:
            self.ctx.check(self.ctx.L.dsnerf_render_train(self.ctx.h, _ptr(ray_o), _ptr(ray_d), _ptr(near), _ptr(far), R, n_samples,
                                                          self._flags(), _ptr(jitter), _ptr(noise), _ptr(rgb), _ptr(depth), _ptr(acc),
                                                          _ptr(disp), _ptr(weights), _ptr(z_vals), self._stream()))
        else:
            self.ctx.check(self.ctx.L.dsnerf_render(self.ctx.h, _ptr(ray_o), _ptr(ray_d), _ptr(near), _ptr(far), R, n_samples,
                                                    self._flags(), _ptr(rgb), _ptr(depth), _ptr(acc), _ptr(disp), _ptr(weights),
                                                    _ptr(z_vals), self._stream()))
        ret = {"color": rgb, "disp_map": disp, "acc_map": acc, "depth_map": depth}
        if want_weights:
            ret["weights"] = weights
            ret["z_vals"] = z_vals
        return ret

    def _render_fine(self, ray_o, ray_d, coarse, n_importance):
        if self.fine_net is not None and self.fine_net is not self.net:
            # the reference would evaluate the second pass with fine_net (can_render.py:232, a branch that cannot run there:
            # Renderer.resampling is undefined); one context stages one weight set, so a distinct fine network is refused
            raise NotImplementedError("hierarchical pass with a separate fine_net is not supported: pass fine_net=None "
                                      "(the second pass then uses net, as can_render.py:77-78 does)")
        R, N = coarse["z_vals"].shape
        mk = lambda *s: torch.empty(*s, device=self.device, dtype=torch.float32)
        z2 = mk(R, N + n_importance)
        self.ctx.check(self.ctx.L.dsnerf_resample(self.ctx.h, _ptr(coarse["z_vals"]), _ptr(coarse["weights"]), R, N, n_importance,
                                                  _ptr(z2), self._stream()))
        rgb, depth, acc, disp, w = mk(R, 3), mk(R), mk(R), mk(R), mk(R, N + n_importance)
        self.ctx.check(self.ctx.L.dsnerf_render_z(self.ctx.h, _ptr(ray_o), _ptr(ray_d), _ptr(z2), R, N + n_importance, self._flags(),
                                                  _ptr(rgb), _ptr(depth), _ptr(acc), _ptr(disp), _ptr(w), self._stream()))
        return {"color": rgb, "disp_map": disp, "acc_map": acc, "depth_map": depth, "weights": w, "z_vals": z2}

    def render(self, batch, jitter=None, noise=None):
        """can_render.py:137-168: {"coarse": {color, disp_map, acc_map, depth_map, weights, z_vals}} on the GPU.

        After ``train()`` this is the training-mode FORWARD (stratified jitter + density noise; the draws come from
        ``_training_draws`` unless given explicitly as (R,N) tensors).  No autograd graph is built: gradients for
        training stay with the caller's PyTorch model."""
        training = getattr(self.net, "training", False)
        if training and int(getattr(self.cfg.MODEL, "FINE_RAY_SAMPLING", -1)) > 0:
            raise NotImplementedError("the hierarchical second pass is an eval-mode feature (own spec, DESIGN.md)")
        with torch.cuda.device(self.device):
            self._set_frame(batch)
            ray_o = self._dev(batch["ray_o"]).reshape(-1, 3)
            ray_d = self._dev(batch["ray_d"]).reshape(-1, 3)
            near = self._dev(batch["near"]).reshape(-1)
            far = self._dev(batch["far"]).reshape(-1)
            N = int(self.cfg.MODEL.COARSE_RAY_SAMPLING)
            if training and jitter is None and noise is None:
                jitter, noise = self._training_draws(ray_o.shape[0], N)
            if jitter is not None:
                jitter = self._dev(jitter).reshape(-1, N)
            if noise is not None:
                noise = self._dev(noise).reshape(-1, N)
            coarse = self._render_rays_device(ray_o, ray_d, near, far, N, True, jitter, noise)
            batch["transparent_mask"] = self._last_transparent_mask(ray_o.shape[0], N)  # can_render.py:156
            out = {"coarse": coarse}
            n_imp = int(getattr(self.cfg.MODEL, "FINE_RAY_SAMPLING", -1))
            if n_imp > 0:
                out["fine"] = self._render_fine(ray_o, ray_d, coarse, n_imp)
        batch["canonical_model"] = self.canonical_model
        batch["face_idx"] = self.face_idx
        return out

    def _last_transparent_mask(self, R, N):
        """(R, N) bool mask of the samples of the render call that has just been enqueued (get_transparent_mask,
        utils/render_utils.py:103-109)."""
        tm = torch.empty(R, N, device=self.device, dtype=torch.uint8)
        if R > 0:
            self.ctx.check(self.ctx.L.dsnerf_last_transparent_mask(self.ctx.h, R, N, _ptr(tm), self._stream()))
        return tm.bool()

    def render_gather(self, batch, exchange, slot=0):
        """Multi-GPU form of ``render``: this rank's rays are rendered and the compositor kernel itself stores
        the per-ray outputs into slot ``slot`` of EVERY rank's frame buffer (``dist.FrameExchange``, peer-mapped symmetric
        memory over NVLink) -- the all-gather is part of the render kernel, not a collective after it.  Returns this GPU's
        (world, 6 * R) view of all ranks' blocks [rgb (R,3) | depth | acc | disp], complete on the current stream."""
        self._check_eval()
        with torch.cuda.device(self.device):
            self._set_frame(batch)
            ro = self._dev(batch["ray_o"]).reshape(-1, 3)
            rd = self._dev(batch["ray_d"]).reshape(-1, 3)
            ne = self._dev(batch["near"]).reshape(-1)
            fa = self._dev(batch["far"]).reshape(-1)
            if ro.shape[0] != exchange.R:
                raise ValueError(f"the exchange was built for {exchange.R} rays per rank, got {ro.shape[0]}")
            own, peers, n_peers, mc = exchange.targets(slot)
            self.ctx.check(self.ctx.L.dsnerf_render_gather(self.ctx.h, _ptr(ro), _ptr(rd), _ptr(ne), _ptr(fa), ro.shape[0],
                                                           int(self.cfg.MODEL.COARSE_RAY_SAMPLING), self._flags(), own, peers, n_peers, mc,
                                                           self._stream()))
            exchange.barrier()
        return exchange.frames(slot)

    def batchify_rays_view(self, ray_o, ray_d, near, far, batch, chunk=1024 * 32):
        """can_render.py:172-245.  ``chunk`` is accepted for signature compatibility; the whole
        batch is rendered by one library call (no intermediates scale with V x rays)."""
        self._check_eval()
        with torch.cuda.device(self.device):
            self._set_frame(batch)
            ro = self._dev(ray_o).reshape(-1, 3)
            rd = self._dev(ray_d).reshape(-1, 3)
            ne = self._dev(near).reshape(-1)
            fa = self._dev(far).reshape(-1)
            N = int(self.cfg.MODEL.COARSE_RAY_SAMPLING)
            n_imp = int(getattr(self.cfg.MODEL, "FINE_RAY_SAMPLING", -1))
            coarse = self._render_rays_device(ro, rd, ne, fa, N, True)
            fine = self._render_fine(ro, rd, coarse, n_imp) if n_imp > 0 else {}
        return coarse, fine

    def render_view(self, batch):
        """can_render.py:248-278: full-image outputs as CPU tensors of shape (H, W, C)."""
        coarse, fine = self.batchify_rays_view(batch["ray_o"], batch["ray_d"], batch["near"], batch["far"], batch)
        _, H, W, _ = batch["img"].shape
        mask = torch.as_tensor(batch["mask_at_box"])[0].to(self.device).bool()

        def post(src, C):  # utils/render_utils.py:466-472 post_process, on the GPU
            img = torch.zeros(H * W, C, device=self.device)
            img[mask] = src.reshape(-1, C)
            return img.reshape(H, W, C).cpu()

        out = {
            "coarse_color": post(coarse["color"], 3),
            "coarse_disp": post(coarse["disp_map"], 1),
            "coarse_acc": post(coarse["acc_map"], 1),
            "coarse_depth": post(coarse["depth_map"], 1),
        }
        if fine:
            out.update(fine_color=post(fine["color"], 3), fine_disp=post(fine["disp_map"], 1),
                       fine_acc=post(fine["acc_map"], 1), fine_depth=post(fine["depth_map"], 1))
        return out

    # ------------------------------------------------------------------ camera in, image out
    def camera_rays(self, H, W, K, R, T, bounds):
        """utils/rays_utils.py:16-30 get_rays + :63-97 get_near_far on the device, over all H*W pixels (row-major):
        returns ray_o, ray_d (H*W,3), near, far (H*W) float32 and mask_at_box (H*W) bool, all on this renderer's GPU."""
        K = np.ascontiguousarray(_np(K), np.float64).reshape(3, 3)
        R = np.ascontiguousarray(_np(R), np.float64).reshape(3, 3)
        T = np.ascontiguousarray(_np(T), np.float64).reshape(3)
        b = np.ascontiguousarray(_np(bounds), np.float32).reshape(6)
        P = int(H) * int(W)
        dev = self.device
        ro, rd = torch.empty(P, 3, device=dev), torch.empty(P, 3, device=dev)
        ne, fa = torch.empty(P, device=dev), torch.empty(P, device=dev)
        mk = torch.empty(P, device=dev, dtype=torch.uint8)
        hp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        with torch.cuda.device(dev):
            self.ctx.check(self.ctx.L.dsnerf_camera_rays(self.ctx.h, int(H), int(W), hp(K), hp(R), hp(T), hp(b), _ptr(ro), _ptr(rd), _ptr(ne),
                                                         _ptr(fa), _ptr(mk), self._stream()))
        return ro, rd, ne, fa, mk.bool()

    def render_view_camera(self, batch):
        """render_view for a batch that carries the camera instead of rays: `K`, `R`, `T`, `bounds` (2,3), `H`, `W` plus the
        usual `xyz`, `poses`, `frame` (`Th`).  Rays, near/far and mask_at_box (what the reference's dataloader computes with
        numpy, rays_utils.py:173-189) are generated on the GPU, the masked rays are rendered, and the per-ray outputs are
        scattered into (H, W, C) images as render_view does.  Saves the per-frame host ray generation and the 8 MB H2D."""
        H, W = int(batch["H"]), int(batch["W"])
        ro, rd, ne, fa, mask = self.camera_rays(H, W, batch["K"], batch["R"], batch["T"], batch["bounds"])
        b = dict(batch)
        b.update(ray_o=ro[mask][None], ray_d=rd[mask][None], near=ne[mask][None], far=fa[mask][None], mask_at_box=mask[None],
                 img=torch.zeros(1, H, W, 3))
        return self.render_view(b)

    # ------------------------------------------------------------------ secondary surface
    def w2l_without_lbs(self, pts_world, batch, canonical_model=None, ray_d_W=None, floor=-4, ceil=5, return_idx=False):
        """can_render.py:333-379.  Returns (pts_smpl_can (P,3), transparent_mask (B,P)[, ray_d_can]).
        ``ray_d_can`` is dead in the reference (SpaceNet.use_dir is False); zeros are returned for it."""
        if floor != -4 or ceil != 5:
            raise ValueError("the transparent-mask thresholds are fixed at the reference defaults (-4, 5, 0.1)")
        with torch.cuda.device(self.device):
            self._set_frame(batch)
            B = pts_world.shape[0]
            pts = self._dev(pts_world).reshape(-1, 3)
            P = pts.shape[0]
            cano = torch.empty(P, 3, device=self.device)
            tm = torch.empty(P, device=self.device, dtype=torch.uint8)
            idx = torch.empty(P, device=self.device, dtype=torch.int32)
            self.ctx.check(self.ctx.L.dsnerf_warp_points(self.ctx.h, _ptr(pts), P, _ptr(cano), _ptr(tm), _ptr(idx), self._stream()))
        res = [cano, tm.bool().reshape(B, -1)]
        if ray_d_W is not None:
            res.append(torch.zeros_like(cano))
        if return_idx:
            res.append(idx)
        return tuple(res)

    def w2l(self, pts_world, ray_o_W, ray_d_W, batch):
        """can_render.py:299-331."""
        B, ray, sp, _ = pts_world.shape
        rd = self._dev(ray_d_W).unsqueeze(2).expand(-1, -1, sp, -1).reshape(B, -1, 3)
        cano, tm, rdc = self.w2l_without_lbs(pts_world, batch, self.canonical_model, ray_d_W=rd)
        rays = torch.cat([rd.reshape(-1, 3), rdc], -1).reshape(B * ray, sp, 6)
        pts6 = torch.cat([self._dev(pts_world).reshape(B * ray, sp, 3), cano.reshape(B * ray, sp, 3)], -1)
        return pts6, rays, tm

    def query_volume(self, pts, code_idx, transparent_mask=None, batch_info={}):
        """can_render.py:280-296: points (B,P,6) = [world xyz | canonical xyz], as the reference's only caller passes them
        (utils/visualizer.py:60-66; DualSpaceNeRF.forward reads pos[..., 3:], model/spacenet.py:219), or canonical points
        (B,P,3); code_idx (B,) = one latent-code row per batch entry -> density (B,P,1), 0 where transparent_mask."""
        pts = torch.as_tensor(pts)
        if pts.dim() != 3 or pts.shape[-1] not in (3, 6):
            raise ValueError(f"query_volume expects pts of shape (B, P, 6) [world | canonical] or (B, P, 3) canonical, got {tuple(pts.shape)}")
        B, P = pts.shape[:2]
        codes = torch.as_tensor(code_idx).reshape(-1).tolist()
        if len(codes) != B:
            raise ValueError(f"code_idx must hold one code index per batch entry ({B}), got {len(codes)}")
        with torch.cuda.device(self.device):
            x = self._dev(pts)[..., -3:].contiguous()
            tm = None if transparent_mask is None else torch.as_tensor(transparent_mask).to(self.device).reshape(B, P).to(torch.uint8).contiguous()
            dens = torch.empty(B, P, device=self.device)
            for b in range(B):  # the latent code is a per-frame constant of the library: one frame set-up per distinct batch entry
                if "xyz" in batch_info:
                    bi = dict(batch_info)
                    bi["frame"] = torch.tensor([int(codes[b])])
                    self._set_frame(bi)
                elif b > 0 and codes[b] != codes[0]:
                    raise ValueError("different code indices per batch entry need batch_info (xyz, poses) to set up each frame")
                self.ctx.check(self.ctx.L.dsnerf_query_density(self.ctx.h, _ptr(x[b]), None if tm is None else _ptr(tm[b]), P,
                                                               _ptr(dens[b]), self.flags_extra, self._stream()))
        return dens.reshape(B, P, 1)

    def _net_forward(self, pos, rays, frame_idx, batch_info, density_only):
        """DualSpaceNeRF.forward (model/spacenet.py:210-266) on (P,6) points/rays."""
        with torch.cuda.device(self.device):
            if "xyz" in batch_info:
                self._set_frame(batch_info)
            pos = self._dev(pos).reshape(-1, 6)
            P = pos.shape[0]
            xw = pos[:, :3].contiguous()
            xc = pos[:, 3:].contiguous()
            dens = torch.empty(P, device=self.device)
            if density_only:
                self.ctx.check(self.ctx.L.dsnerf_query_density(self.ctx.h, _ptr(xc), None, P, _ptr(dens), self.flags_extra, self._stream()))
                return dens.reshape(P, 1)
            vd = self._dev(rays).reshape(-1, 6)[:, :3].contiguous()
            color = torch.empty(P, 3, device=self.device)
            self.ctx.check(self.ctx.L.dsnerf_eval_points(self.ctx.h, _ptr(xw), _ptr(xc), _ptr(vd), P, _ptr(color), _ptr(dens),
                                                         self.flags_extra, self._stream()))
        return color, dens.reshape(P, 1), None

    def render_rays(self, pts, rays, z_vals, frame_idx, net, transparent_mask=None, batch_info=None):
        """can_render.py:97-134 on explicit samples: pts/rays (R,N,6), z_vals (R,N); in training mode the density noise of
        raw2outputs is drawn here (``_training_draws``)."""
        with torch.cuda.device(self.device):
            pts = self._dev(pts)
            rays = self._dev(rays)
            z = self._dev(z_vals)
            R, N = z.shape
            color, dens, _ = self._net_forward(pts.reshape(-1, 6), rays.reshape(-1, 6), frame_idx, batch_info or {}, False)
            raw = torch.cat([color, dens], -1).reshape(R, N, 4)
            if transparent_mask is not None:
                raw[..., 3] = raw[..., 3].masked_fill(torch.as_tensor(transparent_mask).to(self.device).reshape(R, N), 0.0)
            raw = raw.contiguous()
            rd = rays[:, 0, :3].contiguous()
            mk = lambda *s: torch.empty(*s, device=self.device, dtype=torch.float32)
            rgb, depth, acc, disp, w = mk(R, 3), mk(R), mk(R), mk(R), mk(R, N)
            noise = self._training_draws(R, N)[1]
            self.ctx.check(self.ctx.L.dsnerf_composite_noise(self.ctx.h, _ptr(raw), _ptr(z), _ptr(rd), _ptr(noise), R, N, _ptr(rgb),
                                                             _ptr(depth), _ptr(acc), _ptr(disp), _ptr(w), self._stream()))
        return {"color": rgb, "disp_map": disp, "acc_map": acc, "depth_map": depth, "weights": w, "z_vals": z}

    def batchify_pts(self, pts, rays, z_vals, frame_idx, chunk=1024 * 32, net=None, batch_info=None):
        """can_render.py:65-95; no chunking needed."""
        tm = None if batch_info is None else batch_info.get("transparent_mask")
        return self.render_rays(pts, rays, z_vals, frame_idx, net, transparent_mask=tm, batch_info=batch_info)
