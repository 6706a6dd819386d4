"""Host-side weight container mirroring ``model/spacenet.py`` of the reference.

The reference's callers reach into ``Renderer.net`` (``render.net.load_state_dict``,
``render.net.set_light_center``, ``render.net.nerf.w = 0`` -- validate.py:27,
test.py:193-196), so the drop-in keeps a ``DualSpaceNeRF`` object with the very
same ``state_dict()`` layout (33 tensors, 500 021 parameters):

    nerf.embedding.weight (500,8)
    nerf.stage1.{0,2,4,6}   87->256->256->256->256      (spacenet.py:47-56)
    nerf.stage2.{0,2,4}     319->256->256->256           (spacenet.py:58-67)
    nerf.density_net.0      256->1                       (spacenet.py:69-73)
    nerf.rgb_net.{1,3}      256->128->3                  (spacenet.py:75-80)
    lighting_mlp.lights_encoding.{0,2,4}  9->128->128->1 (spacenet.py:165-172)
    pose_mlp.{0,2,4}        92->64->64->16               (spacenet.py:199-205)

It holds parameters only.  There is deliberately **no torch forward pass**:
evaluation happens in the CUDA library (``libdsnerf.so``), and ``forward`` routes
there or raises -- there is no CPU fallback.
"""
from __future__ import annotations

import torch
from torch import nn

MAX_FRAME = 500
CODE_DIM = 8
POSE_FEAT = 16
PE_L = 10
PE_DIM = 3 + 3 * 2 * PE_L  # 63
WIDTH = 256
HEAD = 128
LIGHT_W = 128


def _act():
    # parameter-free placeholder keeping nn.Sequential indices equal to the reference's
    return nn.Identity()


class SpaceNet(nn.Module):
    """Parameter layout of the canonical-space density/essence network."""

    def __init__(self, maxFrame=MAX_FRAME, code_dim=CODE_DIM, essence_dim=3, cfg=None):
        super().__init__()
        self.cfg = cfg
        self.code_dim = code_dim
        self.use_dir = False
        self.embedding = nn.Embedding(maxFrame, code_dim)
        in_dim = PE_DIM + code_dim + POSE_FEAT
        self.stage1 = nn.Sequential(
            nn.Linear(in_dim, WIDTH), _act(), nn.Linear(WIDTH, WIDTH), _act(),
            nn.Linear(WIDTH, WIDTH), _act(), nn.Linear(WIDTH, WIDTH), _act(),
        )
        self.stage2 = nn.Sequential(
            nn.Linear(WIDTH + PE_DIM, WIDTH), _act(), nn.Linear(WIDTH, WIDTH), _act(),
            nn.Linear(WIDTH, WIDTH), _act(),
        )
        self.density_net = nn.Sequential(nn.Linear(WIDTH, 1))
        self.rgb_net = nn.Sequential(_act(), nn.Linear(WIDTH, HEAD), _act(), nn.Linear(HEAD, essence_dim))
        # reference: ``self.w is not None`` zeroes the per-frame code (spacenet.py:126-129)
        self.w = None


class LightingMLP(nn.Module):
    def __init__(self, essence_dim=3):
        super().__init__()
        self.lights_encoding = nn.Sequential(
            nn.Linear(9, LIGHT_W), _act(), nn.Linear(LIGHT_W, LIGHT_W), _act(), nn.Linear(LIGHT_W, 1), _act(),
        )


class DualSpaceNeRF(nn.Module):
    """Weights + render-time switches of the reference model (spacenet.py:191-275)."""

    def __init__(self, cfg=None):
        super().__init__()
        self.nerf = SpaceNet(essence_dim=3, cfg=cfg)
        self.lighting_mlp = LightingMLP(essence_dim=3)
        self.pose_mlp = nn.Sequential(
            nn.Linear(23 * 4, 64), _act(), nn.Linear(64, 64), _act(), nn.Linear(64, POSE_FEAT),
        )
        self.light_center = None
        self.rot_center = None
        self.rot = None
        self._weights_version = 0
        self._point_evaluator = None  # set by Renderer: routes forward() to the CUDA library

    # -- reference setters (spacenet.py:268-275); tensors stay on the host, the
    #    library receives them per frame --------------------------------------
    def set_rot_center(self, center):
        self.rot_center = torch.as_tensor(center, dtype=torch.float32).detach().cpu()

    def set_rot(self, rot):
        self.rot = torch.as_tensor(rot, dtype=torch.float32).detach().cpu()

    def set_light_center(self, center):
        self.light_center = torch.as_tensor(center, dtype=torch.float32).detach().cpu()

    def load_state_dict(self, state_dict, strict=True, **kw):
        out = super().load_state_dict(state_dict, strict=strict, **kw)
        self._weights_version += 1
        return out

    def load_checkpoint(self, path, strict=True):
        """Load a checkpoint file written by the reference's Checkpointer (utils/checkpoint.py:102-125: a dict whose
        "model" entry is DualSpaceNeRF.state_dict(); validate.py:27 / test.py:192 do
        `render.net.load_state_dict(torch.load(path)["model"])`).  A bare state_dict file is accepted too.  A "module."
        prefix left by DataParallel wrappers is stripped."""
        data = torch.load(path, map_location="cpu", weights_only=False)
        sd = data["model"] if isinstance(data, dict) and "model" in data else data
        sd = {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items()}
        self.load_state_dict(sd, strict=strict)
        return data

    def mark_weights_dirty(self):
        """Call after editing parameters in place so the library re-stages them."""
        self._weights_version += 1

    def cuda(self, device=None):  # weights are staged by the library; nothing to move
        return self

    def forward(self, pos, rays, frame_idx=0, batch_info={}, density_only=False):
        if self._point_evaluator is None:
            raise RuntimeError(
                "DualSpaceNeRF has no torch forward pass; attach it to a dual_space_nerf_b200.Renderer "
                "(CUDA library) first"
            )
        return self._point_evaluator(pos, rays, frame_idx, batch_info, density_only)


STATE_DICT_ORDER = [
    "nerf.embedding.weight",
    "nerf.stage1.0.weight", "nerf.stage1.0.bias", "nerf.stage1.2.weight", "nerf.stage1.2.bias",
    "nerf.stage1.4.weight", "nerf.stage1.4.bias", "nerf.stage1.6.weight", "nerf.stage1.6.bias",
    "nerf.stage2.0.weight", "nerf.stage2.0.bias", "nerf.stage2.2.weight", "nerf.stage2.2.bias",
    "nerf.stage2.4.weight", "nerf.stage2.4.bias",
    "nerf.density_net.0.weight", "nerf.density_net.0.bias",
    "nerf.rgb_net.1.weight", "nerf.rgb_net.1.bias", "nerf.rgb_net.3.weight", "nerf.rgb_net.3.bias",
    "lighting_mlp.lights_encoding.0.weight", "lighting_mlp.lights_encoding.0.bias",
    "lighting_mlp.lights_encoding.2.weight", "lighting_mlp.lights_encoding.2.bias",
    "lighting_mlp.lights_encoding.4.weight", "lighting_mlp.lights_encoding.4.bias",
    "pose_mlp.0.weight", "pose_mlp.0.bias", "pose_mlp.2.weight", "pose_mlp.2.bias",
    "pose_mlp.4.weight", "pose_mlp.4.bias",
]


def synthetic_head_rescale_(net):
    """default init yields density < 0 everywhere (image == 0), so
    the synthetic scene rescales the two heads in place."""
    with torch.no_grad():
        net.nerf.density_net[0].weight.mul_(4000.0)
        net.nerf.density_net[0].bias.fill_(120.0)
        net.nerf.rgb_net[3].weight.mul_(4.0)
        net.nerf.rgb_net[3].bias.fill_(0.5)
    if hasattr(net, "mark_weights_dirty"):
        net.mark_weights_dirty()
    return net


def synthetic_net(seed=0):
    """Random-init weights of the reference architecture (torch default init, seed 0)."""
    torch.manual_seed(seed)
    net = DualSpaceNeRF(None)
    return synthetic_head_rescale_(net)
