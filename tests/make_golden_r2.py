"""Generate the round-2 goldens by running the UNMODIFIED reference on CPU.  Run HERE (needs /root/reference):
    python tests/make_golden_r2.py

  render_rot.npz    Renderer.render with net.set_rot_center / net.set_rot (vis_lighting.py:57-58, model/spacenet.py:254-258)
                    plus set_light_center, 128 rays x 32 samples of the 64x64 scene, pose seed 2
  render_bigw.npz   Renderer.render with hidden-layer weights scaled to trained-checkpoint magnitudes (gradients that GROW
                    through the backward chain instead of shrinking as with default init), 112 rays x 32 samples
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

import refharness as RH  # noqa: E402
from dual_space_nerf_b200 import net as N  # noqa: E402
from dual_space_nerf_b200 import scene as S  # noqa: E402
from make_golden import spread_rays  # noqa: E402

ROT_CENTER = np.array([[0.18649693, -0.14180326, 1.7103844]], np.float32)  # vis_lighting.py:57 ("head 313")
ROT_ANGLE = 72.0


def angle2rot(angle):  # vis_lighting.py:86-91
    r = np.pi * angle / 180
    return np.array([[np.cos(r), -np.sin(r)], [np.sin(r), np.cos(r)]])


def big_weight_net(seed=0, hidden=2.6):
    """Synthetic net whose hidden layers carry a per-layer rms gain > 1 (default init: 0.41), as trained checkpoints do:
    activations and chained gradients grow with depth.  The density head is scaled back so that the image stays non-trivial."""
    net = N.synthetic_net(seed)
    with torch.no_grad():
        for seq, idxs in ((net.nerf.stage1, (2, 4, 6)), (net.nerf.stage2, (0, 2, 4))):
            for i in idxs:
                seq[i].weight.mul_(hidden)
        net.nerf.density_net[0].weight.div_(hidden ** 6 / 3.0)
    net.mark_weights_dirty()
    return net


def main():
    out_dir = os.path.join(HERE, "golden")
    sd = N.synthetic_net(0).state_dict()

    sc = S.make_scene(64, 64, pose_seed=2)
    rig = RH.ReferenceRig(sc, 32, sd)
    rig.net.set_light_center(torch.from_numpy(S.LIGHT_CENTER_313))
    rig.net.set_rot_center(torch.Tensor(ROT_CENTER))
    rig.net.set_rot(torch.Tensor(angle2rot(ROT_ANGLE)))
    rays = spread_rays(sc, 112, 16, seed=7)
    out = rig.render(rays)
    rig2 = RH.ReferenceRig(sc, 32, sd)
    rig2.net.set_light_center(torch.from_numpy(S.LIGHT_CENTER_313))
    base = rig2.render(rays)
    print("rot changes colour by", float(np.abs(out["color"] - base["color"]).max()))
    np.savez_compressed(os.path.join(out_dir, "render_rot.npz"), rays=rays, **out)

    sc = S.make_scene(64, 64)
    sdb = big_weight_net(0).state_dict()
    rig = RH.ReferenceRig(sc, 32, sdb)
    rays = spread_rays(sc, 96, 16, seed=9)
    out = rig.render(rays)
    print("bigw acc mean", float(out["acc_map"].mean()), "colour max", float(out["color"].max()))
    np.savez_compressed(os.path.join(out_dir, "render_bigw.npz"), rays=rays, **out)


if __name__ == "__main__":
    main()
