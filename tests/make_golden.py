"""Generate tests/golden/*.npz by running the UNMODIFIED reference on CPU.

Run HERE (needs /root/reference):  python tests/make_golden.py
The GPU box has no /root/reference, so the vectors are committed.  Inputs are
regenerated from dual_space_nerf_b200.scene / .net by the tests (deterministic),
only ray selections and reference outputs are stored.

Files
  stages_64x64x32.npz   stage-boundary tensors of Renderer.render (eval) on a
                         spread of 160 rays of the 64x64 scene, N=32, GG sampling
  render_64x64x32.npz   Renderer.render_view full 64x64 frame, N=32 (config 1)
  render_uniform.npz    Renderer.render, "uniform" sampling, 128 rays, N=16
  render_novelpose.npz  Renderer.render with nerf.w=0 and set_light_center(...)
                         (test.py:193-196), 128 rays, N=32, pose seed 3
  render_128x128x64.npz Renderer.render on 512 rays of a 128x128 scene, N=64
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

import refharness as RH  # noqa: E402
from dual_space_nerf_b200 import net as N  # noqa: E402
from dual_space_nerf_b200 import scene as S  # noqa: E402

OUT = os.path.join(HERE, "golden")


def spread_rays(scene, n_hit, n_miss, seed=0):
    rng = np.random.RandomState(seed)
    hit = np.nonzero(scene["hit_box"])[0]
    miss = np.nonzero(~scene["hit_box"])[0]
    sel = np.concatenate([rng.choice(hit, n_hit, replace=False), rng.choice(miss, n_miss, replace=False)])
    return np.sort(sel).astype(np.int64)


def main():
    os.makedirs(OUT, exist_ok=True)
    sd = N.synthetic_net(0).state_dict()

    sc = S.make_scene(64, 64)
    rig = RH.ReferenceRig(sc, 32, sd)
    rays = spread_rays(sc, 128, 32)
    t = time.time()
    st = rig.stages(rays)
    keep = ["near_gg", "far_gg", "z_vals", "pts", "idx", "uv", "h", "mask", "xyz_cano", "pose_feat",
            "essence", "density", "grad", "normal_world", "color"]
    np.savez_compressed(os.path.join(OUT, "stages_64x64x32.npz"), rays=rays, **{k: st[k] for k in keep})
    print("stages", time.time() - t)

    t = time.time()
    rv = rig.render_view()
    np.savez_compressed(os.path.join(OUT, "render_64x64x32.npz"), **rv)
    print("render_view 64x64x32", time.time() - t, "s ->", 4096 / (time.time() - t), "rays/s")

    rigu = RH.ReferenceRig(sc, 16, sd, mode="uniform")
    rays = spread_rays(sc, 112, 16, seed=1)
    out = rigu.render(rays)
    np.savez_compressed(os.path.join(OUT, "render_uniform.npz"), rays=rays, **out)

    sc3 = S.make_scene(64, 64, pose_seed=3)
    rign = RH.ReferenceRig(sc3, 32, sd)
    rign.net.nerf.w = 0
    rign.net.set_light_center(torch.from_numpy(S.LIGHT_CENTER_313))
    rays = spread_rays(sc3, 112, 16, seed=2)
    out = rign.render(rays)
    np.savez_compressed(os.path.join(OUT, "render_novelpose.npz"), rays=rays, **out)

    sc128 = S.make_scene(128, 128)
    rig128 = RH.ReferenceRig(sc128, 64, sd)
    rays = spread_rays(sc128, 448, 64, seed=3)
    out = rig128.render(rays)
    np.savez_compressed(os.path.join(OUT, "render_128x128x64.npz"), rays=rays, **out)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
