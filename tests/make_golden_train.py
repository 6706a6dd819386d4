"""Generate tests/golden/render_train.npz by running the UNMODIFIED reference's Renderer.render in TRAINING mode on CPU
(stratified jitter + density noise).  Run HERE (needs /root/reference):
    python tests/make_golden_train.py
The random draws are inputs: see ReferenceRig.render_train.  96 rays x 32 samples of the 64x64 scene, GG sampling."""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

import refharness as RH  # noqa: E402
from dual_space_nerf_b200 import net as N  # noqa: E402
from dual_space_nerf_b200 import scene as S  # noqa: E402
from make_golden import spread_rays  # noqa: E402


def main():
    sd = N.synthetic_net(0).state_dict()
    sc = S.make_scene(64, 64)
    n = 32
    rays = spread_rays(sc, 80, 16, seed=5)
    rng = np.random.RandomState(11)
    t_rand = rng.rand(len(rays), n).astype(np.float32)
    noise = rng.randn(len(rays), n).astype(np.float32)  # raw_noise_std = 1
    rig = RH.ReferenceRig(sc, n, sd)
    out = rig.render_train(rays, t_rand, noise)
    jit = rig.render_train(rays, t_rand, np.zeros_like(noise))  # jitter only
    np.savez_compressed(os.path.join(HERE, "golden", "render_train.npz"), rays=rays, t_rand=t_rand, noise=noise,
                        **out, **{"jitter_only_" + k: v for k, v in jit.items()})
    print({k: v.shape for k, v in out.items()}, float(out["acc_map"].mean()), float(jit["acc_map"].mean()))


if __name__ == "__main__":
    main()
