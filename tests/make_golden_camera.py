"""Generate tests/golden/camera_rays.npz by running the UNMODIFIED reference functions utils/rays_utils.py:get_rays (:16-30)
and get_near_far (:63-97) on CPU -- the inference branch of my_sample_ray (:173-189).

Run HERE (needs /root/reference and cv2):  python tests/make_golden_camera.py
"""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))


def cases():
    from dual_space_nerf_b200 import scene as S

    _, posed, _ = S.body_meshes()
    bounds = np.stack([posed.min(0), posed.max(0)]).astype(np.float32)
    bounds[0, 2] -= 0.05
    bounds[1, 2] += 0.05
    out = []
    K, R, T = S.camera(48, 64)
    out.append((48, 64, K, R, T, bounds))
    # an oblique camera with a non-trivial K (skew, off-centre principal point), as calibrated rigs have
    th = 0.4
    Ry = np.array([[np.cos(th), 0, np.sin(th)], [0, 1, 0], [-np.sin(th), 0, np.cos(th)]])
    R2 = Ry @ R
    C2 = np.array([1.3, -2.6, 1.4])
    K2 = np.array([[70.5, 0.3, 30.2], [0, 69.1, 25.7], [0, 0, 1.0]])
    out.append((50, 60, K2, R2, (-R2 @ C2), bounds))
    return out


def main():
    spec = importlib.util.spec_from_file_location("ref_rays_utils", "/root/reference/utils/rays_utils.py")
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    save = {}
    for c, (H, W, K, R, T, bounds) in enumerate(cases()):
        ray_o, ray_d = m.get_rays(H, W, K, R, T.reshape(3, 1))
        ray_o = ray_o.reshape(-1, 3).astype(np.float32)
        ray_d = ray_d.reshape(-1, 3).astype(np.float32)
        near, far, mask = m.get_near_far(bounds, ray_o, ray_d)
        save.update({f"H{c}": H, f"W{c}": W, f"K{c}": K, f"R{c}": R, f"T{c}": T, f"bounds{c}": bounds, f"ray_o{c}": ray_o[0],
                     f"ray_d{c}": ray_d, f"near{c}": near.astype(np.float32), f"far{c}": far.astype(np.float32), f"mask{c}": mask})
        print(c, H, W, int(mask.sum()), "rays hit the box")
    np.savez_compressed(os.path.join(HERE, "golden", "camera_rays.npz"), **save)


if __name__ == "__main__":
    main()
