"""Manual GPU check of the tcgen05 MLP kernel against the fp32 SIMT kernel (not a pytest file)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, ctypes
from types import SimpleNamespace
from dual_space_nerf_b200 import net as N, scene as S, lib
from dual_space_nerf_b200.renderer import Renderer

sc = S.make_scene(64, 64)
cfg = SimpleNamespace(MODEL=SimpleNamespace(TYPE="nerf", COARSE_RAY_SAMPLING=32, FINE_RAY_SAMPLING=-1, sample_points_mode="GG", perturb=1.0, raw_noise_std=1.0), DATASETS=SimpleNamespace(SMPL_PATH=None))
r = Renderer(N.synthetic_net(0), None, cfg, torch.from_numpy(sc["canonical"]), device=0, faces=sc["faces"])
r.eval()
batch = S.to_batch(sc, torch)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
rng = np.random.RandomState(0)
# canonical points near the canonical surface
vid = rng.randint(0, sc["canonical"].shape[0], n)
xc = (sc["canonical"][vid] + rng.randn(n, 3).astype(np.float32) * 0.02).astype(np.float32)
xw = (xc + np.array([0.2, -0.1, 1.0], np.float32)).astype(np.float32)
vd = rng.randn(n, 3).astype(np.float32)
pts = torch.from_numpy(xc)[None]
res = {}
for name, fl in (("simt", lib.MLP_FP32_SIMT), ("tc", 0)):
    r.flags_extra = fl
    d = r.query_volume(pts, torch.tensor([5]), None, batch)
    torch.cuda.synchronize()
    res[name + "_d"] = d.cpu().numpy().ravel()
    print(name, "density-only ok", res[name + "_d"][:4])
print("density-only max diff", np.abs(res["simt_d"] - res["tc_d"]).max(), "scale", np.abs(res["simt_d"]).max())
pos = torch.from_numpy(np.concatenate([xw, xc], 1))
rays = torch.from_numpy(np.concatenate([vd, vd], 1))
for name, fl in (("simt", lib.MLP_FP32_SIMT), ("tc", 0)):
    r.flags_extra = fl
    c, d, _ = r._net_forward(pos, rays, None, batch, False)
    torch.cuda.synchronize()
    res[name + "_c"] = c.cpu().numpy(); res[name + "_s"] = d.cpu().numpy().ravel()
    print(name, "full ok", res[name + "_c"][:2])
print("sigma max diff", np.abs(res["simt_s"] - res["tc_s"]).max())
cd = np.abs(res["simt_c"] - res["tc_c"]).max(1)
print("color max diff", cd.max(), "p99", np.percentile(cd, 99), "median", np.median(cd))
