"""Manual GPU check of the two-tiles-in-flight tcgen05 kernel (mlp_tc2.cuh) against the one-tile kernel and the fp32 SIMT
kernel, point-wise through dsnerf_eval_points (not a pytest file).  usage: tc2_debug.py [n_points ...]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("DSNERF_TC_WATCHDOG", "1")
import numpy as np, torch
from types import SimpleNamespace
from dual_space_nerf_b200 import net as N, scene as S, lib
from dual_space_nerf_b200.renderer import Renderer

sc = S.make_scene(64, 64)
cfg = SimpleNamespace(MODEL=SimpleNamespace(TYPE="nerf", COARSE_RAY_SAMPLING=32, FINE_RAY_SAMPLING=-1, sample_points_mode="GG", perturb=1.0, raw_noise_std=1.0), DATASETS=SimpleNamespace(SMPL_PATH=None))
def make(variant):
    os.environ["DSNERF_MLP_VARIANT"] = str(variant)
    r = Renderer(N.synthetic_net(0), None, cfg, torch.from_numpy(sc["canonical"]), device=0, faces=sc["faces"])
    r.eval()
    return r
r1, r2 = make(1), make(2)
batch = S.to_batch(sc, torch)
for n in [int(a) for a in sys.argv[1:]] or [1000, 40000, 300000]:
    rng = np.random.RandomState(n)
    vid = rng.randint(0, sc["canonical"].shape[0], n)
    xc = (sc["canonical"][vid] + rng.randn(n, 3).astype(np.float32) * 0.02).astype(np.float32)
    xw = (xc + np.array([0.2, -0.1, 1.0], np.float32)).astype(np.float32)
    vd = rng.randn(n, 3).astype(np.float32)
    pos = torch.from_numpy(np.concatenate([xw, xc], 1))
    rays = torch.from_numpy(np.concatenate([vd, vd], 1))
    res = {}
    for name, r, fl in (("simt", r1, lib.MLP_FP32_SIMT), ("tc1", r1, 0), ("tc2", r2, 0), ("tc2b", r2, 0)):
        r.flags_extra = fl
        c, d, _ = r._net_forward(pos, rays, None, batch, False)
        torch.cuda.synchronize()
        res[name] = (c.cpu().numpy(), d.cpu().numpy().ravel())
    for a, b in (("tc1", "simt"), ("tc2", "simt"), ("tc2", "tc1"), ("tc2", "tc2b")):
        ds = np.abs(res[a][1] - res[b][1]); dc = np.abs(res[a][0] - res[b][0]).max(1)
        print(f"n={n:7d} {a:4s} vs {b:4s}: sigma max {ds.max():.3e} (bad rows {np.flatnonzero(ds > 1e-2)[:6]})  color max {dc.max():.3e} p99 {np.percentile(dc, 99):.3e}"
              f"  nan {int(np.isnan(res[a][0]).sum() + np.isnan(res[a][1]).sum())}", flush=True)
