"""Manual GPU probe (not a pytest file): BASELINE configs[3] shape, 1024x1024 rays x 64 samples, on one GPU."""
import sys, time, numpy as np, torch
sys.path.insert(0,'/root/repo')
from types import SimpleNamespace
from dual_space_nerf_b200 import net as N, scene as S
from dual_space_nerf_b200.renderer import Renderer
sc = S.make_scene(1024, 1024)
cfg = SimpleNamespace(MODEL=SimpleNamespace(TYPE="nerf", COARSE_RAY_SAMPLING=64, FINE_RAY_SAMPLING=-1, sample_points_mode="GG", perturb=1.0, raw_noise_std=1.0), DATASETS=SimpleNamespace(SMPL_PATH=None))
r = Renderer(N.synthetic_net(0), None, cfg, torch.from_numpy(sc["canonical"]), device=0, faces=sc["faces"]); r.eval()
b = S.to_batch(sc, torch, device="cuda")
for i in range(3):
    torch.cuda.synchronize(); t=time.time(); out = r.render(b)["coarse"]; torch.cuda.synchronize(); dt=time.time()-t
    print("1024x1024x64:", dt*1e3, "ms", sc["ray_o"].shape[0]/dt/1e6, "M rays/s", r.ctx.stats()["evaluated_samples"])
print("acc mean", float(out["acc_map"].mean()), "finite", bool(torch.isfinite(out["color"]).all()))
