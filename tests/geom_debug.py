"""Manual GPU probe: search statistics of the warp kernel (not a pytest file)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from types import SimpleNamespace
from dual_space_nerf_b200 import net as N, scene as S
from dual_space_nerf_b200.renderer import Renderer
H = int(sys.argv[1]) if len(sys.argv) > 1 else 512
sc = S.make_scene(H, H)
cfg = SimpleNamespace(MODEL=SimpleNamespace(TYPE="nerf", COARSE_RAY_SAMPLING=64, FINE_RAY_SAMPLING=-1, sample_points_mode="GG", perturb=1.0, raw_noise_std=1.0), DATASETS=SimpleNamespace(SMPL_PATH=None))
r = Renderer(N.synthetic_net(0), None, cfg, torch.from_numpy(sc["canonical"]), device=0, faces=sc["faces"])
r.eval()
r.ctx.profile(2)
out = r.render(S.to_batch(sc, torch))
torch.cuda.synchronize()
st = r.ctx.stats()
print(st)
print("searched/sample %.3f  cand/searched %.1f  evaluated/searched %.3f" % (st["searched_samples"] / st["samples"], st["nn_candidates"] / max(1, st["searched_samples"]), st["evaluated_samples"] / max(1, st["searched_samples"])))
