"""Manual GPU probe: outcome counters of the lazily built lookup tables (not a pytest file)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from types import SimpleNamespace
from dual_space_nerf_b200 import net as N, scene as S
from dual_space_nerf_b200.renderer import Renderer
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
sc = S.make_scene(n, n)
cfg = SimpleNamespace(MODEL=SimpleNamespace(TYPE="nerf", COARSE_RAY_SAMPLING=64, FINE_RAY_SAMPLING=-1, sample_points_mode="GG", perturb=1.0, raw_noise_std=1.0), DATASETS=SimpleNamespace(SMPL_PATH=None))
r = Renderer(N.synthetic_net(0), None, cfg, torch.from_numpy(sc["canonical"]), device=0, faces=sc["faces"])
r.eval()
r.ctx.profile(2)
faces = np.ascontiguousarray(sc["faces"], np.int32)
canon = np.ascontiguousarray(sc["canonical"], np.float32)
r.ctx.check(r.ctx.L.dsnerf_set_mesh(r.ctx.h, faces.ctypes.data_as(ctypes.c_void_p), faces.shape[0], canon.ctypes.data_as(ctypes.c_void_p), canon.shape[0]))  # rebuild the canonical grid with the counters on
r.render(S.to_batch(sc, torch)); torch.cuda.synchronize()
names = ["pool entries", "requested", "enum requested", "level-2 leftovers", "far|parent", "certified", "lists", "fallback", "list entries", "L1 visits", "L2 visits", "table cells", "enum cells"]
for which, nm in ((0, "posed"), (1, "canonical")):
    buf = (ctypes.c_int * 16)()
    r.ctx.check(r.ctx.L.dsnerf_debug_table(r.ctx.h, which, buf))
    print(nm, {k: buf[i] for i, k in enumerate(names)})
    print(nm, "lookups by path: list", buf[13], "scan", buf[14], "exhaustive / no search", buf[15])
print(r.ctx.stats())
