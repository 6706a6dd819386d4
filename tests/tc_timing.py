"""Manual GPU probe: per-op clock64 stamps of the tcgen05 kernel's first tile (not a pytest file)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from types import SimpleNamespace
from dual_space_nerf_b200 import net as N, scene as S
from dual_space_nerf_b200.renderer import Renderer
sc = S.make_scene(int(os.environ.get("DSNERF_TIMING_HW", "256")), int(os.environ.get("DSNERF_TIMING_HW", "256")))
cfg = SimpleNamespace(MODEL=SimpleNamespace(TYPE="nerf", COARSE_RAY_SAMPLING=64, FINE_RAY_SAMPLING=-1, sample_points_mode="GG", perturb=1.0, raw_noise_std=1.0), DATASETS=SimpleNamespace(SMPL_PATH=None))
r = Renderer(N.synthetic_net(0), None, cfg, torch.from_numpy(sc["canonical"]), device=0, faces=sc["faces"])
r.eval()
for _ in range(3): r.render(S.to_batch(sc, torch))
r.ctx.profile(4 | int(os.environ.get("DSNERF_DEBUG_PROFILE_BITS", "0")))
r.render(S.to_batch(sc, torch)); torch.cuda.synchronize()
buf = (ctypes.c_longlong * 128)()
r.ctx.check(r.ctx.L.dsnerf_debug_tc_timing(r.ctx.h, buf))
t = np.array(buf[:128], dtype=np.int64)
names = ["L0","L1","L2","L3","L4","L5","L6","rgb1","bW6","bW5","bW4","bW3","bW2","bW1","bW0"]
print("op     accwait  q0+q1  q2+q3 (cycles)")
prev = t[0]
for op in range(15):
    a0, a1, a3 = t[1 + 4 * op], t[2 + 4 * op], t[3 + 4 * op]
    if op == 14:
        print(f"{names[op]:6s} (consumed by the tile-output stage)")
        break
    print(f"{names[op]:6s} {a0 - prev:8d} {a1 - a0:8d} {a3 - a1:8d}")
    prev = a3
print("tile total", t[62] - t[0])
print("MMA thread per op: wait_full  wait_a  total_issue_loop")
for op in range(15):
    print(f"{names[op]:6s} {t[64+3*op]:9d} {t[65+3*op]:9d} {t[66+3*op]:9d}")
dc, dt = t[122] - t[120], t[123] - t[121]
print(f"kernel (CTA 0): {dc} cycles in {dt / 1e3:.1f} us -> SM clock inside the kernel {dc / max(dt, 1) * 1e3:.0f} MHz")
