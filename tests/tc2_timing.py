"""Manual GPU probe: clock64 stamps of mlp_tc2_kernel's first tile pair on CTA 0 (not a pytest file)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from types import SimpleNamespace
from dual_space_nerf_b200 import net as N, scene as S
from dual_space_nerf_b200.renderer import Renderer
hw = int(os.environ.get("DSNERF_TIMING_HW", "512"))
sc = S.make_scene(hw, hw)
cfg = SimpleNamespace(MODEL=SimpleNamespace(TYPE="nerf", COARSE_RAY_SAMPLING=64, FINE_RAY_SAMPLING=-1, sample_points_mode="GG", perturb=1.0, raw_noise_std=1.0), DATASETS=SimpleNamespace(SMPL_PATH=None))
r = Renderer(N.synthetic_net(0), None, cfg, torch.from_numpy(sc["canonical"]), device=0, faces=sc["faces"])
r.eval()
for _ in range(3): r.render(S.to_batch(sc, torch))
r.ctx.profile(4 | int(os.environ.get("DSNERF_DEBUG_PROFILE_BITS", "0")))
r.render(S.to_batch(sc, torch)); torch.cuda.synchronize()
buf = (ctypes.c_longlong * 128)()
r.ctx.check(r.ctx.L.dsnerf_debug_tc_timing(r.ctx.h, buf))
t = np.array(buf[:128], dtype=np.int64)
names = ["L0","L1","L2","L3","L4","L5","L6","rgb","bW6","bW5","E4","bW4","bW3","bW2","bW1","bW0"]
print("phase   start(rel)  accwait->   work   (cycles; start = when the accumulator was seen, relative to the tile pair's start)")
prev_end = t[0]
for op in range(16):
    for s in range(2):
        a, b = t[2 + 2 * (2 * op + s)], t[3 + 2 * (2 * op + s)]
        if a == 0: continue
        print(f"{names[op]:4s} s{s}  {a - t[0]:9d}  wait {a - prev_end:7d}  work {b - a:7d}")
        prev_end = b
print("MMA warp per phase: start (accumulator free seen, rel), issue loop, -> accumulator seen by the epilogue after issue end")
t0lo = int(t[0]) & 0xffffffff
for op in range(16):
    for s in range(2):
        e = int(t[66 + 2 * op + s])
        if e == 0: continue
        st, en = e & 0xffffffff, (e >> 32) & 0xffffffff
        a = int(t[2 + 2 * (2 * op + s)])
        rel = (st - t0lo) & 0xffffffff
        rel = rel - (1 << 32) if rel > (1 << 31) else rel
        print(f"{names[op]:4s} s{s}  start {rel:9d}  issue {(en - st) & 0xffffffff:6d}  acc seen {a - t[0] - rel - ((en - st) & 0xffffffff) if a else 0:6d} after issue end")
print("cycles the MMA warp waited for weight slabs, tile slot 0, ops 0..15:", [int(x) for x in t[98:114]])
print("pair total", t[1] - t[0])
dc, dt = t[122] - t[120], t[123] - t[121]
print(f"kernel (CTA 0): {dc} cycles in {dt / 1e3:.1f} us -> SM clock inside the kernel {dc / max(dt, 1) * 1e3:.0f} MHz")
