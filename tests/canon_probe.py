"""Probe: which path the canonical-space nearest-triangle lookups of a 512x512x64 frame take."""
import ctypes, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from dual_space_nerf_b200 import scene as S
from test_gpu_parity import make_renderer
H = int(os.environ.get("HW", 512))
sc = S.make_scene(H, H)
r = make_renderer(sc, 64)
r.ctx.L.dsnerf_profile(r.ctx.h, 2)
f32 = np.ascontiguousarray(sc["canonical"], np.float32); i32 = np.ascontiguousarray(sc["faces"], np.int32)
r.ctx.check(r.ctx.L.dsnerf_set_mesh(r.ctx.h, i32.ctypes.data_as(ctypes.c_void_p), len(i32), f32.ctypes.data_as(ctypes.c_void_p), len(f32)))
b = S.to_batch(sc, torch)
r.render(b)
st = r.ctx.stats()
for which in (0, 1):
    out = (ctypes.c_int * 16)()
    r.ctx.L.dsnerf_debug_table(r.ctx.h, which, out)
    print("posed" if which == 0 else "canon", list(out))
print(st)
