"""Manual GPU probe: frames alternating between two contexts on two streams (tail kernels of frame k next to the head kernels
of frame k + 1) against one context, device-resident, 512x512x64 (not a pytest file)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from types import SimpleNamespace
import bench as B
from dual_space_nerf_b200 import scene as S

args = SimpleNamespace(simt=False, early_stop=False)
sc = S.make_scene(B.H, B.W)
R = B.H * B.W
dev = torch.device("cuda", 0)
rigs, streams = [], []
for k in range(2):
    st = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(st):
        rig = B.Rig(args, 0, 0, 1)
        rig.stream, rig.sp = st, ctypes.c_void_p(st.cuda_stream)
        rig.set_mesh(sc)
        rig.sf = rig.frame_setter(sc, False)
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        rig.inp = [t(sc[k_]) for k_ in ("ray_o", "ray_d", "near", "far")]
        rig.out = [torch.empty(R, 3, device=dev), torch.empty(R, device=dev), torch.empty(R, device=dev), torch.empty(R, device=dev)]
    rigs.append(rig); streams.append(st)
torch.cuda.synchronize()
P = B.P

def frame(rig):
    with torch.cuda.stream(rig.stream):
        rig.flush.fill_(1)
        rig.sf()
        rig.ctx.check(rig.L.dsnerf_render(rig.ctx.h, P(rig.inp[0]), P(rig.inp[1]), P(rig.inp[2]), P(rig.inp[3]), R, B.N_SAMPLES, rig.flags,
                                          P(rig.out[0]), P(rig.out[1]), P(rig.out[2]), P(rig.out[3]), None, None, rig.sp))

def run(n, two):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(torch.cuda.current_stream())
    for s in streams: s.wait_stream(torch.cuda.current_stream())
    for i in range(n):
        frame(rigs[i & 1] if two else rigs[0])
    for s in streams: torch.cuda.current_stream().wait_stream(s)
    e1.record(torch.cuda.current_stream())
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

for two in (False, True, False, True):
    run(6, two)
    ms = run(30, two)
    print("two contexts" if two else "one context ", f"{ms:.3f} ms/frame  {R / ms / 1e3:.2f} M rays/s", flush=True)
a, b = rigs[0].out[0].cpu().numpy(), rigs[1].out[0].cpu().numpy()
print("frames identical across contexts:", bool(np.array_equal(a, b)), float(np.nansum(a)))
