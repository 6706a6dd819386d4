"""Manual GPU probe (not a pytest file): per-sample errors of the tcgen05 path on the trained-magnitude weights of
tests/make_golden_r2.py against the oracle and against the fp32 SIMT kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np, torch
import common as C
from dual_space_nerf_b200 import net as N, scene as S, lib
from oracle import oracle as O
from make_golden_r2 import big_weight_net
from test_gpu_parity import make_renderer, oracle_run, to_np

g = C.golden("render_bigw.npz")
rays = g["rays"]
sc = S.make_scene(64, 64)
net = big_weight_net(0)
n = 32
ref, st = oracle_run(sc, net.state_dict(), n, rays)
act = ~st["mask"]
kink_s = st["kink_margin"] < C.KINK_MARGIN
b = S.to_batch(sc, torch, rays=rays)
pos = torch.from_numpy(np.concatenate([st["pts"].reshape(-1, 3), st["xyz_cano"]], 1))
rd = torch.from_numpy(np.concatenate([np.repeat(sc["ray_d"][rays], n, 0)] * 2, 1))
for name, fl in (("tc", 0), ("simt", lib.MLP_FP32_SIMT)):
    r = make_renderer(sc, n, net=net)
    r.flags_extra = fl
    out = to_np(r.render(S.to_batch(sc, torch, rays=rays))["coarse"])
    print(name, "mode", r.ctx.L.dsnerf_tensor_path_active(r.ctx.h), "ray rgb err max", np.abs(out["color"] - g["color"]).max(),
          "depth", np.abs(out["depth_map"] - g["depth_map"]).max())
    b2 = dict(b); b2["canonical_model"], b2["face_idx"] = r.canonical_model, r.face_idx
    c, d, _ = r._net_forward(pos, rd, None, b2, False)
    c, d = c.cpu().numpy(), d.cpu().numpy().ravel()
    ce = np.abs(c - st["color"]).max(1)[act]
    de = np.abs(d - st["density"])[act]
    nk = ~kink_s[act]
    print(f"  per-sample (non-transparent {act.sum()}): colour err max {ce.max():.2e} p99 {np.percentile(ce, 99):.2e} median {np.median(ce):.2e}; "
          f"no-kink samples max {ce[nk].max():.2e}; sigma err max {de.max():.2e} (|sigma| max {np.abs(st['density'][act]).max():.1f}); |colour| max {np.abs(st['color']).max():.2f}")
    lightf = np.where(np.abs(st["essence"]).max(1) > 1e-3, np.abs(st["color"]).max(1) / np.maximum(np.abs(st["essence"]).max(1), 1e-9), 0)
    print("  light factor range", lightf[act].min(), lightf[act].max(), " essence max", np.abs(st["essence"][act]).max())
    worst = np.argsort(-ce)[:5]
    ai = np.nonzero(act)[0][worst]
    for w, i in zip(worst, ai):
        print("   worst sample", i, "err", ce[w], "kink", bool(kink_s[i]), "margin", st["kink_margin"][i], "colour", st["color"][i], "got", c[i], "essence", st["essence"][i])
