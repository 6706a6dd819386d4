"""torchrun worker (not a pytest file): BASELINE configs[3] style -- one frame, rays sharded across the ranks over NCCL
(dual_space_nerf_b200.dist.render_sharded) -- against the unsharded render of the same frame on every rank.
Usage: torchrun --nproc-per-node N tests/sharded_worker.py [H]"""
import os
import sys
import time
from types import SimpleNamespace

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dual_space_nerf_b200 import dist as D  # noqa: E402
from dual_space_nerf_b200 import net as N  # noqa: E402
from dual_space_nerf_b200 import scene as S  # noqa: E402
from dual_space_nerf_b200.renderer import Renderer  # noqa: E402


def main():
    H = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sc = S.make_scene(H, H)
    cfg = SimpleNamespace(MODEL=SimpleNamespace(TYPE="nerf", COARSE_RAY_SAMPLING=64, FINE_RAY_SAMPLING=-1, sample_points_mode="GG",
                                                perturb=1.0, raw_noise_std=1.0), DATASETS=SimpleNamespace(SMPL_PATH=None))
    r = Renderer(N.synthetic_net(0), None, cfg, torch.from_numpy(sc["canonical"]), device=local, faces=sc["faces"])
    r.eval()
    b = S.to_batch(sc, torch, device=f"cuda:{local}")
    full = r.render(b)["coarse"]
    times = []
    for _ in range(4):
        dist.barrier()
        torch.cuda.synchronize()
        t = time.perf_counter()
        sh = D.render_sharded(r, b)
        torch.cuda.synchronize()
        dist.barrier()
        times.append(time.perf_counter() - t)
    same = lambda a: all(torch.equal(a[k].nan_to_num(-1.0), full[k].nan_to_num(-1.0)) for k in ("color", "depth_map", "acc_map", "disp_map"))
    ok = same(sh)
    # rows dealt round-robin over NCCL, then the same with the reassembly fused into the compositor's peer stores
    ok = ok and same(D.render_sharded(r, b, interleave=H))
    fused = "unavailable"
    try:
        fx = D.FrameExchange(H * H // world, torch.device("cuda", local), n_slots=2)
    except Exception as e:  # no symmetric memory: reported, not fatal
        fx = None
        fused = f"unavailable ({type(e).__name__})"
    if fx is not None:
        ok_f = True
        for slot in (0, 1, 0):
            ok_f = ok_f and same(D.render_sharded(r, b, interleave=H, exchange=fx, slot=slot))
        fused = str(ok_f)
        ok = ok and ok_f
        # the same through the NVSwitch multicast address (multimem.st), where the fabric offers one
        fxm = D.FrameExchange(H * H // world, torch.device("cuda", local), n_slots=1, multicast=True)
        if fxm.mc:
            ok_m = same(D.render_sharded(r, b, interleave=H, exchange=fxm, slot=0))
            fused += f" multicast={ok_m}"
            ok = ok and ok_m
        else:
            fused += " multicast=unavailable"
    flag = torch.tensor([1 if ok else 0], device=f"cuda:{local}")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"SHARDED world={world} {H}x{H}x64 bit_identical={bool(flag.item())} fused_gather={fused} ms={min(times) * 1e3:.2f} "
              f"rays_per_s={H * H / min(times):.4g}", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
