#!/usr/bin/env python
"""Benchmark of the volume-rendering hot path (BASELINE.json metric: rays/sec at
512x512, 64 samples/ray).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one full 512x512x64 frame of the synthetic scene (SURVEY.md 8d,
config 2) through the render path: geometry-guided sampling -> nearest-triangle
warp -> SpaceNet + density-gradient normal -> lighting -> compositing.  With
N > 1 (torchrun, one rank per GPU) every rank renders its own frame (config 5:
novel-pose batch, pose seed = rank) and one NCCL all-gather reassembles the
per-ray outputs on every rank: weak scaling, value = N * rays / time.

`--impl reference` times the CPU restatement of the reference (oracle/, the
reference itself is Python and does not exist on the GPU box) on the host cores
on a bounded ray slice of the same frame.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H = W = 512
N_SAMPLES = 64
FLOP_PER_SAMPLE = 1804544.0  # 902 272 MAC forward + input-gradient per evaluated sample
METRIC = "rays/sec (512x512, 64 samples/ray)"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get("bf16_tflops_sustained", 1369.6), d.get("bf16_tflops", 1629.8), d.get("hbm_gbs", 6550.7), "measured"
    return 1400.0, 1590.0, 6650.0, "fallback"


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md).  Polls NVML (the library behind
    nvidia-smi) from a thread every 50 ms; falls back to one `nvidia-smi --query-gpu` call if pynvml is unavailable."""

    def __init__(self, index):
        self.index = index
        self.sm, self.mx, self.reasons = [], [], set()
        self._stop = threading.Event()
        self.t = None
        self.nv = None

    def start(self):
        try:
            import pynvml as nv

            nv.nvmlInit()
            self.nv = nv
            self.h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
        except Exception:
            self.nv = None

    def _poll(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8)),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40)),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20)),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)),
        }
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.05)

    def stop(self):
        if self.nv is None:
            try:
                out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=10).stdout.strip().split(",")
                return {"sm_mhz": float(out[0]), "sm_max_mhz": float(out[1]), "reasons": [], "samples": 1, "source": "nvidia-smi after the run"}
            except Exception:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"]}
        self._stop.set()
        if self.t:
            self.t.join(timeout=2)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                "reasons": sorted(self.reasons), "samples": len(self.sm), "source": "NVML polled every 50 ms during the timed region"}


def cpu_port(rays_per_step, steps, warmup, sc=None, sd=None):
    """Time the oracle (CPU restatement of the reference) on a bounded slice of the 512x512x64 frame."""
    import torch

    from dual_space_nerf_b200 import net as N
    from dual_space_nerf_b200 import scene as S
    from oracle import clib
    from oracle import oracle as O

    sc = sc or S.make_scene(H, W)
    sd = sd or N.synthetic_net(0).state_dict()
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    orc = O.Oracle(sd, sc["canonical"], sc["faces"], N_SAMPLES)
    # contiguous scanlines through the middle of the frame (hit and miss rays in the frame's proportion)
    r0 = (H // 2) * W
    times = []
    for s in range(warmup + steps):
        sel = slice(r0 + s * rays_per_step, r0 + (s + 1) * rays_per_step)
        t = time.perf_counter()
        orc.render(sc["ray_o"][sel], sc["ray_d"][sel], sc["near"][sel], sc["far"][sel], sc["posed"], sc["poses"], sc["frame"])
        if s >= warmup:
            times.append(time.perf_counter() - t)
    sec = float(np.sum(times))
    return {
        "value": rays_per_step * steps / sec, "unit": "rays/s", "cores": int(max(cores, clib.lib().dso_num_threads())), "kind": "port",
        "sample": f"{steps} x {rays_per_step}-ray scanline slices of the 512x512x64 frame (numpy/OpenBLAS MLP + OpenMP C "
                  f"brute-force nearest triangle, fp32), {sec:.1f} s",
        "seconds": sec,
    }


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    rays = 4096
    res = cpu_port(rays, max(1, args.steps), min(1, args.warmup))
    line = {
        "impl": "reference", "metric": METRIC, "value": res["value"], "unit": "rays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * res["seconds"] / max(1, args.steps), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "ZJU-Mocap 313 config shape: 512x512 rays, 64 samples/ray, GG sampling, random-init weights",
                   "step": f"{rays}-ray slice of the frame on the host CPU"},
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": res["value"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "the reference is pure Python + pytorch3d and is absent on the GPU box; this arm times oracle/ (its CPU "
                "restatement, pinned to it bit-exactly on the geometry stages) on all host cores",
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--simt", action="store_true", help="debug: fp32 SIMT MLP kernel instead of tcgen05")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist

    from dual_space_nerf_b200 import lib
    from dual_space_nerf_b200 import net as N
    from dual_space_nerf_b200 import scene as S

    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (B200); there is no CPU fallback for the product path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    warmup = max(args.warmup, 3)
    sc = S.make_scene(H, W, pose_seed=rank)  # config 5 for N > 1: one novel-pose frame per GPU
    net = N.synthetic_net(0)
    sd = net.state_dict()
    R = H * W
    ctx = lib.Context(local_rank)
    L = ctx.L
    arrs = [np.ascontiguousarray(sd[k].detach().numpy(), dtype=np.float32) for k in N.STATE_DICT_ORDER]
    ptrs = (ctypes.c_void_p * len(arrs))(*[a.ctypes.data_as(ctypes.c_void_p) for a in arrs])
    ctx.check(L.dsnerf_set_weights(ctx.h, ptrs, len(arrs)))
    faces = np.ascontiguousarray(sc["faces"], dtype=np.int32)
    ctx.check(L.dsnerf_set_mesh(ctx.h, faces.ctypes.data_as(ctypes.c_void_p), faces.shape[0],
                                sc["canonical"].ctypes.data_as(ctypes.c_void_p), sc["canonical"].shape[0]))
    flags = lib.SAMPLE_GG | (lib.MLP_FP32_SIMT if args.simt else 0)
    stream = torch.cuda.current_stream(dev)
    sp = ctypes.c_void_p(stream.cuda_stream)
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    posed = np.ascontiguousarray(sc["posed"])
    poses = np.ascontiguousarray(sc["poses"])

    def set_frame():
        ctx.check(L.dsnerf_set_frame(ctx.h, posed.ctypes.data_as(ctypes.c_void_p), poses.ctypes.data_as(ctypes.c_void_p),
                                     sc["frame"], 0, None, None, None, sp))

    # ---- device-resident arm ("value"): inputs already in HBM, outputs stay in HBM
    d_o, d_d = torch.from_numpy(sc["ray_o"]).to(dev), torch.from_numpy(sc["ray_d"]).to(dev)
    d_n, d_f = torch.from_numpy(sc["near"]).to(dev), torch.from_numpy(sc["far"]).to(dev)
    out = torch.empty(R, 6, device=dev)  # rgb(3) depth acc disp packed per ray -> one all-gather
    o_rgb, o_dep, o_acc, o_dsp = (torch.empty(R, 3, device=dev), torch.empty(R, device=dev), torch.empty(R, device=dev),
                                  torch.empty(R, device=dev))
    gathered = torch.empty(world * R, 6, device=dev) if world > 1 else None
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def step_device():
        set_frame()
        ctx.check(L.dsnerf_render(ctx.h, P(d_o), P(d_d), P(d_n), P(d_f), R, N_SAMPLES, flags, P(o_rgb), P(o_dep), P(o_acc),
                                  P(o_dsp), None, None, sp))
        if world > 1:
            out[:, :3] = o_rgb
            out[:, 3] = o_dep
            out[:, 4] = o_acc
            out[:, 5] = o_dsp
            dist.all_gather_into_tensor(gathered, out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        step_device()
    torch.cuda.synchronize()
    st = ctx.stats()
    launches_per_step = st["kernel_launches"] + 5  # + the five grid-build kernels of dsnerf_set_frame
    evaluated = st["evaluated_samples"]

    ctx.profile(1 | int(os.environ.get('DSNERF_DEBUG_PROFILE_BITS', '0')))
    ctx.profile_read(reset=True)
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        flush.fill_(1)  # L2 flush between timed iterations (inside the bracket: ~0.1 ms of 256 MB writes per step)
        step_device()
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop()
    mlp_ms, mlp_n = ctx.profile_read(reset=True)
    ctx.profile(0)
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = world * R * args.steps / (ms_total * 1e-3)

    # ---- end-to-end arm: pinned HOST buffers through the C ABI (dsnerf_render_host), H2D + D2H inside the timed region
    h_o, h_d = torch.from_numpy(sc["ray_o"]).pin_memory(), torch.from_numpy(sc["ray_d"]).pin_memory()
    h_n, h_f = torch.from_numpy(sc["near"]).pin_memory(), torch.from_numpy(sc["far"]).pin_memory()
    h_rgb, h_dep, h_acc, h_dsp = (torch.empty(R, 3).pin_memory(), torch.empty(R).pin_memory(), torch.empty(R).pin_memory(),
                                  torch.empty(R).pin_memory())

    def step_e2e():
        set_frame()
        ctx.check(L.dsnerf_render_host(ctx.h, P(h_o), P(h_d), P(h_n), P(h_f), R, N_SAMPLES, flags, P(h_rgb), P(h_dep), P(h_acc),
                                       P(h_dsp), None, None, sp))

    for _ in range(2):
        step_e2e()
    barrier()
    e0.record(stream)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        flush.fill_(1)
        step_e2e()
    e1.record(stream)
    barrier()
    wall = time.perf_counter() - t0
    e2e_ms = max(e0.elapsed_time(e1), wall * 1e3)
    t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    e2e_value = world * R * args.steps / (e2e_ms * 1e-3)
    frame_checksum = float(h_rgb.double().sum())

    if rank == 0:
        sustained, burst, hbm, src = peaks()
        achieved = (evaluated * FLOP_PER_SAMPLE) / (mlp_ms / max(mlp_n, 1) * 1e-3) / 1e12 if mlp_ms > 0 else None
        traffic = None
        tp = os.path.join(ROOT, "profiles", "mlp_dram_bytes.json")
        if os.path.exists(tp):
            with open(tp) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
        line = {
            "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.simt else "f16x3-split operands, f32 accumulate (tcgen05)", "data": "synthetic",
            "config": {
                "workload": "ZJU-Mocap 313 config shape (configs[1]): 512x512 = 262144 rays, 64 samples/ray, GG sampling, "
                            "random-init SpaceNet (head rescale of SURVEY.md 8d), synthetic SMPL-sized mesh (V=6890, F=13776)",
                "rays_per_gpu_per_step": R, "samples_per_ray": N_SAMPLES, "evaluated_samples_per_step": int(evaluated),
                "evaluated_fraction": evaluated / float(R * N_SAMPLES),
                "parallelism": f"{world} x (one frame per GPU) + NCCL all-gather of 6 floats/ray" if world > 1 else "1 GPU",
                "l2": "256 MB buffer written between timed steps (L2 flush)",
                "mlp_kernel": "fp32 SIMT (debug)" if args.simt else "tcgen05",
            },
            "roofline": {
                "bound": "tensor", "achieved": achieved, "peak": sustained, "unit": "TFLOP/s",
                "frac": (achieved / sustained) if achieved else None, "traffic": traffic,
                "kernel": "mlp_simt_kernel" if args.simt else "mlp_tc_kernel", "kernel_ms_per_launch": mlp_ms / max(mlp_n, 1),
                "kernel_share_of_step": (mlp_ms / max(mlp_n, 1)) / ms_step, "peak_source": f"bf16_tflops_sustained of {src} (MEASURED_PEAKS.json)",
                "algorithmic_flop_per_launch": evaluated * FLOP_PER_SAMPLE,
            },
            "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": R * 8 * 4 + posed.nbytes + 256 * 4,
                    "d2h_bytes_per_step": R * 6 * 4, "ms_per_step": e2e_ms / args.steps,
                    "api": "dsnerf_set_frame + dsnerf_render_host (C ABI, pinned host buffers)", "rgb_checksum": frame_checksum},
            "gpu_launches": int(launches_per_step * args.steps),
            "clocks": clocks,
        }
        if not args.no_cpu_baseline:
            res = cpu_port(8192, 3, 1, sc=S.make_scene(H, W), sd=sd)
            line["cpu_baseline"] = {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
