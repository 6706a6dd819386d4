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
    """SM clock and throttle reasons DURING the timed region, the way B200_PROFILING.md prescribes: an `nvidia-smi
    --query-gpu ... -lms` child process started before the region and killed after it.  (An in-process NVML polling
    thread was measured to stretch the timed region by 2-30 ms per step: its queries serialise with kernel launches.)"""

    FIELDS = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index, period_ms=100):
        self.index, self.period_ms, self.proc, self.t0, self.t1 = index, period_ms, None, None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", str(self.period_ms)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def mark(self, which):  # wall-clock bounds of the timed region
        if which == 0:
            self.t0 = time.time()
        else:
            self.t1 = time.time()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"]}
        try:
            self.proc.terminate()
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            out = ""
        import datetime

        rows = []
        for ln in out.strip().splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(f[1]), float(f[2]), float(f[3]), f[4:8]))
            except Exception:
                continue
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"]}
        inside = [r for r in rows if self.t0 is not None and self.t0 - 0.02 <= r[0] <= self.t1 + 0.02]
        # a short region may fall between two samples: then take the samples that bracket it
        use = inside if inside else sorted(rows, key=lambda r: min(abs(r[0] - self.t0), abs(r[0] - self.t1)))[:2]
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        reasons = sorted({n for r in use for n, v in zip(names, r[4]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median([r[1] for r in use])), "sm_max_mhz": max(r[2] for r in use), "power_w": max(r[3] for r in use),
                "reasons": reasons, "samples": len(use), "samples_inside_region": len(inside),
                "source": f"nvidia-smi --query-gpu -lms {self.period_ms} child process running across the timed region"}


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
    ap.add_argument("--early-stop", action="store_true", help="optional DSNERF_EARLY_STOP mode (not the headline: the default evaluates every sample)")
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
    flags = lib.SAMPLE_GG | (lib.MLP_FP32_SIMT if args.simt else 0) | (lib.EARLY_STOP if args.early_stop else 0)
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
    # one flat block per rank, [rgb (R,3) | depth (R) | acc (R) | disp (R)]: the compositor writes straight into it and a single
    # all-gather (6 floats per ray) reassembles every rank's frame on every rank
    out = torch.empty(6 * R, device=dev)
    o_rgb, o_dep, o_acc, o_dsp = out[: 3 * R].view(R, 3), out[3 * R: 4 * R], out[4 * R: 5 * R], out[5 * R:]
    gathered = torch.empty(world, 6 * R, device=dev) if world > 1 else None
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def step_device():
        set_frame()
        ctx.check(L.dsnerf_render(ctx.h, P(d_o), P(d_d), P(d_n), P(d_f), R, N_SAMPLES, flags, P(o_rgb), P(o_dep), P(o_acc),
                                  P(o_dsp), None, None, sp))
        if world > 1:
            dist.all_gather_into_tensor(gathered, out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        step_device()
    torch.cuda.synchronize()
    st = ctx.stats()
    launches_per_step = st["kernel_launches"] + 1  # render + per-frame grid build (counted by the library) + the clock probe
    evaluated = st["evaluated_samples"]

    ctx.profile(1 | int(os.environ.get('DSNERF_DEBUG_PROFILE_BITS', '0')))
    ctx.profile_read(reset=True)
    # SM clock inside the timed region: a 30 us one-warp probe kernel per step on the launch stream (clock64 vs global timer).
    # NVML / nvidia-smi queries are NOT issued inside the timed region: each one was measured to stall kernel launches for
    # 30-100 ms on this driver (ms_per_step 13.8 -> 23..88 ms); throttle reasons and power are sampled during an identical,
    # untimed repeat of the region right after it.
    d_clk = torch.zeros(args.steps, 2, device=dev)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(args.steps):
        flush.fill_(1)  # L2 flush between timed iterations (inside the bracket: ~0.1 ms of 256 MB writes per step)
        step_device()
        ctx.check(L.dsnerf_debug_sm_clock(ctx.h, ctypes.c_void_p(d_clk.data_ptr() + 8 * i), sp))
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    mlp_ms, mlp_n = ctx.profile_read(reset=True)
    ctx.profile(0)
    probe = d_clk.cpu().numpy()
    sampler = ClockSampler(local_rank)
    if not os.environ.get('DSNERF_NO_CLOCK_SAMPLER'):
        sampler.start()
        time.sleep(0.4)
        sampler.mark(0)
        t_rep = time.time()
        while time.time() - t_rep < 1.0:  # identical load, untimed, long enough for several samples
            flush.fill_(1)
            step_device()
            torch.cuda.synchronize()
        sampler.mark(1)
    clocks = sampler.stop()
    clocks["sm_mhz_nvidia_smi_repeat"] = clocks.get("sm_mhz")
    clocks["sm_mhz"] = float(np.median(probe[:, 0]))
    clocks["sm_mhz_min"] = float(probe[:, 0].min())
    clocks["source"] = ("sm_mhz: clock64/globaltimer probe kernel after every timed step (inside the timed region); reasons, power, "
                        "sm_max_mhz: " + str(clocks.get("source")) + " during an untimed repeat of the same loop")
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
                "early_stop": bool(args.early_stop),
            },
            "roofline": {
                "bound": "tensor", "achieved": achieved, "peak": sustained, "unit": "TFLOP/s",
                "frac": (achieved / sustained) if achieved else None, "traffic": traffic,
                "kernel": "mlp_simt_kernel" if args.simt else "mlp_tc_kernel", "kernel_ms_per_launch": mlp_ms / max(mlp_n, 1),
                "kernel_share_of_step": (mlp_ms / max(mlp_n, 1)) / ms_step, "peak_source": f"bf16_tflops_sustained of {src} (MEASURED_PEAKS.json)",
                "algorithmic_flop_per_launch": evaluated * FLOP_PER_SAMPLE,
                # the 1e-4 parity bound forces three fp16 MMAs per forward k-step: the tensor pipe executes
                # 1 734 656 MAC per sample for 902 272 algorithmic ones; reported alongside, never instead
                "executed_tensor_tflops": None if (achieved is None or args.simt) else achieved * (1734656.0 / 902272.0),
                "executed_frac_of_peak": None if (achieved is None or args.simt) else achieved * (1734656.0 / 902272.0) / sustained,
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
