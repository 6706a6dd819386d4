#!/usr/bin/env python
"""torchrun worker: config-4 strong scaling (bench.bench_strong) for several shard block heights.
Usage: torchrun --nproc-per-node N tools/strong_sweep.py [rows ...]"""
import json
import os
import sys
from types import SimpleNamespace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
args = SimpleNamespace(simt=False, early_stop=False)
rig = B.Rig(args, rank, local, world)
rig.fx = object() if os.environ.get("DSNERF_SWEEP_GATHER", "fused") == "fused" else None
for rows in [int(a) for a in sys.argv[1:]] or [1, 16, 32]:
    os.environ["DSNERF_STRONG_ROWS"] = str(rows)
    res = B.bench_strong(rig, 10, 3)
    if rank == 0:
        print(json.dumps({k: res[k] for k in ("rows_per_block", "ms", "ms_1gpu", "speedup_vs_1gpu", "bit_identical", "evaluated_samples_per_rank")}), flush=True)
dist.barrier()
dist.destroy_process_group()
