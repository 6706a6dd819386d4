#!/bin/bash
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 tools/strong_sweep.py 1 8 16 32 2>&1 | grep -v "^\*\|OMP_NUM" | tee gpurun_out/r02_strong_sweep_8gpu.log
