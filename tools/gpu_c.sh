#!/bin/bash
# round-2 GPU batch C (2 GPUs): all gpu tests incl. the NCCL / fused-gather worker, bench N=2 with fused and NCCL gather
python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -60 > gpurun_out/r02_tests_c.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/sharded_worker.py 256 > gpurun_out/r02_sharded_worker.log 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_2gpu_fused.json 2> gpurun_out/r02_bench_2gpu_fused.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 --gather nccl > gpurun_out/r02_bench_2gpu_nccl.json 2> gpurun_out/r02_bench_2gpu_nccl.err
tail -5 gpurun_out/r02_tests_c.log; tail -3 gpurun_out/r02_sharded_worker.log; tail -c 600 gpurun_out/r02_bench_2gpu_fused.err
