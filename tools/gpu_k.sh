#!/bin/bash
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tests/sharded_worker.py 256 2>&1 | grep -v "^\*\|OMP_NUM" | tee gpurun_out/r02_sharded_worker2.log
DSNERF_GATHER_MULTICAST=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_2gpu_multicast.json 2> gpurun_out/r02_bench_2gpu_multicast.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_2gpu_multicast.json') if l.startswith('{')][-1]); s=d['strong']
print('multicast weak', round(d['value']/1e6,2), d['ms_per_step'], d['config']['gather'], d['config']['gather_verified'], 'strong', s['ms'], s['speedup_vs_1gpu'], s['bit_identical'])
PY
python -m pytest tests -m gpu -q --tb=short -k nccl 2>&1 | tail -3
