#!/bin/bash
# usage: multi_gpu_bench.sh <N> [tag]   : torchrun bench.py --gpus N on the box's N GPUs, JSON line into gpurun_out/<tag>_bench_<N>gpu.json
N=$1; TAG=${2:-r02}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err
python - $N $TAG <<'PY'
import json,sys
N,TAG=sys.argv[1],sys.argv[2]
d=json.loads([l for l in open(f'gpurun_out/{TAG}_bench_{N}gpu.json') if l.startswith('{')][-1]); s=d['strong']
print(N,'GPUs weak', round(d['value']/1e6,2), d['ms_per_step'], d['config']['gather'], d['config']['gather_verified'], 'strong', s['ms'], s['ms_1gpu'], s['speedup_vs_1gpu'], s['bit_identical'], 'e2e', d['e2e']['value']/1e6)
print(d.get('parity'))
PY
