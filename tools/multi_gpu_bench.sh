#!/bin/bash
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_bench_${N}gpu_final.json 2> gpurun_out/r02_bench_${N}gpu_final.err
python - $N <<'PY'
import json,sys
N=sys.argv[1]
d=json.loads([l for l in open(f'gpurun_out/r02_bench_{N}gpu_final.json') if l.startswith('{')][-1]); s=d['strong']
print(N,'GPUs weak', round(d['value']/1e6,2), d['ms_per_step'], d['config']['gather'], d['config']['gather_verified'], 'strong', s['ms'], s['ms_1gpu'], s['speedup_vs_1gpu'], s['bit_identical'], 'e2e', d['e2e']['value']/1e6)
PY
if [ "$N" = "2" ]; then python bench.py --config 3 --steps 10 --warmup 3 > gpurun_out/r02_bench_config3_final.json 2>/dev/null; python -c "
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_config3_final.json') if l.startswith('{')][-1]); print('config3', d['value']/1e6, d['ms_per_step'], d['roofline']['frac'])"; fi
