#!/bin/bash
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_bench_8gpu_multicast.json 2> gpurun_out/r02_bench_8gpu_multicast.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_8gpu_multicast.json') if l.startswith('{')][-1]); s=d['strong']
print('8 GPUs weak', round(d['value']/1e6,2), d['ms_per_step'], d['config']['gather'], d['config']['gather_verified'], 'strong', s['ms'], s['speedup_vs_1gpu'], s['bit_identical'], s['rows_per_block'], 'e2e', d['e2e']['value']/1e6)
print(d['parity'])
PY
