#!/bin/bash
# 8-GPU validation: weak (config 5) + strong (config 4) with the fused gather and with NCCL, plus N = 4
N=${1:-8}
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus $1 --steps 10 --warmup 3 $3 > gpurun_out/r02_bench_$1gpu_$4.json 2> gpurun_out/r02_bench_$1gpu_$4.err; }
run $N 29521 "" fused
run $N 29522 "--gather nccl" nccl
if [ "$N" = "8" ]; then run 4 29523 "" fused; fi
for f in gpurun_out/r02_bench_*gpu_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1]); s=d.get('strong',{})
    print(sys.argv[1], 'weak', round(d['value']/1e6,2), round(d['ms_per_step'],3), d['config']['gather'][:20], d['config']['gather_verified'], '| strong', round(s.get('rays_s',0)/1e6,2), round(s.get('ms',0),3), round(s.get('speedup_vs_1gpu',0),3), s.get('bit_identical'), '| parity', d['parity']['rays_over_tol_unexplained'], d['parity']['max_abs_depth'])
except Exception as e: print(sys.argv[1],'ERR',e)
PY
done
