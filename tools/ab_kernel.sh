#!/bin/bash
export DSNERF_NO_CLOCK_SAMPLER=1
K=${1:-light_tc}
for v in new old; do
  if [ $v = old ]; then export DSNERF_LIB=$PWD/dual_space_nerf_b200/libdsnerf_base.so; fi
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:$K -s 4 -c 6 --csv --log-file gpurun_out/r02_k_$v.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
  python - $v <<'PY'
import csv,sys
rows=[r for r in csv.reader(l for l in open(f'gpurun_out/r02_k_{sys.argv[1]}.csv') if not l.startswith('=='))]
h=rows[0]; vi=h.index('Metric Value'); ki=h.index('Kernel Name')
print(sys.argv[1], [(r[ki].split('(')[0][-24:], round(float(r[vi].replace(',',''))/1e3,1)) for r in rows[1:] if len(r)>vi])
PY
done
unset DSNERF_LIB
python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -3
for i in 1 2; do
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ab_q_new_$i.json 2>/dev/null
DSNERF_LIB=$PWD/dual_space_nerf_b200/libdsnerf_base.so python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ab_q_old_$i.json 2>/dev/null
done
for f in gpurun_out/r02_ab_q_*.json; do python - "$f" <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1]); r=d['roofline']
print(sys.argv[1], round(d['ms_per_step'],3), round(r['kernel_ms_per_launch'],3), 'tail', round(d['ms_per_step']-r['kernel_ms_per_launch'],3), d['e2e']['rgb_checksum'])
PY
done
