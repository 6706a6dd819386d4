#!/bin/bash
# usage: ab_variants.sh <kernel regex> <lib suffix> ...   ("" = libdsnerf.so): ncu kernel time of each variant on one box
export DSNERF_NO_CLOCK_SAMPLER=1
K=$1; shift
for v in "$@"; do
  if [ "$v" = "main" ]; then unset DSNERF_LIB; else export DSNERF_LIB=$PWD/dual_space_nerf_b200/libdsnerf_$v.so; fi
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:$K -s 4 -c 6 --csv --log-file gpurun_out/ab_$v.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
  python - $v <<'PY'
import csv,sys
rows=[r for r in csv.reader(l for l in open(f'gpurun_out/ab_{sys.argv[1]}.csv') if not l.startswith('=='))]
h=rows[0]; vi=h.index('Metric Value')
print(sys.argv[1], [round(float(r[vi].replace(',',''))/1e3,1) for r in rows[1:] if len(r)>vi])
PY
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('   bench', round(d['ms_per_step'],3), round(d['roofline']['kernel_ms_per_launch'],3), d['e2e']['rgb_checksum'])"
done
