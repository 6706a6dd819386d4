#!/bin/bash
# usage: ab_libs.sh <lib suffix> ...   ("main" = libdsnerf.so): bench.py frame / MLP-kernel time of each library variant on one box, twice
export DSNERF_NO_CLOCK_SAMPLER=1
for rep in 1 2; do
for v in "$@"; do
  if [ "$v" = "main" ]; then unset DSNERF_LIB; else export DSNERF_LIB=$PWD/dual_space_nerf_b200/libdsnerf_$v.so; fi
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$v', 'frame', round(d['ms_per_step'],3), 'mlp', round(d['roofline']['kernel_ms_per_launch'],3), 'frac', round(d['roofline']['frac'],4), d['e2e']['rgb_checksum'], d['clocks']['sm_mhz'], d['clocks'].get('power_w'))"
done
done
