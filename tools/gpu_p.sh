#!/bin/bash
for i in 1 2; do
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ab_sw_new_$i.json 2>/dev/null
  DSNERF_LIB=$PWD/dual_space_nerf_b200/libdsnerf_base.so python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ab_sw_old_$i.json 2>/dev/null
done
python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -8 > gpurun_out/r02_tests_p.log
tail -3 gpurun_out/r02_tests_p.log
for f in gpurun_out/r02_ab_sw_*.json; do python - "$f" <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1]); r=d['roofline']
print(sys.argv[1], round(d['ms_per_step'],3), round(r['kernel_ms_per_launch'],3), round(d['e2e']['ms_per_step'],3), d['e2e']['rgb_checksum'])
PY
done
