#!/bin/bash
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_1gpu_last.json 2> gpurun_out/r02_bench_1gpu_last.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_1gpu_last.json') if l.startswith('{')][-1]); r=d['roofline']
print('value', round(d['value']/1e6,3), d['ms_per_step'], 'mlp', r['kernel_ms_per_launch'], r['frac'], r['whole_step_frac'], 'e2e', d['e2e']['value']/1e6, d['e2e']['ms_per_step'], d['clocks']['sm_mhz'], d['clocks']['sm_mhz_min'], d['clocks']['reasons'], d['gpu_launches'])
print(d['parity']['rays_over_tol_unexplained'], d['parity']['max_abs_rgb_within_tol_rays'], d['parity']['max_abs_depth'])
PY
python -m pytest tests -m gpu -q 2>&1 | tail -2
