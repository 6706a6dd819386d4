#!/bin/bash
for w in 2 3 4; do
  DSNERF_ERT_WAVES=$w python bench.py --steps 10 --warmup 3 --no-cpu-baseline --early-stop > gpurun_out/r02_ert_w$w.json 2>/dev/null
done
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ert_off.json 2>/dev/null
python -m pytest tests -m gpu -q --tb=short -k early_stop 2>&1 | tail -5 > gpurun_out/r02_tests_g.log
tail -3 gpurun_out/r02_tests_g.log
for f in gpurun_out/r02_ert_*.json; do python - "$f" <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1]); r=d['roofline']
print(sys.argv[1], round(d['ms_per_step'],3), 'mlp/launch', round(r['kernel_ms_per_launch'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'evaluated', d['config']['evaluated_samples_per_step'], 'launches', d['gpu_launches'])
PY
done
