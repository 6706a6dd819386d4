#!/bin/bash
for st in 10 10 20; do
  python bench.py --config 3 --steps $st --warmup 3 2>/dev/null | python -c "
import sys,json
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('steps $st', round(d['ms_per_step'],2), round(d['roofline']['kernel_ms_per_frame'],2))"
done
