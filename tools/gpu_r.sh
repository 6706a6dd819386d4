#!/bin/bash
export DSNERF_NO_CLOCK_SAMPLER=1
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 90 --csv --log-file gpurun_out/r02_cfg3_launches.csv python bench.py --config 3 --steps 3 --warmup 3 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r02_cfg3_launches.csv | head -14
