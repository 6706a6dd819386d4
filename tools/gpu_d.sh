#!/bin/bash
python tests/bigw_debug.py > gpurun_out/r02_bigw_debug.log 2>&1
DSNERF_RGB_PASSES=1 python tests/bigw_debug.py > gpurun_out/r02_bigw_debug_rgb1.log 2>&1
python -m pytest tests -m gpu -q --tb=short -x -k "render_host or restaged or config1" 2>&1 | tail -15 > gpurun_out/r02_tests_d.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_d.json 2> gpurun_out/r02_bench_d.err
DSNERF_NO_SIDE_STREAM=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_d_noside.json 2>/dev/null
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_d2.json 2>/dev/null
cat gpurun_out/r02_bigw_debug.log | tail -30; tail -3 gpurun_out/r02_tests_d.log
