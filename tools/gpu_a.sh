#!/bin/bash
# round-2 GPU batch A: parity tests, headline bench, pass-count experiments, config 3
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r02_tests_a.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err
DSNERF_DEBUG_PROFILE_BITS=128 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_2pass.json 2>/dev/null
DSNERF_DEBUG_PROFILE_BITS=64 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_1pass.json 2>/dev/null
python bench.py --config 3 --steps 5 > gpurun_out/r02_bench_cfg3.json 2> gpurun_out/r02_bench_cfg3.err
tail -5 gpurun_out/r02_tests_a.log
