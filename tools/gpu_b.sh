#!/bin/bash
# round-2 GPU batch B: full test log, in-kernel clock for the pass-count variants, bench with rgb 1-pass vs 3-pass
python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -150 > gpurun_out/r02_tests_b.log
for bits in 0 128 64; do
  echo "== profile bits $bits" >> gpurun_out/r02_tc_timing.log
  DSNERF_TIMING_HW=512 DSNERF_DEBUG_PROFILE_BITS=$bits python tests/tc_timing.py >> gpurun_out/r02_tc_timing.log 2>&1
done
DSNERF_RGB_PASSES=3 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_rgb3.json 2>/dev/null
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_b.json 2> gpurun_out/r02_bench_b.err
tail -12 gpurun_out/r02_tests_b.log
