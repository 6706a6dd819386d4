#!/bin/bash
# round-2 profiling batch (1 GPU): launch list with DRAM bytes, ncu --set full of the MLP kernel and of the tail kernels
export DSNERF_NO_CLOCK_SAMPLER=1
TAG=${1:-r02}
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 150 -c 140 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_prof_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mlp_tc -s 4 -c 1 -f -o gpurun_out/${TAG}_mlp \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_prof_mlp.log 2>&1
ncu --set full --clock-control none -k regex:"build_cells|sample_warp|light_tc|canon_nearest|composite|mark_samples|gg_bounds" -s 28 -c 9 -f -o gpurun_out/${TAG}_tail \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_prof_tail.log 2>&1
ncu -i gpurun_out/${TAG}_mlp.ncu-rep --page raw --csv > gpurun_out/${TAG}_mlp_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_tail.ncu-rep --page raw --csv > gpurun_out/${TAG}_tail_raw.csv 2>/dev/null
ls -la gpurun_out/${TAG}_*
python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -15 > gpurun_out/${TAG}_tests_e.log
python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_e.json 2> gpurun_out/${TAG}_bench_e.err
tail -4 gpurun_out/${TAG}_tests_e.log
