#!/bin/bash
# last batch of the round (1 GPU): launch list with DRAM bytes, ncu --set full of the search kernels, bench lines of the final build
TAG=${1:-r02z}
python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_1gpu.json 2> gpurun_out/${TAG}_bench_1gpu.err
python bench.py --steps 10 --warmup 3 --config 3 > gpurun_out/${TAG}_bench_config3.json 2>/dev/null
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1
export DSNERF_NO_CLOCK_SAMPLER=1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 150 -c 144 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-early-stop-line > /dev/null 2>&1
ncu --set full --clock-control none -k regex:"build_cells|sample_warp|canon_" -s 14 -c 7 -f -o gpurun_out/${TAG}_tail \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-early-stop-line > /dev/null 2>&1
ncu -i gpurun_out/${TAG}_tail.ncu-rep --page raw --csv > gpurun_out/${TAG}_tail_raw.csv 2>/dev/null
rm -f gpurun_out/${TAG}_tail.ncu-rep
tail -2 gpurun_out/${TAG}_smoke.log; tail -c 300 gpurun_out/${TAG}_bench_config3.json
