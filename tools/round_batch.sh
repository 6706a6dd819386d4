#!/bin/bash
# final round-2 batch (1 GPU): tests, headline bench, config 3, early-stop line, reference arm, ncu captures of the final build
TAG=${1:-r02f}
python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -20 > gpurun_out/${TAG}_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_1gpu.json 2> gpurun_out/${TAG}_bench_1gpu.err
python bench.py --steps 10 --warmup 3 --config 3 > gpurun_out/${TAG}_bench_config3.json 2>/dev/null
python bench.py --steps 10 --warmup 3 --early-stop --no-cpu-baseline > gpurun_out/${TAG}_bench_early_stop.json 2>/dev/null
DSNERF_RGB_PASSES=3 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_precise_mode.json 2>/dev/null
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>/dev/null
export DSNERF_NO_CLOCK_SAMPLER=1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 150 -c 140 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:mlp_tc -s 4 -c 1 -f -o gpurun_out/${TAG}_mlp \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none -k regex:"build_cells|sample_warp|light_tc|canon_nearest|composite|mark_samples|gg_bounds" -s 28 -c 9 -f -o gpurun_out/${TAG}_tail \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu -i gpurun_out/${TAG}_mlp.ncu-rep --page raw --csv > gpurun_out/${TAG}_mlp_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_tail.ncu-rep --page raw --csv > gpurun_out/${TAG}_tail_raw.csv 2>/dev/null
# per-phase clock stamps of the two-tile MLP kernel (measurement build: make -C dual_space_nerf_b200/csrc ../libdsnerf_timing.so)
if [ -f dual_space_nerf_b200/libdsnerf_timing.so ]; then
  DSNERF_LIB=$PWD/dual_space_nerf_b200/libdsnerf_timing.so python tests/tc2_timing.py > gpurun_out/${TAG}_tc2_phase_stamps.log 2>&1
fi
tail -4 gpurun_out/${TAG}_tests.log; cat gpurun_out/${TAG}_smoke.log | tail -2
