TAG=r02q
python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -20 > gpurun_out/${TAG}_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>/dev/null
cp profiles/r02_parity.json gpurun_out/${TAG}_parity.json
tail -4 gpurun_out/${TAG}_tests.log; tail -2 gpurun_out/${TAG}_smoke.log; tail -c 400 gpurun_out/${TAG}_bench_reference.json
