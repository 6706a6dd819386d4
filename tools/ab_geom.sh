#!/bin/bash
# same-box A/B of the search kernels: ncu kernel times (new library, then libdsnerf_base.so), then frame times twice each
export DSNERF_NO_CLOCK_SAMPLER=1
for v in new old; do
  if [ $v = old ]; then export DSNERF_LIB=$PWD/dual_space_nerf_b200/libdsnerf_base.so; fi
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"canon_|sample_warp|build_cells" -s 14 -c 14 --csv --log-file gpurun_out/r02p_geom_$v.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-early-stop-line > /dev/null 2>&1
  python - $v <<'PY'
import csv,sys,collections
rows=[r for r in csv.reader(l for l in open(f'gpurun_out/r02p_geom_{sys.argv[1]}.csv') if not l.startswith('=='))]
h=rows[0]; vi=h.index('Metric Value'); ki=h.index('Kernel Name')
d=collections.defaultdict(list)
for r in rows[1:]:
    if len(r)>vi: d[r[ki].split('(')[0][-22:]].append(round(float(r[vi].replace(',',''))/1e3,1))
print(sys.argv[1], dict(d))
PY
done
unset DSNERF_LIB
tools/ab_libs.sh base main
