#!/bin/bash
export DSNERF_NO_CLOCK_SAMPLER=1
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -s 150 -c 70 --csv \
    --log-file gpurun_out/r02g_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(l for l in open('gpurun_out/r02g_launches.csv') if not l.startswith('=='))]
h=rows[0]; ki,mi,vi,ii=h.index('Kernel Name'),h.index('Metric Name'),h.index('Metric Value'),h.index('ID')
agg=collections.OrderedDict()
for r in rows[1:]:
    if len(r)<=vi: continue
    n=r[ki].split('(')[0].replace('void ','').replace('dsn::','')
    a=agg.setdefault(n,collections.defaultdict(list)); a[r[mi]].append(float(r[vi].replace(',','')))
for n,a in sorted(agg.items(), key=lambda kv:-sum(kv[1]['gpu__time_duration.sum'])):
    t=a['gpu__time_duration.sum']
    print('%-34s n=%2d mean %8.1f us max %8.1f inst(max) %11.0f thr/inst %5.1f issue %5.1f warps %5.1f'%(n[:34],len(t),sum(t)/len(t)/1e3,max(t)/1e3,max(a['smsp__inst_executed.sum']),max(a['smsp__thread_inst_executed_per_inst_executed.ratio']),max(a['smsp__issue_active.avg.pct_of_peak_sustained_active']),max(a['sm__warps_active.avg.pct_of_peak_sustained_active'])))
PY
