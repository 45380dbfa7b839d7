#!/bin/bash
# ncu launch list (device time per launch) of one bench workload: tools/gpu_launchlist.sh <tag> <workload>
TAG=$1; WL=$2
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --workload $WL --steps 2 --warmup 3 --extras 0 --cpu-budget 0.1 --e2e 0 > gpurun_out/ncu_launch_$TAG.log 2>&1
python - <<PY
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/launches_$TAG.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
d=collections.defaultdict(list)
for r in rows[1:]:
    try: d[r[ki][:60]].append(float(r[vi].replace(',','')))
    except: pass
for k,v in d.items(): print(f'{k:60s} n={len(v):4d} mean={sum(v)/len(v):10.1f} min={min(v):10.1f} max={max(v):10.1f}')
PY
