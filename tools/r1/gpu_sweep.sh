#!/bin/bash
# sweep an env knob over one bench workload: tools/gpu_sweep.sh <workload> <ENVNAME> v1 v2 ...
WL=$1; ENVN=$2; shift 2
mkdir -p gpurun_out
for v in "$@"; do
  echo "== $ENVN=$v"
  env $ENVN=$v timeout 300 python bench.py --workload $WL --extras 0 --cpu-budget 0.1 --steps 20 --e2e 0 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('  value %.0f frac %.3f ms %.4f' % (d['value'], d['roofline']['frac'], d['ms_per_step']))
"
done
