#!/bin/bash
# all GPU tests + the default bench line (with extras): tools/gpu_tb.sh <tag>
TAG=${1:-tb}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $OUT/pytest_$TAG.txt
timeout 900 python bench.py 2> $OUT/bench_$TAG.err | tee $OUT/bench_$TAG.json | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l)
        print('MAIN  %-22s value %9.0f frac %.3f e2e %7.0f sparse %s cpu %6.0f' % (d['config']['workload'], d['value'], d['roofline']['frac'], d['e2e']['value'], d['e2e'].get('sparse',{}).get('value'), d['cpu_baseline']['value']))
        for k,x in d.get('extras',{}).items():
            if 'error' in x: print('EXTRA', k, x['error']); continue
            if 'roofline_frac' in x:
                print('EXTRA %-22s value %9.0f frac %.3f e2e %7.0f sparse %s cpu %6.0f' % (k, x['value'], x['roofline_frac'], x['e2e']['value'], x['e2e'].get('sparse',{}).get('value'), x['cpu_baseline']['value']))
            else:
                print('EXTRA %-22s value %9.0f cpu %6.0f' % (k, x['value'], x['cpu_baseline']['value']))
"
tail -3 $OUT/bench_$TAG.err
