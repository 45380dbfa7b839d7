#!/bin/bash
# tok_store_kernel grid size against the sparse encode e2e leg and the Encoder object: tools/gpu_store_sweep.sh v1 v2 ...
for v in "$@"; do
  echo "== PFV_TOK_STORE_CTAS=$v"
  PFV_TOK_STORE_CTAS=$v timeout 300 python bench.py --workload encode_p_1080p --steps 6 --warmup 3 --extras 0 --cpu-budget 0.1 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('  e2e dense %.0f sparse %.0f' % (d['e2e']['value'], d['e2e']['sparse']['value']))
"
done
