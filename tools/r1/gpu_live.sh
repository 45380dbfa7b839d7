#!/bin/bash
export PFV_DECODE_P_VARIANT=live
for cc in 3 4 5 6; do for rc in 1 2 3; do
  echo "copy $cc resid $rc"
  PFV_LIVE_COPY_CTAS=$cc PFV_LIVE_RESID_CTAS=$rc timeout 200 python bench.py --workload decode_p_1080p --extras 0 --cpu-budget 0.1 --steps 10 --e2e 0 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('  value %.0f frac %.3f ms %.4f' % (d['value'], d['roofline']['frac'], d['ms_per_step']))
"
done; done
