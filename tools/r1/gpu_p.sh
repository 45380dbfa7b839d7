#!/bin/bash
# decode-P iteration: parity tests that touch decode-P, then the P workload with the v1 and v2 kernels, launch list
TAG=${1:-p}
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_codec.py -x -q -k "pframe or long_motion or live or variants or stream or batched or decoder or sparse or full_size" 2>&1 | tail -5
for V in ${VARIANTS:-win winll two1}; do
  echo "== PFV_DECODE_P_VARIANT=$V"
  PFV_DECODE_P_VARIANT=$V timeout 300 python bench.py --workload decode_p_1080p --extras 0 --cpu-budget 0.1 --steps 10 --e2e 0 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('  value %.0f frac %.3f ms %.4f launches %d' % (d['value'], d['roofline']['frac'], d['ms_per_step'], d['roofline']['launches_per_step']))
"
done
bash tools/gpu_launchlist.sh p_$TAG decode_p_1080p
