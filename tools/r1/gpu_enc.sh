#!/bin/bash
# encode path: parity tests that touch it, then the encode bench line: tools/gpu_enc.sh
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_codec.py -x -q -k "encode or encoder or chain or round_trip or config1 or full_size" 2>&1 | tail -3
timeout 600 python bench.py --workload encode_p_1080p --steps 10 --warmup 3 --extras 0 --cpu-budget 2 --e2e 0 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l)
        print('encode_p value %.0f frac %.3f ms/step %.3f' % (d['value'], d['roofline']['frac'], d['ms_per_step']))
"
