#!/bin/bash
# round 2, visit zs: staging reserved at Decoder / Encoder open; codec tests; traces
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2zs; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_codec.py -m gpu -q -x > $O/t.log 2>&1
echo "rc=$?" >> $O/t.log
PFV_TRACE=1 timeout 300 python tools/exp/dec_trace.py > $O/dec_trace.txt 2>&1
timeout 300 python tools/exp/enc_trace.py > $O/enc_trace.txt 2>&1
tail -n 3 $O/t.log; grep "fps\|decode submits" $O/dec_trace.txt | cut -c1-300; grep "frames/s" $O/enc_trace.txt
