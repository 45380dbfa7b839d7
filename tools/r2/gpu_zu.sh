#!/bin/bash
# round 2, visit zu: Decoder with a polling wait on the calling thread; codec tests
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2zu; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_codec.py -m gpu -q -x > $O/t.log 2>&1
echo "rc=$?" >> $O/t.log
timeout 300 python tools/exp/dec_trace.py 12 16 > $O/dec_sweep.txt 2>&1
PFV_TRACE=1 timeout 300 python tools/exp/dec_trace.py 16 > $O/dec_trace.txt 2>&1
tail -n 2 $O/t.log; grep "fps" $O/dec_sweep.txt; grep "pfv_decoder\|decode submits" $O/dec_trace.txt | cut -c1-330
