#!/bin/bash
# round 2, visit zt: Decoder throughput against the size of the entropy pool on the box's 16 hardware threads
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2zt; mkdir -p $O
nproc > $O/dec_sweep.txt
timeout 300 python tools/exp/dec_trace.py 8 10 12 13 14 15 16 >> $O/dec_sweep.txt 2>&1
grep "fps\|^[0-9]" $O/dec_sweep.txt
