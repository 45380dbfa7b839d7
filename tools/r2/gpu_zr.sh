#!/bin/bash
# round 2, visit zr: where a long single-stream decode spends its time (PFV_TRACE)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2zr; mkdir -p $O
PFV_TRACE=1 timeout 300 python tools/exp/dec_trace.py > $O/dec_trace.txt 2>&1
cat $O/dec_trace.txt
