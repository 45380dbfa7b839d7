#!/bin/bash
# round 2, visit zm: where the Encoder's time goes (PFV_TRACE) at 16 and 8 entropy threads
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2zm; mkdir -p $O
nproc > $O/trace.txt
timeout 300 python tools/exp/enc_trace.py >> $O/trace.txt 2>&1
timeout 300 python tools/exp/enc_trace.py 8 >> $O/trace.txt 2>&1
timeout 300 python tools/exp/enc_trace.py 12 >> $O/trace.txt 2>&1
cat $O/trace.txt
