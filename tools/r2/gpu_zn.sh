#!/bin/bash
# round 2, visit zn: device time of ONE 1080p frame through the sparse encode seam (what the Encoder object submits)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2zn; mkdir -p $O
timeout 300 python tools/exp/tok_cost.py 1 > $O/tok_cost_1.txt 2>&1
timeout 300 python tools/exp/tok_cost.py 2 > $O/tok_cost_2.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_1.csv python tools/exp/tok_cost.py 1 > /dev/null 2>&1
cat $O/tok_cost_1.txt $O/tok_cost_2.txt
