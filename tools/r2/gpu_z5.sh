#!/bin/bash
# round 2, visit z5: decode-P 1080p with 64 GOPs per launch
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2z5; mkdir -p $O
B="timeout 400 python bench.py --steps 10 --warmup 3 --extras 0 --e2e 0 --cpu-budget 0"
$B --workload decode_p_1080p_64 > $O/dp64.json 2> $O/dp64.err
$B --workload decode_p_1080p > $O/dp32.json 2> $O/dp32.err
tail -n 2 $O/dp64.err
