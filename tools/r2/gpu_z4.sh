#!/bin/bash
# round 2, visit z4: encode-P search without the discarded middle candidate
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2z4; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x -k "encode_p or encode_pframe or encoder or smoke or fullsize or full_gop or chain" > $O/t_ep.log 2>&1
echo "rc=$?" >> $O/t_ep.log
B="timeout 300 python bench.py --steps 10 --warmup 3 --extras 0 --e2e 0 --cpu-budget 0.5"
$B --workload encode_p_1080p > $O/ep.json 2> $O/ep.err
$B --workload encode_p_1080p > $O/ep2.json 2> $O/ep2.err
tail -n 3 $O/t_ep.log
