#!/bin/bash
# round 2, visit zc: encode-P with 13 warps per CTA (window pitch 160 leaves room for one more) against 12
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2zc; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x -k "encode_p or encode_pframe or sparse_encode or encoder_chain" > $O/t.log 2>&1
echo "rc=$?" >> $O/t.log
B="timeout 300 python bench.py --steps 10 --warmup 3 --extras 0 --e2e 0 --cpu-budget 0.5"
$B --workload encode_p_1080p > $O/ep_13.json 2> $O/ep_13.err
PFV_EP2_WARPS=12 $B --workload encode_p_1080p > $O/ep_12.json 2> $O/ep_12.err
$B --workload encode_p_1080p > $O/ep_13b.json 2> $O/ep_13b.err
tail -n 3 $O/t.log
