#!/bin/bash
# round 2, twelfth visit: the small-submit path A/B on one box, encode-P (batched predictor copy, prefetched source rows) and
# its sensitivity to resident warps
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2l; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x -k "encode or smoke or encoder" > $O/t_enc.log 2>&1
echo "rc=$?" >> $O/t_enc.log
for lean in 1 0 1 0; do
  PFV_LEAN_SUBMIT=$lean timeout 300 python tools/exp/small_submit.py >> $O/small_submit.txt 2>> $O/small_submit.err
done
B="timeout 300 python bench.py --steps 10 --warmup 3 --extras 0 --e2e 0 --cpu-budget 0.5"
for w in 12 10 8; do
  PFV_EP2_WARPS=$w $B --workload encode_p_1080p > $O/ep_w$w.json 2> $O/ep_w$w.err
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:encode_p2 -s 6 -c 1 -o $O/prof_ep2 python bench.py --steps 2 --warmup 3 --extras 0 --e2e 0 --cpu-budget 0 --workload encode_p_1080p > /dev/null 2>&1
ls -la $O
