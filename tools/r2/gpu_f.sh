#!/bin/bash
# round 2, sixth visit: Encoder with submitter / writer threads; fused decode-P as 4 pipelines of 3-row windows; whole bench
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2f; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/t_all.log 2>&1
echo "rc=$?" >> $O/t_all.log
B="timeout 300 python bench.py --steps 10 --warmup 3 --extras 0 --e2e 0 --cpu-budget 0.5"
$B --workload decode_p_1080p > $O/dp.json 2> $O/dp.err
$B --workload decode_p_4k > $O/dp4k.json 2> $O/dp4k.err
PFV_TRACE=1 timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err
echo "rc=$?" >> $O/bench.err
ls -la $O
