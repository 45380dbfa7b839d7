#!/bin/bash
# round 2, eleventh visit: all GPU tests on the new kernels; decode-P with the rows fetched before the stores; encode-I with the
# unconditional prefetch; the bench line
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2k; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/t_all.log 2>&1
echo "rc=$?" >> $O/t_all.log
B="timeout 300 python bench.py --steps 10 --warmup 3 --extras 0 --e2e 0 --cpu-budget 0.5"
$B --workload encode_i_1080p > $O/ei.json 2> $O/ei.err
$B --workload decode_p_1080p > $O/dp.json 2> $O/dp.err
$B --workload decode_p_4k > $O/dp4k.json 2> $O/dp4k.err
$B --workload decode_i_1080p > $O/di.json 2> $O/di.err
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_p_fused -s 20 -c 1 -o $O/prof_dp_fused python bench.py --steps 2 --warmup 3 --extras 0 --e2e 0 --cpu-budget 0 --workload decode_p_1080p > /dev/null 2>&1
ls -la $O
