#!/bin/bash
# round 2, fifth visit: fused decode-P without the per-window pipeline barrier; encode-I with coalesced coefficient stores
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2e; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/t_all.log 2>&1
echo "rc=$?" >> $O/t_all.log
B="timeout 300 python bench.py --steps 10 --warmup 3 --extras 0 --e2e 0 --cpu-budget 0.5"
$B --workload decode_p_1080p > $O/dp.json 2> $O/dp.err
$B --workload decode_p_4k > $O/dp4k.json 2> $O/dp4k.err
$B --workload encode_i_1080p > $O/ei.json 2> $O/ei.err
$B --workload encode_p_1080p > $O/ep.json 2> $O/ep.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_p_fused -s 20 -c 1 -o $O/prof_dp_fused python bench.py --steps 2 --warmup 3 --extras 0 --e2e 0 --cpu-budget 0 --workload decode_p_1080p > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:encode_i_stream -s 3 -c 1 -o $O/prof_ei_stream python bench.py --steps 2 --warmup 3 --extras 0 --e2e 0 --cpu-budget 0 --workload encode_i_1080p > /dev/null 2>&1
ls -la $O
