#!/bin/bash
# round 2, second visit: fused decode-P with the job cache, encode-I with one transform copy, decode-I drain / dense hint
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2b; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/t_all.log 2>&1
echo "rc=$?" >> $O/t_all.log
B="timeout 300 python bench.py --steps 10 --warmup 3 --extras 0 --e2e 0 --cpu-budget 0.5"
$B --workload decode_p_1080p > $O/dp.json 2> $O/dp.err
$B --workload decode_p_4k > $O/dp4k.json 2> $O/dp4k.err
PFV_ENCODE_I_CTAS=4 $B --workload encode_i_1080p > $O/ei_4.json 2> $O/ei_4.err
PFV_ENCODE_I_CTAS=3 $B --workload encode_i_1080p > $O/ei_3.json 2> $O/ei_3.err
PFV_DECODE_I_DRAIN=0 $B --workload decode_i_1080p > $O/di_pool.json 2> $O/di_pool.err
PFV_DECODE_I_DRAIN=1 $B --workload decode_i_1080p > $O/di_drain.json 2> $O/di_drain.err
$B --workload decode_i_1080p_dense > $O/dense_hint.json 2> $O/dense_hint.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_p_fused -s 20 -c 1 -o $O/prof_dp_fused python bench.py --steps 2 --warmup 3 --extras 0 --e2e 0 --cpu-budget 0 --workload decode_p_1080p > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:encode_i_stream -s 3 -c 1 -o $O/prof_ei_stream python bench.py --steps 2 --warmup 3 --extras 0 --e2e 0 --cpu-budget 0 --workload encode_i_1080p > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:encode_p_kernel -s 8 -c 1 -o $O/prof_ep python bench.py --steps 2 --warmup 3 --extras 0 --e2e 0 --cpu-budget 0 --workload encode_p_1080p > /dev/null 2>&1
ls -la $O
