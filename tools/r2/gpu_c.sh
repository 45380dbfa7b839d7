#!/bin/bash
# round 2, third visit: encode-P column-strip kernel, fused decode-P with the item table, encode-I with exact forward divisions
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2c; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "encode_p" > $O/t_ep.log 2>&1
echo "rc=$?" >> $O/t_ep.log
timeout 900 python -m pytest tests -m gpu -q > $O/t_all.log 2>&1
echo "rc=$?" >> $O/t_all.log
B="timeout 300 python bench.py --steps 10 --warmup 3 --extras 0 --e2e 0 --cpu-budget 0.5"
$B --workload encode_p_1080p > $O/ep.json 2> $O/ep.err
PFV_ENCODE_P_VARIANT=v1 $B --workload encode_p_1080p > $O/ep_v1.json 2> $O/ep_v1.err
$B --workload decode_p_1080p > $O/dp.json 2> $O/dp.err
$B --workload decode_p_4k > $O/dp4k.json 2> $O/dp4k.err
PFV_DECODE_P_VARIANT=win $B --workload decode_p_4k > $O/dp4k_win.json 2> $O/dp4k_win.err
$B --workload encode_i_1080p > $O/ei.json 2> $O/ei.err
$B --workload decode_i_1080p > $O/di.json 2> $O/di.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:encode_p2 -s 8 -c 1 -o $O/prof_ep2 python bench.py --steps 2 --warmup 3 --extras 0 --e2e 0 --cpu-budget 0 --workload encode_p_1080p > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_p_fused -s 20 -c 1 -o $O/prof_dp_fused python bench.py --steps 2 --warmup 3 --extras 0 --e2e 0 --cpu-budget 0 --workload decode_p_1080p > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:encode_i_stream -s 3 -c 1 -o $O/prof_ei_stream python bench.py --steps 2 --warmup 3 --extras 0 --e2e 0 --cpu-budget 0 --workload encode_i_1080p > /dev/null 2>&1
ls -la $O
