#!/bin/bash
# round 2, tenth visit: encode-I (fp32 quantiser, plane-specialised loop, rolled inverse transform), the cp.async-staged dense
# decode-I kernel, the lean small-submit path
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2j; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/t_all.log 2>&1
echo "rc=$?" >> $O/t_all.log
B="timeout 300 python bench.py --steps 10 --warmup 3 --extras 0 --e2e 0 --cpu-budget 0.5"
PFV_ENCODE_I_ROLLED=1 $B --workload encode_i_1080p > $O/ei_rolled.json 2> $O/ei_rolled.err
PFV_ENCODE_I_ROLLED=0 $B --workload encode_i_1080p > $O/ei_flat.json 2> $O/ei_flat.err
$B --workload decode_i_1080p_dense > $O/di_dense.json 2> $O/di_dense.err
PFV_DECODE_I_VARIANT=sb $B --workload decode_i_1080p_dense > $O/di_dense_sb.json 2> $O/di_dense_sb.err
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:encode_i_stream -s 3 -c 1 -o $O/prof_ei_rolled python bench.py --steps 2 --warmup 3 --extras 0 --e2e 0 --cpu-budget 0 --workload encode_i_1080p > /dev/null 2>&1
PFV_ENCODE_I_ROLLED=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:encode_i_stream -s 3 -c 1 -o $O/prof_ei_flat python bench.py --steps 2 --warmup 3 --extras 0 --e2e 0 --cpu-budget 0 --workload encode_i_1080p > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_i_direct -s 3 -c 1 -o $O/prof_di_direct python bench.py --steps 2 --warmup 3 --extras 0 --e2e 0 --cpu-budget 0 --workload decode_i_1080p_dense > /dev/null 2>&1
ls -la $O
