#!/bin/bash
# round 2, first visit: parity of the new kernels (fused decode-P, encode-I stream), full-size shapes, first numbers
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2a; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/smi.txt 2>&1
echo "== fused decode-P quick" > $O/t0.log
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "decode_pframe or variants or chained or all_skipped or bad_motion" >> $O/t0.log 2>&1
echo "rc=$?" >> $O/t0.log
echo "== encode quick" > $O/t1.log
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "encode" >> $O/t1.log 2>&1
echo "rc=$?" >> $O/t1.log
timeout 1500 python -m pytest tests -m gpu -q > $O/t_all.log 2>&1
echo "rc=$?" >> $O/t_all.log
timeout 900 python bench.py --steps 10 --warmup 3 --cpu-budget 2 > $O/bench.json 2> $O/bench.err
echo "rc=$?" >> $O/bench.err
for shape in 72 62; do
  PFV_DECODE_I_SHAPE=$shape timeout 300 python bench.py --steps 10 --warmup 3 --extras 0 --e2e 0 --cpu-budget 0.5 --workload decode_i_1080p_dense > $O/dense_$shape.json 2> $O/dense_$shape.err
  PFV_DECODE_I_SHAPE=$shape timeout 300 python bench.py --steps 10 --warmup 3 --extras 0 --e2e 0 --cpu-budget 0.5 --workload decode_i_1080p > $O/di_$shape.json 2> $O/di_$shape.err
done
PFV_DECODE_I_VARIANT=sb timeout 300 python bench.py --steps 10 --warmup 3 --extras 0 --e2e 0 --cpu-budget 0.5 --workload decode_i_1080p_dense > $O/dense_sb.json 2> $O/dense_sb.err
PFV_DECODE_P_VARIANT=win timeout 300 python bench.py --steps 10 --warmup 3 --extras 0 --e2e 0 --cpu-budget 0.5 --workload decode_p_1080p > $O/dp_win.json 2> $O/dp_win.err
PFV_ENCODE_I_VARIANT=warp timeout 300 python bench.py --steps 10 --warmup 3 --extras 0 --e2e 0 --cpu-budget 0.5 --workload encode_i_1080p > $O/ei_warp.json 2> $O/ei_warp.err
# ncu: launch list + full sets of the two new kernels
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_dp.csv python bench.py --steps 2 --warmup 3 --extras 0 --e2e 0 --cpu-budget 0 --workload decode_p_1080p > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_p_fused -s 20 -c 2 -o $O/prof_dp_fused python bench.py --steps 2 --warmup 3 --extras 0 --e2e 0 --cpu-budget 0 --workload decode_p_1080p > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:encode_i_stream -s 3 -c 1 -o $O/prof_ei_stream python bench.py --steps 2 --warmup 3 --extras 0 --e2e 0 --cpu-budget 0 --workload encode_i_1080p > /dev/null 2>&1
ls -la $O
