#!/bin/bash
# round 2, thirteenth visit: L2 prefetch beyond the staged tiles / windows (decode-I stream, decode-P fused, encode-I), the
# cheaper candidate masks of encode-P; all GPU tests again
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2m; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/t_all.log 2>&1
echo "rc=$?" >> $O/t_all.log
B="timeout 300 python bench.py --steps 10 --warmup 3 --extras 0 --e2e 0 --cpu-budget 0.5"
for wl in decode_i_1080p decode_p_1080p decode_p_4k encode_i_1080p encode_p_1080p decode_i_1080p_dense; do
  $B --workload $wl > $O/$wl.json 2> $O/$wl.err
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_i_stream -s 3 -c 1 -o $O/prof_decode_i python bench.py --steps 2 --warmup 3 --extras 0 --e2e 0 --cpu-budget 0 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_p_fused -s 20 -c 1 -o $O/prof_dp_fused python bench.py --steps 2 --warmup 3 --extras 0 --e2e 0 --cpu-budget 0 --workload decode_p_1080p > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:encode_i_stream -s 3 -c 1 -o $O/prof_ei python bench.py --steps 2 --warmup 3 --extras 0 --e2e 0 --cpu-budget 0 --workload encode_i_1080p > /dev/null 2>&1
ls -la $O
