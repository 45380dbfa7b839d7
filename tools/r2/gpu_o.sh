#!/bin/bash
# round 2, fifteenth visit: windows / tiles handed out by a device counter (decode-P fused, encode-P); realistic PCIe ceiling
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2o; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/t_all.log 2>&1
echo "rc=$?" >> $O/t_all.log
B="timeout 300 python bench.py --steps 10 --warmup 3 --extras 0 --e2e 0 --cpu-budget 0.5"
for wl in decode_p_1080p decode_p_4k encode_p_1080p; do
  $B --workload $wl > $O/$wl.json 2> $O/$wl.err
done
K="not 1080p and not full_size and not config1 and not size4 and not size3 and not 1918 and not two_devices"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "(variants or long_motion or chained or encode_pframe or decode_pframe) and $K" 2>&1 | tail -8 > $O/sanitize_racecheck.txt
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "(chained or encode_pframe or decode_pframe or decode_iframe) and $K" 2>&1 | tail -8 > $O/sanitize_memcheck.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_p_fused -s 20 -c 1 -o $O/prof_dp_fused python bench.py --steps 2 --warmup 3 --extras 0 --e2e 0 --cpu-budget 0 --workload decode_p_1080p > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:encode_p2 -s 6 -c 1 -o $O/prof_ep2 python bench.py --steps 2 --warmup 3 --extras 0 --e2e 0 --cpu-budget 0 --workload encode_p_1080p > /dev/null 2>&1
timeout 600 python bench.py --extras 0 --cpu-budget 1 > $O/bench_main.json 2> $O/bench_main.err
ls -la $O
