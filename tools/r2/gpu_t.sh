#!/bin/bash
# round 2, twentieth visit: the persistent decode-I kernel (tests, A/B against the grid form, racecheck, ncu)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2t; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/t_all.log 2>&1
echo "rc=$?" >> $O/t_all.log
B="timeout 300 python bench.py --steps 10 --warmup 3 --extras 0 --e2e 0 --cpu-budget 0.5"
$B --workload decode_i_1080p > $O/di_persist.json 2> $O/di_persist.err
PFV_DECODE_I_VARIANT=stream $B --workload decode_i_1080p > $O/di_stream.json 2> $O/di_stream.err
$B --workload decode_i_1080p > $O/di_persist2.json 2> $O/di_persist2.err
K="not 1080p and not full_size and not config1 and not size4 and not size3 and not 1918 and not two_devices"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "(decode_iframe or variants) and $K" 2>&1 | tail -8 > $O/sanitize_racecheck_di.txt
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "(decode_iframe or variants or chained) and $K" 2>&1 | tail -8 > $O/sanitize_memcheck_di.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_i_persist -s 3 -c 1 -o $O/prof_di_persist python bench.py --steps 2 --warmup 3 --extras 0 --e2e 0 --cpu-budget 0 > /dev/null 2>&1
ls -la $O
