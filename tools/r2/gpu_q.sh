#!/bin/bash
# round 2, seventeenth visit: the decode_p_4k bench failure again, then under memcheck
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2q; mkdir -p $O
timeout 300 python bench.py --steps 10 --warmup 3 --extras 0 --e2e 0 --cpu-budget 0 --workload decode_p_4k > $O/dp4k.json 2> $O/dp4k.err
echo "rc=$?" >> $O/dp4k.err
PFV_DECODE_P_VARIANT=win timeout 300 python bench.py --steps 10 --warmup 3 --extras 0 --e2e 0 --cpu-budget 0 --workload decode_p_4k > $O/dp4k_win.json 2> $O/dp4k_win.err
echo "rc=$?" >> $O/dp4k_win.err
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python bench.py --steps 2 --warmup 3 --extras 0 --e2e 0 --cpu-budget 0 --workload decode_p_4k > $O/memcheck.txt 2>&1
head -60 $O/memcheck.txt > $O/memcheck_head.txt
tail -5 $O/dp4k.err $O/dp4k_win.err; head -40 $O/memcheck.txt
