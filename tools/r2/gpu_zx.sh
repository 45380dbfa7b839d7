#!/bin/bash
# round 2, visit zx: fused decode-P, the CTA's windows handed to its pipelines dynamically (shared-memory counter)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2zx; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -x -k "decode_p or pframes or full_gop or round_trip or gop_sharded or smoke" > $O/t.log 2>&1
echo "rc=$?" >> $O/t.log
B="timeout 300 python bench.py --steps 10 --warmup 3 --extras 0 --e2e 0 --cpu-budget 0.5"
$B --workload decode_p_1080p > $O/dp.json 2> $O/dp.err
$B --workload decode_p_1080p_64 > $O/dp64.json 2> $O/dp64.err
$B --workload decode_p_4k > $O/dp4k.json 2> $O/dp4k.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:decode_p_fused -s 20 -c 1 -o $O/prof_dp python bench.py --workload decode_p_1080p --steps 2 --warmup 1 --extras 0 --e2e 0 --cpu-budget 0 > /dev/null 2>&1
tail -n 3 $O/t.log
