#!/bin/bash
# round 2, fourteenth visit: pictures back in one copy (padded plane strides in the caller's buffer) on two copy streams; key-frame
# workloads at four launches per step; the bench line
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2n; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x -k "not full_size and not fullsize and not 4k" > $O/t_quick.log 2>&1
echo "rc=$?" >> $O/t_quick.log
PFV_D2H_STREAMS=1 timeout 600 python bench.py --extras 0 --cpu-budget 1 > $O/bench_1stream.json 2> $O/bench_1stream.err
timeout 600 python bench.py --extras 0 --cpu-budget 1 > $O/bench_2streams.json 2> $O/bench_2streams.err
PFV_D2H_STREAMS=1 timeout 600 python bench.py --extras 0 --cpu-budget 1 --workload decode_p_1080p > $O/dp_1stream.json 2> $O/dp_1stream.err
timeout 600 python bench.py --extras 0 --cpu-budget 1 --workload decode_p_1080p > $O/dp_2streams.json 2> $O/dp_2streams.err
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
ls -la $O
