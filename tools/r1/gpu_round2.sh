#!/bin/bash
# GPU-box visit: parity tests, smoke, default bench line, launch list, and ONE ncu --set full run that captures both
# decode-P kernels (mc_copy_kernel + residual_sb_kernel).  gpurun --timeout 1500 -- bash tools/gpu_round2.sh <tag>
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi_$TAG.txt 2>&1
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $OUT/pytest_$TAG.txt
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee $OUT/smoke_$TAG.txt
echo "== bench"; timeout 900 python bench.py 2> $OUT/bench_$TAG.err | tee $OUT/bench_$TAG.json
echo "== ncu launch list (decode_p workload)"
bash tools/gpu_launchlist.sh p_$TAG decode_p_1080p
echo "== ncu full decode-P"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mc_copy|residual_sb" -s 12 -c 4 -f -o $OUT/prof_decode_p_two_$TAG \
    python bench.py --workload decode_p_1080p --steps 2 --warmup 3 --extras 0 --cpu-budget 0.2 --e2e 0 > $OUT/ncu_full_p_$TAG.log 2>&1
tail -2 $OUT/ncu_full_p_$TAG.log
ls -la $OUT
