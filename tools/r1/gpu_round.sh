#!/bin/bash
# One GPU-box visit: parity tests, the default bench line, the ncu launch list of the same command and one
# `--set full` capture of each hot kernel.  Run as: gpurun --timeout 1500 -- bash tools/gpu_round.sh <tag>
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi_$TAG.txt 2>&1
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $OUT/pytest_$TAG.txt
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee $OUT/smoke_$TAG.txt
echo "== bench"; timeout 900 python bench.py 2> $OUT/bench_$TAG.err | tee $OUT/bench_$TAG.json
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --cpu-budget 0.5 > $OUT/ncu_launch_$TAG.log 2>&1
for K in ${KERNELS:-decode_i_stream mc_copy residual_sb encode_p_kernel}; do
  WL=decode_i_1080p; SK=3
  case $K in mc_copy*|residual_sb*|decode_p*) WL=decode_p_1080p; SK=6;; encode_p*) WL=encode_p_1080p; SK=6;; esac
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s $SK -c 2 -f -o $OUT/prof_${K}_$TAG \
      python bench.py --workload $WL --steps 2 --warmup 3 --extras 0 --cpu-budget 0.2 > $OUT/ncu_full_${K}_$TAG.log 2>&1
done
ls -la $OUT
