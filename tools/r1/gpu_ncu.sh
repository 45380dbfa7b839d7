#!/bin/bash
# ncu --set full of kernels matching a regex on one bench workload:
#   tools/gpu_ncu.sh <tag> <kernel-regex> <workload> [skip] [count]   (extra env vars are passed through)
TAG=$1; K=$2; WL=$3; SK=${4:-6}; CNT=${5:-2}
OUT=gpurun_out; mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$K" -s $SK -c $CNT -f -o $OUT/prof_$TAG \
    python bench.py --workload $WL --steps 2 --warmup 3 --extras 0 --cpu-budget 0.1 --e2e 0 > $OUT/ncu_full_$TAG.log 2>&1
tail -2 $OUT/ncu_full_$TAG.log
