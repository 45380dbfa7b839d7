#!/bin/bash
# ncu --set full of one kernel on one bench workload: tools/gpu_prof.sh <tag> <kernel-regex> <workload> [skip]
TAG=$1; K=$2; WL=$3; SK=${4:-3}
OUT=gpurun_out; mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s $SK -c 2 -f -o $OUT/prof_$TAG \
    python bench.py --workload $WL --steps 2 --warmup 3 --extras 0 --cpu-budget 0.2 > $OUT/ncu_full_$TAG.log 2>&1
tail -3 $OUT/ncu_full_$TAG.log
