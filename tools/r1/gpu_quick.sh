#!/bin/bash
# quick GPU visit: parity tests + one bench line (no profiling)
TAG=${1:-q}; shift
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/pytest_$TAG.txt
timeout 900 python bench.py "$@" 2> $OUT/bench_$TAG.err | tee $OUT/bench_$TAG.json
tail -5 $OUT/bench_$TAG.err
