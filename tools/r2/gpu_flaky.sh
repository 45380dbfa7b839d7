#!/bin/bash
# round 2: the intermittent failure of the 4K round-trip test - how often, alone and after the other codec tests
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2flaky; mkdir -p $O
: > $O/loop.txt
for i in 1 2 3 4 5 6 7 8; do
  timeout 120 python -m pytest tests/test_gpu_codec.py -m gpu -q -x -k "round_trip_4k" 2>&1 | tail -n 1 >> $O/loop.txt
done
for i in 1 2 3 4; do
  timeout 300 python -m pytest tests/test_gpu_codec.py tests/test_gpu_fullsize.py -m gpu -q 2>&1 | tail -n 3 >> $O/loop.txt
done
cat $O/loop.txt
