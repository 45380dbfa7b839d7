#!/bin/bash
# round 2: the new buffer-growth test and the codec tests on the final library
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2zz4; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_codec.py -m gpu -q -x > $O/t.log 2>&1
echo "rc=$?" >> $O/t.log
tail -n 3 $O/t.log
