#!/bin/bash
# round 2: the whole GPU suite twice more (data points for the one intermittent failure)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2zz_twice; mkdir -p $O
for i in 1 2 3; do timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -n 4 >> $O/t.log; done
cat $O/t.log
