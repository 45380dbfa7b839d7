#!/bin/bash
# round 2: the intermittent failure of the 4K round-trip test - repetitions that compare both sides with the oracle
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2flaky; mkdir -p $O
: > $O/flaky.txt
for th in 4 2 8 3 4; do timeout 200 python tools/exp/flaky_4k.py 30 $th >> $O/flaky.txt 2>&1; done
cat $O/flaky.txt | cut -c1-300
