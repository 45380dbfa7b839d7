#!/bin/bash
# round 2, last visit: every GPU test and the smoke test on the code as committed
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2zz; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/t_all.log 2>&1
echo "rc=$?" >> $O/t_all.log
timeout 300 python __graft_entry__.py smoke > $O/smoke.txt 2>&1
tail -n 3 $O/t_all.log; tail -n 1 $O/smoke.txt
