#!/bin/bash
# round 2, sixteenth visit: repro of the 4K x 16 decode-P failure of the dynamic window hand-out
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2p; mkdir -p $O
for v in fused win; do
  PFV_DECODE_P_VARIANT=$v timeout 300 python tools/exp/dp_stress.py 3840 2160 16 20 >> $O/stress.txt 2>> $O/stress.err
  PFV_DECODE_P_VARIANT=$v timeout 300 python tools/exp/dp_stress.py 1920 1080 32 20 >> $O/stress.txt 2>> $O/stress.err
  PFV_DECODE_P_VARIANT=$v timeout 300 python tools/exp/dp_stress.py 3840 2160 2 6 >> $O/stress.txt 2>> $O/stress.err
done
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/exp/dp_stress.py 3840 2160 16 3 > $O/memcheck_4k.txt 2>&1
tail -30 $O/memcheck_4k.txt > $O/memcheck_4k_tail.txt
ls -la $O; cat $O/stress.txt; tail -5 $O/stress.err
