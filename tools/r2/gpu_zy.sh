#!/bin/bash
# round 2, visit zy: a submit's pictures as ONE strided D2H copy (equally spaced slots and host buffers) against one copy per picture
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2zy; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x -k "decode or decoder or round_trip or key_frames or full_gop" > $O/t.log 2>&1
echo "rc=$?" >> $O/t.log
B="timeout 400 python bench.py --steps 20 --warmup 3 --extras 0 --cpu-budget 0.5"
$B > $O/di_batch.json 2> $O/di_batch.err
PFV_D2H_BATCH=0 $B > $O/di_single.json 2> $O/di_single.err
$B > $O/di_batch2.json 2> $O/di_batch2.err
$B --workload decode_p_1080p > $O/dp_batch.json 2> $O/dp_batch.err
PFV_D2H_BATCH=0 $B --workload decode_p_1080p > $O/dp_single.json 2> $O/dp_single.err
tail -n 3 $O/t.log
