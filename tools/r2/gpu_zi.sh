#!/bin/bash
# round 2, visit zi: fused decode-P with a 160-byte window pitch; ONE window per pipeline with 4 and 5 pipelines
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2zi; mkdir -p $O
B="timeout 300 python bench.py --steps 10 --warmup 3 --extras 0 --e2e 0 --cpu-budget 0.5"
$B --workload decode_p_1080p > $O/dp_base.json 2> $O/dp_base.err
$B --workload decode_p_4k > $O/dp4k_base.json 2> $O/dp4k_base.err
cp pretty_fast_video_b200/libpfv_b200.so /tmp/new.so
for v in pfw160 pf4x1 pf5x1; do
  if [ -f tools/exp/libpfv_b200_$v.so ]; then
    cp tools/exp/libpfv_b200_$v.so pretty_fast_video_b200/libpfv_b200.so
    $B --workload decode_p_1080p > $O/dp_$v.json 2> $O/dp_$v.err
    $B --workload decode_p_4k > $O/dp4k_$v.json 2> $O/dp4k_$v.err
    timeout 600 python -m pytest tests -m gpu -q -x -k "decode_p or pframes or full_gop" > $O/t_$v.log 2>&1; echo "rc=$?" >> $O/t_$v.log
  fi
done
cp /tmp/new.so pretty_fast_video_b200/libpfv_b200.so
tail -n 2 $O/t_*.log
