#!/bin/bash
# round 2, visit zd: encode-P warps-per-SM experiment: 16 warps with ONE window stage each (no window in flight during the search)
# against 12 warps x 1 stage and the default 12 x 2
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2zd; mkdir -p $O
B="timeout 300 python bench.py --steps 10 --warmup 3 --extras 0 --e2e 0 --cpu-budget 0.5"
$B --workload encode_p_1080p > $O/ep_12x2.json 2> $O/ep_12x2.err
cp pretty_fast_video_b200/libpfv_b200.so /tmp/new.so
for v in 16x1 12x1; do
  if [ -f tools/exp/libpfv_b200_$v.so ]; then
    cp tools/exp/libpfv_b200_$v.so pretty_fast_video_b200/libpfv_b200.so
    $B --workload encode_p_1080p > $O/ep_$v.json 2> $O/ep_$v.err
    if [ $v = 16x1 ]; then
      timeout 600 python -m pytest tests -m gpu -q -x -k "encode_p or encode_pframe" > $O/t_$v.log 2>&1; echo "rc=$?" >> $O/t_$v.log
    fi
  fi
done
cp /tmp/new.so pretty_fast_video_b200/libpfv_b200.so
tail -n 2 $O/t_16x1.log
