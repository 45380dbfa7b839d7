#!/bin/bash
# round 2, visit ze: encode-P with ONE window stage per warp: 16 / 18 / 20 / 22 warps per SM
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2ze; mkdir -p $O
B="timeout 300 python bench.py --steps 10 --warmup 3 --extras 0 --e2e 0 --cpu-budget 0.5"
cp pretty_fast_video_b200/libpfv_b200.so /tmp/new.so
for v in 16x1 18x1 20x1 22x1; do
  if [ -f tools/exp/libpfv_b200_$v.so ]; then
    cp tools/exp/libpfv_b200_$v.so pretty_fast_video_b200/libpfv_b200.so
    $B --workload encode_p_1080p > $O/ep_$v.json 2> $O/ep_$v.err
    timeout 600 python -m pytest tests -m gpu -q -x -k "encode_p or encode_pframe" > $O/t_$v.log 2>&1; echo "rc=$?" >> $O/t_$v.log
  fi
done
cp /tmp/new.so pretty_fast_video_b200/libpfv_b200.so
tail -n 2 $O/t_*.log
