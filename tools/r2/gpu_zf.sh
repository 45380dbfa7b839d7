#!/bin/bash
# round 2, visit zf: encode-P 16 warps x 1 window buffer with the glue fast paths; window pitch 256 and 20 warps as variants
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2zf; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q -x -k "encode_p or encode_pframe or sparse_encode or encoder_chain or encoder_stream or round_trip" > $O/t.log 2>&1
echo "rc=$?" >> $O/t.log
B="timeout 300 python bench.py --steps 10 --warmup 3 --extras 0 --e2e 0 --cpu-budget 0.5"
$B --workload encode_p_1080p > $O/ep.json 2> $O/ep.err
cp pretty_fast_video_b200/libpfv_b200.so /tmp/new.so
for v in w256 20; do
  if [ -f tools/exp/libpfv_b200_$v.so ]; then
    cp tools/exp/libpfv_b200_$v.so pretty_fast_video_b200/libpfv_b200.so
    $B --workload encode_p_1080p > $O/ep_$v.json 2> $O/ep_$v.err
    timeout 600 python -m pytest tests -m gpu -q -x -k "encode_p or encode_pframe" > $O/t_$v.log 2>&1; echo "rc=$?" >> $O/t_$v.log
  fi
done
cp /tmp/new.so pretty_fast_video_b200/libpfv_b200.so
timeout 600 ncu --set full --clock-control none --import-source on -k regex:encode_p2 -c 1 -o $O/prof_ep2 python bench.py --workload encode_p_1080p --steps 1 --warmup 1 --extras 0 --e2e 0 --cpu-budget 0 > $O/ncu.log 2>&1
tail -n 2 $O/t*.log
