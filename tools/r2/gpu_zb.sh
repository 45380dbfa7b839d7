#!/bin/bash
# round 2, visit zb: encode-P fine levels with hoisted loads, window pitch 160 against 176; decode-P frame index rotated by the window index
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2zb; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q -x -k "encode_p or encode_pframe or sparse_encode or encoder_chain or encoder_stream or decode_p or pframes or fullsize or full_gop or round_trip" > $O/t.log 2>&1
echo "rc=$?" >> $O/t.log
B="timeout 300 python bench.py --steps 10 --warmup 3 --extras 0 --e2e 0 --cpu-budget 0.5"
$B --workload encode_p_1080p > $O/ep_w160.json 2> $O/ep_w160.err
$B --workload decode_p_1080p > $O/dp.json 2> $O/dp.err
$B --workload decode_p_1080p_64 > $O/dp64.json 2> $O/dp64.err
$B --workload decode_p_4k > $O/dp4k.json 2> $O/dp4k.err
cp pretty_fast_video_b200/libpfv_b200.so /tmp/new.so
if [ -f tools/exp/libpfv_b200_w176.so ]; then
  cp tools/exp/libpfv_b200_w176.so pretty_fast_video_b200/libpfv_b200.so
  $B --workload encode_p_1080p > $O/ep_w176.json 2> $O/ep_w176.err
  cp /tmp/new.so pretty_fast_video_b200/libpfv_b200.so
fi
timeout 600 ncu --set full --clock-control none --import-source on -k regex:encode_p2 -c 1 -o $O/prof_ep2 python bench.py --workload encode_p_1080p --steps 1 --warmup 1 --extras 0 --e2e 0 --cpu-budget 0 > $O/ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_p_fused -s 20 -c 1 -o $O/prof_dp python bench.py --workload decode_p_1080p --steps 2 --warmup 1 --extras 0 --e2e 0 --cpu-budget 0 > $O/ncu_dp.log 2>&1
tail -n 3 $O/t.log
