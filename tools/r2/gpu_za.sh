#!/bin/bash
# round 2, visit za: encode-P search with the error taken apart (sum a^2 + sum b^2 - 2 sum a b), A/B against the previous library
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2za; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x -k "encode_p or encode_pframe or sparse_encode or encoder_chain or encoder_stream" > $O/t_ep.log 2>&1
echo "rc=$?" >> $O/t_ep.log
B="timeout 300 python bench.py --steps 10 --warmup 3 --extras 0 --e2e 0 --cpu-budget 0.5"
$B --workload encode_p_1080p > $O/ep_new.json 2> $O/ep_new.err
$B --workload encode_p_1080p > $O/ep_new2.json 2> $O/ep_new2.err
if [ -f tools/exp/libpfv_b200_base.so ]; then
  cp pretty_fast_video_b200/libpfv_b200.so /tmp/new.so
  cp tools/exp/libpfv_b200_base.so pretty_fast_video_b200/libpfv_b200.so
  $B --workload encode_p_1080p > $O/ep_base.json 2> $O/ep_base.err
  cp /tmp/new.so pretty_fast_video_b200/libpfv_b200.so
fi
timeout 600 ncu --set full --clock-control none --import-source on -k regex:encode_p2 -c 1 -o $O/prof_ep2 python bench.py --workload encode_p_1080p --steps 1 --warmup 1 --extras 0 --e2e 0 --cpu-budget 0 > $O/ncu.log 2>&1
tail -n 3 $O/t_ep.log
