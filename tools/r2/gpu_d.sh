#!/bin/bash
# round 2, fourth visit: fused decode-P with three copy pipelines per CTA; the whole bench line
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2d; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -q -x -k "decode_p or variants or chained or gops or gop_4k or all_skipped or bad_motion" > $O/t_dp.log 2>&1
echo "rc=$?" >> $O/t_dp.log
timeout 300 python -m pytest tests/test_gpu_codec.py -q -x -k "writer or round_trip or decoder_matches" > $O/t_codec.log 2>&1
echo "rc=$?" >> $O/t_codec.log
B="timeout 300 python bench.py --steps 10 --warmup 3 --extras 0 --e2e 0 --cpu-budget 0.5"
$B --workload decode_p_1080p > $O/dp.json 2> $O/dp.err
$B --workload decode_p_4k > $O/dp4k.json 2> $O/dp4k.err
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err
echo "rc=$?" >> $O/bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_p_fused -s 20 -c 1 -o $O/prof_dp_fused python bench.py --steps 2 --warmup 3 --extras 0 --e2e 0 --cpu-budget 0 --workload decode_p_1080p > /dev/null 2>&1
ls -la $O
