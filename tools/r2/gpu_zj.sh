#!/bin/bash
# round 2, visit zj: decode-I work split by whole waves (tiles per warp chosen to minimise ceil(CTAs / resident) x (tiles + 1))
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2zj; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x -k "decode_i or iframe or key_frames or three_tables or smoke" > $O/t.log 2>&1
echo "rc=$?" >> $O/t.log
B="timeout 300 python bench.py --steps 20 --warmup 3 --extras 0 --e2e 0 --cpu-budget 0.5"
$B > $O/di_auto.json 2> $O/di_auto.err
for t in 9 8 10 12 15 4; do PFV_DECODE_I_TPW=$t $B > $O/di_t$t.json 2> $O/di_t$t.err; done
$B --workload decode_i_1080p_dense > $O/dd_auto.json 2> $O/dd_auto.err
for t in 3 5 9 16; do PFV_DECODE_I_TPW=$t $B --workload decode_i_1080p_dense > $O/dd_t$t.json 2> $O/dd_t$t.err; done
timeout 400 ncu --set full --clock-control none --import-source on -c 1 -k regex:decode_i_stream -s 6 -o $O/prof_decode_i python bench.py --steps 2 --warmup 1 --extras 0 --e2e 0 --cpu-budget 0 > /dev/null 2>&1
tail -n 3 $O/t.log
