#!/bin/bash
# round 2, visit v: persistent encode-I with the chunk counter two chunks ahead and the job record prefetched; waves of the decode-I grid
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2v; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x -k "encode or encoder or smoke or fullsize or 64_key" > $O/t_enc.log 2>&1
echo "rc=$?" >> $O/t_enc.log
B="timeout 300 python bench.py --steps 10 --warmup 3 --extras 0 --e2e 0 --cpu-budget 0.5"
$B --workload encode_i_1080p > $O/ei.json 2> $O/ei.err
for wv in 6 4 9 14; do
  PFV_DECODE_I_WAVES=$wv $B --workload decode_i_1080p > $O/di_w$wv.json 2> $O/di_w$wv.err
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:encode_i_persist -s 3 -c 1 -o $O/prof_ei_persist python bench.py --steps 2 --warmup 3 --extras 0 --e2e 0 --cpu-budget 0 --workload encode_i_1080p > /dev/null 2>&1
ls -la $O
