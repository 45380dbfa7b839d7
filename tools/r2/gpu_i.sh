#!/bin/bash
# round 2, ninth visit: transform warps per CTA in the fused decode-P kernel, fp32 forward transform in encode-I, racecheck again
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2i; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/t_all.log 2>&1
echo "rc=$?" >> $O/t_all.log
B="timeout 300 python bench.py --steps 10 --warmup 3 --extras 0 --e2e 0 --cpu-budget 0.5"
for xf in 4 6 8; do
  PFV_PF_XF=$xf $B --workload decode_p_1080p > $O/dp_xf$xf.json 2> $O/dp_xf$xf.err
  PFV_PF_XF=$xf $B --workload decode_p_4k > $O/dp4k_xf$xf.json 2> $O/dp4k_xf$xf.err
done
$B --workload encode_i_1080p > $O/ei.json 2> $O/ei.err
K="not 1080p and not full_size and not config1 and not size4 and not size3 and not 1918 and not two_devices"
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "(variants or long_motion or chained or encode_pframe or encode_iframe or encode_i_kernel) and $K" 2>&1 | tail -8 > $O/sanitize_racecheck.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_p_fused -s 20 -c 1 -o $O/prof_dp_fused python bench.py --steps 2 --warmup 3 --extras 0 --e2e 0 --cpu-budget 0 --workload decode_p_1080p > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:encode_i_stream -s 3 -c 1 -o $O/prof_ei_stream python bench.py --steps 2 --warmup 3 --extras 0 --e2e 0 --cpu-budget 0 --workload encode_i_1080p > /dev/null 2>&1
ls -la $O
