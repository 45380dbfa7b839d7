#!/bin/bash
# round 2, eighteenth visit: the persistent encode-I kernel (tests, A/B against the grid form, ncu), everything else re-tested
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2r; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/t_all.log 2>&1
echo "rc=$?" >> $O/t_all.log
B="timeout 300 python bench.py --steps 10 --warmup 3 --extras 0 --e2e 0 --cpu-budget 0.5"
$B --workload encode_i_1080p > $O/ei_persist.json 2> $O/ei_persist.err
PFV_ENCODE_I_VARIANT=stream $B --workload encode_i_1080p > $O/ei_stream.json 2> $O/ei_stream.err
$B --workload decode_p_1080p > $O/dp.json 2> $O/dp.err
$B --workload decode_p_4k > $O/dp4k.json 2> $O/dp4k.err
$B --workload encode_p_1080p > $O/ep.json 2> $O/ep.err
K="not 1080p and not full_size and not config1 and not size4 and not size3 and not 1918 and not two_devices"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "(encode_iframe or encode_i_kernel) and $K" 2>&1 | tail -8 > $O/sanitize_racecheck_ei.txt
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_codec.py -x -q -k "(encode_iframe or encode_i_kernel or sparse_encode or encoder_stream) and $K and not 512" 2>&1 | tail -8 > $O/sanitize_memcheck_ei.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:encode_i_persist -s 3 -c 1 -o $O/prof_ei_persist python bench.py --steps 2 --warmup 3 --extras 0 --e2e 0 --cpu-budget 0 --workload encode_i_1080p > /dev/null 2>&1
ls -la $O
