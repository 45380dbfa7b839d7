#!/bin/bash
# round 2, visit zv (round-end visit after the Decoder / Encoder host-side work: the code as committed): every GPU test, the bench line (both arms), the launch list of the
# bench command, ncu captures of the kernels the roofline lines quote (traffic.json), sanitizers over the small parity tests
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2zv; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/t_all.log 2>&1
echo "rc=$?" >> $O/t_all.log
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err
echo "rc=$?" >> $O/bench.err
timeout 400 python bench.py --impl reference > $O/bench_ref.json 2> $O/bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 5 --warmup 3 --extras 0 --e2e 0 --cpu-budget 0 > /dev/null 2>&1
N="timeout 400 ncu --set full --clock-control none --import-source on -c 1"
Q="--steps 2 --warmup 1 --extras 0 --e2e 0 --cpu-budget 0"
$N -k regex:decode_i_stream -s 6 -o $O/prof_decode_i python bench.py $Q > /dev/null 2>&1
$N -k regex:decode_p_fused -s 20 -o $O/prof_dp_1080p python bench.py --workload decode_p_1080p $Q > /dev/null 2>&1
$N -k regex:encode_p2 -o $O/prof_ep2 python bench.py --workload encode_p_1080p $Q > /dev/null 2>&1
$N -k regex:rgb_batch8 -o $O/prof_rgb8 python -m pytest tests/test_format_helpers.py -m gpu -q -x -k 1920 > /dev/null 2>&1
timeout 300 python tools/exp/tok_cost.py > $O/tok_cost.txt 2>&1
K="not 1080p and not full_size and not config1 and not size4 and not size3 and not 1918 and not two_devices"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "(encode_pframe or encode_p_kernel or encoder_chain) and $K" 2>&1 | tail -8 > $O/sanitize_memcheck_encode_p.txt
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "(encode_pframe or encode_p_kernel) and $K" 2>&1 | tail -8 > $O/sanitize_racecheck_encode_p.txt
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "(encode_pframe) and $K" 2>&1 | tail -8 > $O/sanitize_synccheck_encode_p.txt
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_format_helpers.py -m gpu -x -q -k "not 1920" 2>&1 | tail -8 > $O/sanitize_memcheck_rgb.txt
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_format_helpers.py -m gpu -x -q -k "not 1920" 2>&1 | tail -8 > $O/sanitize_racecheck_rgb.txt
timeout 300 python tools/exp/enc_trace.py > $O/enc_trace.txt 2>&1
timeout 300 python __graft_entry__.py smoke > $O/smoke.txt 2>&1
ls -la $O; tail -n 3 $O/t_all.log; tail -n 2 $O/bench.err; cat $O/smoke.txt | tail -2
