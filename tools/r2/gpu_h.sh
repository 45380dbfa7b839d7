#!/bin/bash
# round 2, eighth visit: compute-sanitizer over the new kernels, the launch list and full captures of the bench command
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2h; mkdir -p $O
K="not 1080p and not full_size and not config1 and not size4 and not size3 and not 1918 and not two_devices"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "$K" 2>&1 | tail -8 > $O/sanitize_memcheck.txt
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "(variants or long_motion or chained or encode_pframe or encode_iframe or encode_i_kernel) and $K" 2>&1 | tail -8 > $O/sanitize_racecheck.txt
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_codec.py -x -q -k "(sparse_encode or interleaved or encoder_stream or writer or decoder_matches) and not 512" 2>&1 | tail -8 > $O/sanitize_codec_memcheck.txt
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "(variants or chained) and $K" 2>&1 | tail -8 > $O/sanitize_synccheck.txt
# launch list of the bench command (kernel share of the step) and full captures of the headline kernel
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 5 --warmup 3 --extras 0 --e2e 0 --cpu-budget 0 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_i_stream -s 3 -c 2 -o $O/prof_decode_i python bench.py --steps 2 --warmup 3 --extras 0 --e2e 0 --cpu-budget 0 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_i_sb -s 3 -c 1 -o $O/prof_decode_i_dense python bench.py --steps 2 --warmup 3 --extras 0 --e2e 0 --cpu-budget 0 --workload decode_i_1080p_dense > /dev/null 2>&1
timeout 300 python tools/exp/tok_cost.py > $O/tok_cost.txt 2>&1
ls -la $O
