#!/bin/bash
# round 2, final visit: every GPU test, the bench line (both arms), the launch list of the bench command, sanitizers, tokenizer cost
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2z; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/t_all.log 2>&1
echo "rc=$?" >> $O/t_all.log
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err
echo "rc=$?" >> $O/bench.err
timeout 400 python bench.py --impl reference > $O/bench_ref.json 2> $O/bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 5 --warmup 3 --extras 0 --e2e 0 --cpu-budget 0 > /dev/null 2>&1
timeout 300 python tools/exp/tok_cost.py > $O/tok_cost.txt 2>&1
K="not 1080p and not full_size and not config1 and not size4 and not size3 and not 1918 and not two_devices"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "$K" 2>&1 | tail -8 > $O/sanitize_memcheck.txt
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "(variants or long_motion or chained or encode_pframe or encode_iframe or encode_i_kernel or three_tables) and $K" 2>&1 | tail -8 > $O/sanitize_racecheck.txt
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_codec.py -x -q -k "(sparse_encode or interleaved or encoder_stream or writer or decoder_matches) and not 512" 2>&1 | tail -8 > $O/sanitize_codec_memcheck.txt
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "(variants or chained or encode_iframe) and $K" 2>&1 | tail -8 > $O/sanitize_synccheck.txt
timeout 300 python __graft_entry__.py smoke > $O/smoke.txt 2>&1
ls -la $O; tail -n 3 $O/t_all.log; tail -n 2 $O/bench.err; cat $O/smoke.txt | tail -2
