#!/bin/bash
# round 2, visit y: the key-frame tokenizer (thread per sub-block, padded slots, gather in the store)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2y; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/t_all.log 2>&1
echo "rc=$?" >> $O/t_all.log
timeout 300 python tools/exp/tok_cost.py > $O/tok_cost.txt 2>&1
K="not 1080p and not full_size and not config1 and not size4 and not size3 and not 1918 and not two_devices"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_codec.py -x -q -k "(sparse_encode or interleaved or encoder_stream or writer) and not 512" 2>&1 | tail -8 > $O/sanitize_tok_memcheck.txt
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_codec.py -x -q -k "(sparse_encode or encoder_stream) and not 512" 2>&1 | tail -8 > $O/sanitize_tok_racecheck.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tok_emit_sb -s 3 -c 1 -o $O/prof_tok_emit_sb python tools/exp/tok_cost.py > /dev/null 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --extras 0 --cpu-budget 0.5 --workload encode_i_1080p > $O/ei.json 2> $O/ei.err
ls -la $O; cat $O/tok_cost.txt
