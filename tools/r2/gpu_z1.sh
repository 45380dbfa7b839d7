#!/bin/bash
# round 2, visit z1: key-frame tokenizer with entries staged in shared memory
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2z2; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x -k "sparse or encoder or smoke or codec or interleaved" > $O/t_tok.log 2>&1
echo "rc=$?" >> $O/t_tok.log
timeout 300 python tools/exp/tok_cost.py > $O/tok_cost.txt 2>&1
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_codec.py -x -q -k "(sparse_encode or encoder_stream) and not 512" 2>&1 | tail -8 > $O/sanitize_tok_racecheck.txt
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_codec.py -x -q -k "(sparse_encode or interleaved or encoder_stream or writer) and not 512" 2>&1 | tail -8 > $O/sanitize_tok_memcheck.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tok_emit_sb -s 3 -c 1 -o $O/prof_tok_emit_sb python tools/exp/tok_cost.py > /dev/null 2>&1
cat $O/tok_cost.txt; tail -n 3 $O/t_tok.log; cat $O/sanitize_tok_racecheck.txt $O/sanitize_tok_memcheck.txt
