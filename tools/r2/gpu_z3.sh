#!/bin/bash
# round 2, visit z3: packed host store of the key-frame entries; the whole GPU suite three times (flakiness)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2z3; mkdir -p $O
for i in 1 2 3; do
  timeout 900 python -m pytest tests -m gpu -q -x > $O/t_all_$i.log 2>&1
  echo "rc=$?" >> $O/t_all_$i.log
done
timeout 400 python bench.py --steps 10 --warmup 3 --extras 0 --cpu-budget 0.5 --workload encode_i_1080p > $O/ei.json 2> $O/ei.err
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_codec.py -x -q -k "(sparse_encode or interleaved or encoder_stream or writer) and not 512" 2>&1 | tail -8 > $O/sanitize_tok_memcheck.txt
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_codec.py -x -q -k "(sparse_encode or encoder_stream) and not 512" 2>&1 | tail -8 > $O/sanitize_tok_racecheck.txt
tail -n 2 $O/t_all_*.log; cat $O/sanitize_tok_*.txt | tail -8
