#!/bin/bash
# round 2, visit x: the tokenizer kernels under ncu (key frames, 32 x 1080p per submit)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2x; mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tok_emit -s 8 -c 1 -o $O/prof_tok_emit_i python tools/exp/tok_cost.py > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches_tok.csv python tools/exp/tok_cost.py > /dev/null 2>&1
ls -la $O
