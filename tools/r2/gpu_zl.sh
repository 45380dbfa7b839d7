#!/bin/bash
# round 2, visit zl: batched RGB conversion with 8 pixels per thread, no conversion instructions, coalesced 16-byte stores
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2zl; mkdir -p $O
timeout 900 python -m pytest tests/test_format_helpers.py -m gpu -q -x > $O/t.log 2>&1
echo "rc=$?" >> $O/t.log
timeout 600 python - > $O/rgb.json 2> $O/rgb.err <<'PY'
import json, sys, torch
sys.path.insert(0, ".")
import bench
stream = torch.cuda.Stream()
with torch.cuda.stream(stream):
    r = bench.run_format_steps(torch, bench.Dist(1), stream, 10)
print(json.dumps(r))
PY
timeout 400 ncu --set full --clock-control none --import-source on -c 1 -k regex:rgb_batch8 -o $O/prof_rgb8 python -m pytest tests/test_format_helpers.py -m gpu -q -x -k "1920" > /dev/null 2>&1
tail -n 3 $O/t.log; tail -n 5 $O/rgb.err; cat $O/rgb.json
