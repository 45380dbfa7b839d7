#!/bin/bash
# round 2, visit zq: RGB paths after the single-picture call became a batch of one; an ncu capture of encode-I with source
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2zq; mkdir -p $O
timeout 900 python -m pytest tests/test_format_helpers.py tests/test_gpu_codec.py -m gpu -q -x > $O/t.log 2>&1
echo "rc=$?" >> $O/t.log
timeout 400 ncu --set full --clock-control none --import-source on -c 1 -k regex:encode_i_persist -o $O/prof_ei python bench.py --workload encode_i_1080p --steps 2 --warmup 1 --extras 0 --e2e 0 --cpu-budget 0 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -c 1 -k regex:decode_i_direct -o $O/prof_dd python bench.py --workload decode_i_1080p_dense --steps 2 --warmup 1 --extras 0 --e2e 0 --cpu-budget 0 > /dev/null 2>&1
tail -n 3 $O/t.log; ls -la $O
