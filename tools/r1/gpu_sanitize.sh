#!/bin/bash
# compute-sanitizer over the small-size parity tests (memcheck, then racecheck on a subset)
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "not 1080p and not full_size and not config1 and not size4 and not size3" 2>&1 | tail -6 | tee $OUT/sanitize_memcheck.txt
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "(variants or long_motion or host_compaction or encode_pframe) and not size4 and not size3" 2>&1 | tail -6 | tee $OUT/sanitize_racecheck.txt
