#!/bin/bash
# compute-sanitizer over the sparse encode seam (tokenizer kernels write pinned host memory) and the re-worked encode-P kernel
OUT=gpurun_out; mkdir -p $OUT
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_codec.py tests/test_gpu_parity.py -x -q \
    -k "(sparse_encode or interleaved or encode_pframe or encoder_stream) and not size3 and not size4 and not 1918 and not 512" 2>&1 | tail -6 | tee $OUT/sanitize_tok_memcheck.txt
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_codec.py tests/test_gpu_parity.py -x -q \
    -k "(sparse_encode or encode_pframe) and not size3 and not size4 and not 1918 and not 512" 2>&1 | tail -6 | tee $OUT/sanitize_tok_racecheck.txt
