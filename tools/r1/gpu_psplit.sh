#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_codec.py -x -q -k "pframe or long_motion or variants or stream or batched or decoder or sparse or full_size or chain" 2>&1 | tail -3
bash tools/gpu_sweep.sh decode_p_1080p PFV_DECODE_P_SPLIT 1 2 4 8
bash tools/gpu_sweep.sh decode_p_4k PFV_DECODE_P_SPLIT 1 2
