#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "pframe or long_motion or variants or batched or full_size" 2>&1 | tail -2
bash tools/gpu_sweep.sh decode_p_1080p PFV_DECODE_P_SPLIT 1
bash tools/gpu_sweep.sh decode_i_1080p PFV_TILES_PER_WARP 6 8 12 16 24 32 64
bash tools/gpu_launchlist.sh pm1 decode_p_1080p 2>&1 | grep "mc_copy4\|residual"
