#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_codec.py -x -q -k "pframe or long_motion or variants or stream or batched or decoder or sparse or full_size or chain" 2>&1 | tail -2
bash tools/gpu_sweep.sh decode_p_1080p PFV_RESIDUAL_VARIANT 4 2
bash tools/gpu_launchlist.sh pr4 decode_p_1080p 2>&1 | grep "mc_copy4\|residual"
