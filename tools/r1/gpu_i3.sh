#!/bin/bash
echo "== parity with 3 stages / 48-slot queue"
PFV_DECODE_I_STAGES=3 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_codec.py -x -q -k "iframe or variants or stream or config1 or full_size or sparse or decoder" 2>&1 | tail -2
bash tools/gpu_sweep.sh decode_i_1080p PFV_DECODE_I_STAGES 2 3 2 3
bash tools/gpu_sweep.sh decode_i_1080p_dense PFV_DECODE_I_STAGES 2 3
