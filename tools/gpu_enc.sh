#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_codec.py -x -q -k "encode or encoder or chain or round_trip or config1" 2>&1 | tail -3
bash tools/gpu_sweep.sh encode_p_1080p PFV_NOP 0
