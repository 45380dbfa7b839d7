#!/usr/bin/env python
"""A/B of the small-submit path (PFV_LEAN_SUBMIT=0/1 in the environment): config 1 (512x384, 161 frames) through
Decoder.advance_frame and the pageable one-frame-per-call drop-in leg of bench.py, three passes each."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch

import bench

dist = bench.Dist(1)
torch.cuda.set_device(0)
stream = torch.cuda.Stream()
out = {"lean": os.environ.get("PFV_LEAN_SUBMIT", "1"), "config1": [], "pageable": []}
for _ in range(3):
    r = bench.run_config1_stream(torch, dist, 0.0, 16)
    out["config1"].append(round(r["value"]))
    r = bench.run_pageable_drop_in(torch, dist, stream)
    out["pageable"].append(round(r["value"]))
print(json.dumps(out))
dist.close()
