#!/usr/bin/env python
"""Stress / repro for the decode-P kernels at large geometries: `lanes` frames per submit, random legal motion vectors, ~15 % coded
macroblocks, a chain of `steps` P frames on ping-pong slots; prints a checksum of every lane's final frame (compare
PFV_DECODE_P_VARIANT=fused / win / warp; run the fused one under compute-sanitizer).

  python tools/exp/dp_stress.py 3840 2160 16 20
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from pretty_fast_video_b200 import PFV_FRAME_P, Engine, make_qtables  # noqa: E402
from pretty_fast_video_b200.engine import DecodeJob  # noqa: E402

w, h, lanes, steps = (int(x) for x in (sys.argv[1:5] + ["3840", "2160", "16", "20"][len(sys.argv) - 1:]))
rng = np.random.default_rng(1234)
qt, _ = make_qtables(5)
with Engine(w, h, qt, nslots=2 * lanes, max_jobs=lanes, device=0) as e:
    g = e.geometry
    hdr = np.zeros((g.nb, 4), np.uint8)
    mv = hdr[:, :2].view(np.int8)
    idx = 0
    for (pw, ph) in ((g.pw, g.ph), (g.cpw, g.cph), (g.cpw, g.cph)):
        bw, bh = pw // 16, ph // 16
        bx = np.tile(np.arange(bw), bh) * 16
        by = np.repeat(np.arange(bh), bw) * 16
        n = bw * bh
        lo_x, hi_x = np.maximum(-15, -bx), np.minimum(15, pw - 16 - bx)
        lo_y, hi_y = np.maximum(-15, -by), np.minimum(15, ph - 16 - by)
        mv[idx:idx + n, 0] = (lo_x + (rng.random(n) * (hi_x - lo_x + 1)).astype(np.int64)).astype(np.int8)
        mv[idx:idx + n, 1] = (lo_y + (rng.random(n) * (hi_y - lo_y + 1)).astype(np.int64)).astype(np.int8)
        idx += n
    hdr[:, 2] = rng.random(g.nb) < 0.15
    coeff = (rng.integers(-40, 41, (g.nb, 256)) * (rng.random((g.nb, 256)) < 0.05)).astype(np.int16)
    coeff[hdr[:, 2] == 0] = 0
    d_hdr = torch.from_numpy(hdr).cuda()
    d_coeff = torch.from_numpy(coeff).cuda()
    ref = rng.integers(0, 256, g.frame_bytes).astype(np.uint8)
    for l in range(lanes):
        e.slot_write(2 * l, np.roll(ref, 977 * l))
    cur = [2 * l for l in range(lanes)]
    for s in range(steps):
        jobs = []
        for l in range(lanes):
            jobs.append(DecodeJob(PFV_FRAME_P, cur[l] ^ 1, d_coeff.data_ptr(), (2, 3, 3), ref_slot=cur[l], hdr=d_hdr.data_ptr(),
                                  device_ptrs=True))
            cur[l] ^= 1
        e.decode_submit(jobs)
    e.sync()
    hs = hashlib.sha1()
    for l in range(lanes):
        hs.update(e.slot_read(cur[l]).tobytes())
    print(f"{w}x{h} lanes={lanes} steps={steps} variant={os.environ.get('PFV_DECODE_P_VARIANT', 'fused')} sha1={hs.hexdigest()}")
