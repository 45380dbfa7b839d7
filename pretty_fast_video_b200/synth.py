"""Deterministic synthetic YUV420 content for tests and benchmarks.

Per frame t: luma = smooth gradient + 12 textured rectangles translating at integer velocities in
[-12, 12] px/frame + uniform noise; chroma = smooth half-resolution fields + noise.  The mix gives
skipped macroblocks, non-zero motion vectors up to the +-15 search limit and coded residuals.
"random" (every macroblock coded, dense coefficients) and "static" (everything skipped) are the two
stress streams.  There is no network in the build environment, so all benchmark input is made here.
"""
from __future__ import annotations

import numpy as np

SEED_I_1080P = 0x50465601
SEED_P_1080P = 0x50465602
SEED_P_4K = 0x50465603


class SynthVideo:
    def __init__(self, width: int, height: int, seed: int, kind: str = "moving", noise: int = 4):
        assert width % 2 == 0 and height % 2 == 0
        self.w, self.h, self.kind, self.noise = width, height, kind, noise
        self.seed = seed
        rng = np.random.Generator(np.random.PCG64(seed))
        self.nrect = 12
        self.rw = rng.integers(max(16, width // 16), max(24, width // 4), self.nrect)
        self.rh = rng.integers(max(16, height // 16), max(24, height // 4), self.nrect)
        self.x0 = rng.integers(0, width, self.nrect)
        self.y0 = rng.integers(0, height, self.nrect)
        self.vx = rng.integers(-12, 13, self.nrect)
        self.vy = rng.integers(-12, 13, self.nrect)
        self.tex = []
        for i in range(self.nrect):
            coarse = rng.integers(0, 256, ((self.rh[i] + 7) // 8, (self.rw[i] + 7) // 8)).astype(np.float32)
            t = np.kron(coarse, np.ones((8, 8), np.float32))[: self.rh[i], : self.rw[i]]
            fine = rng.integers(-24, 25, (self.rh[i], self.rw[i])).astype(np.float32)
            self.tex.append(np.clip(0.6 * t + 50 + fine, 0, 255))
        yy, xx = np.mgrid[0:height, 0:width].astype(np.float32)
        self.bg = 110 + 60 * np.sin(xx * (3.0 / width)) * np.cos(yy * (2.0 / height)) + 0.02 * xx
        cy, cx = np.mgrid[0:height // 2, 0:width // 2].astype(np.float32)
        self.ubg = 128 + 40 * np.sin(cx * (4.0 / width) + 1.0)
        self.vbg = 128 + 40 * np.cos(cy * (4.0 / height) - 0.5)

    def frame(self, t: int):
        w, h = self.w, self.h
        rng = np.random.Generator(np.random.PCG64([self.seed, t]))
        if self.kind == "random":
            return (rng.integers(0, 256, (h, w), dtype=np.uint8),
                    rng.integers(0, 256, (h // 2, w // 2), dtype=np.uint8),
                    rng.integers(0, 256, (h // 2, w // 2), dtype=np.uint8))
        tt = 0 if self.kind == "static" else t
        y = self.bg.copy()
        u = self.ubg.copy()
        v = self.vbg.copy()
        for i in range(self.nrect):
            x = int((self.x0[i] + self.vx[i] * tt) % w)
            yp = int((self.y0[i] + self.vy[i] * tt) % h)
            x1, y1 = min(w, x + self.rw[i]), min(h, yp + self.rh[i])
            y[yp:y1, x:x1] = self.tex[i][: y1 - yp, : x1 - x]
            u[yp // 2:y1 // 2, x // 2:x1 // 2] = 96 + 8 * (i % 5)
            v[yp // 2:y1 // 2, x // 2:x1 // 2] = 160 - 8 * (i % 7)
        if self.kind != "static" and self.noise:
            n = self.noise
            y += rng.integers(-n, n + 1, (h, w))
            u += rng.integers(-(n // 2), n // 2 + 1, (h // 2, w // 2))
            v += rng.integers(-(n // 2), n // 2 + 1, (h // 2, w // 2))
        return (np.clip(y, 0, 255).astype(np.uint8), np.clip(u, 0, 255).astype(np.uint8),
                np.clip(v, 0, 255).astype(np.uint8))
