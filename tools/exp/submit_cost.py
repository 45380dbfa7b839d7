"""host cost of one pfv_decode_submit_sparse call vs jobs per call (diagnostic)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from pretty_fast_video_b200 import Engine, PinnedArena, make_qtables, geometry_for, PFV_FRAME_I, PFV_FRAME_P, codec
from pretty_fast_video_b200.engine import SparseDecodeJob
w, h = 1920, 1080
qt, _ = make_qtables(5)
g = geometry_for(w, h)
rng = np.random.default_rng(1)
c = np.zeros(g.nb * 256, np.int16)
idx = rng.choice(c.size, 60000, replace=False); c[idx] = rng.integers(-50, 50, idx.size)
mo, tk = codec.dense_to_tokens(c, g.nb)
hdr = np.zeros((g.nb, 4), np.uint8); hdr[:, 2] = 1
ysz, csz = w * h, (w // 2) * (h // 2)
for nj in (1, 2, 4):
    for with_out in (False, True):
        ar = PinnedArena(nj * (tk.nbytes + mo.nbytes + hdr.nbytes + ysz + 2 * csz) + 65536)
        jobs = []
        for j in range(nj):
            pmo = ar.take(mo.shape, np.uint32); pmo[...] = mo
            ptk = ar.take(tk.shape, np.uint32); ptk[...] = tk
            ph = ar.take(hdr.shape, np.uint8); ph[...] = hdr
            out = ar.take((ysz + 2 * csz,), np.uint8)
            b = out.ctypes.data
            jobs.append((pmo, ptk, ph, (b, b + ysz, b + ysz + csz)))
        with Engine(w, h, qt, nslots=2 * nj + 2, max_jobs=nj) as e:
            cur = [0] * nj
            def mk(kind):
                js = []
                for j, (pmo, ptk, ph, outs) in enumerate(jobs):
                    js.append(SparseDecodeJob(kind, 2 * j + 1 - cur[j], pmo, ptk, (0, 1, 1) if kind == PFV_FRAME_I else (2, 3, 3),
                                              ref_slot=2 * j + cur[j], hdr=ph if kind == PFV_FRAME_P else None, out=outs if with_out else None))
                    cur[j] ^= 1
                return e.build_sparse_decode_jobs(js), js
            arr, js = mk(PFV_FRAME_I); e.decode_submit_sparse(js, prebuilt=arr); e.sync()
            for kind, name in ((PFV_FRAME_I, "I"), (PFV_FRAME_P, "P")):
                pre = [mk(kind) for _ in range(40)]
                e.sync()
                t0 = time.perf_counter()
                for arr, js in pre:
                    e.decode_submit_sparse(js, prebuilt=arr)
                t1 = time.perf_counter()
                e.sync()
                t2 = time.perf_counter()
                print("jobs/call %d out=%d %s: host %.1f us per call (%.1f per job); incl. drain %.1f us per frame" % (
                    nj, with_out, name, (t1 - t0) / 40 * 1e6, (t1 - t0) / 40 / nj * 1e6, (t2 - t0) / 40 / nj * 1e6))
        ar.close()
