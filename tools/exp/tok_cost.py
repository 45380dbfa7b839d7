"""Device time of the sparse encode seam next to the dense one: L x 1080p frames per submit (default 32; the Encoder object
   submits ONE), everything resident.
   python tools/exp/tok_cost.py [L]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from pretty_fast_video_b200 import PFV_FRAME_I, PFV_FRAME_P, Engine, make_qtables, geometry_for
from pretty_fast_video_b200 import _native as N
from pretty_fast_video_b200.engine import EncodeJob, SparseEncodeJob
from pretty_fast_video_b200.synth import SynthVideo

w, h, L = 1920, 1080, (int(sys.argv[1]) if len(sys.argv) > 1 else 32)
qt, px_err = make_qtables(5)
geo = geometry_for(w, h)
nb = geo.nb
sv = SynthVideo(w, h, 0x50465602)
ysz, csz = w * h, (w // 2) * (h // 2)
f = [sv.frame(t) for t in range(3)]
def dev_planes(fr):
    t = torch.from_numpy(np.concatenate([p.reshape(-1) for p in fr])).cuda()
    return t
d0, d1, d2 = dev_planes(f[0]), dev_planes(f[1]), dev_planes(f[2])
d_c = torch.zeros((L, nb * 256), dtype=torch.int16, device="cuda")
d_h = torch.zeros((L, nb, 4), dtype=torch.uint8, device="cuda")
d_tok = torch.zeros((L, nb * 256), dtype=torch.int32, device="cuda")
d_st = torch.zeros((L, 64), dtype=torch.int32, device="cuda")
with Engine(w, h, qt, nslots=2 * L + 2, max_jobs=L) as e:
    def planes(t): b = t.data_ptr(); return (b, b + ysz, b + ysz + csz)
    e.enable_kernel_timing()
    e.encode_submit([EncodeJob(PFV_FRAME_I, 2 * i, planes(d0), d_c[i].data_ptr(), device_ptrs=True) for i in range(L)])
    e.sync()
    for kind, src, name in ((PFV_FRAME_P, d1, "P"), (PFV_FRAME_I, d2, "I")):
        dense = [EncodeJob(kind, 2 * i + 1, planes(src), d_c[i].data_ptr(), ref_slot=2 * i, px_err=px_err, hdr_out=d_h[i].data_ptr(), device_ptrs=True) for i in range(L)]
        sparse = [SparseEncodeJob(kind, 2 * i + 1, planes(src), d_tok[i].data_ptr(), d_st[i].data_ptr(), tok_cap=nb * 256, ref_slot=2 * i, px_err=px_err,
                                  hdr_out=d_h[i].data_ptr(), device_ptrs=True) for i in range(L)]
        for label, fn, jobs in (("dense ", e.encode_submit, dense), ("sparse", e.encode_submit_sparse, sparse)):
            ms = []
            for rep in range(6):
                fn(jobs); e.sync(); ms.append(e.last_kernel_ms())
            print(f"{name} frames x{L} {label}: compute-stream device time {min(ms[2:])*1e3:8.1f} us  ({L / (min(ms[2:])*1e-3):9.0f} frames/s)")
        print("   entries per frame:", int(d_st[0, N.PFV_TOKSTATS_NTOK]))
