"""Generates tests/golden/*.npz.  The reference cannot run here (no Rust toolchain; its fixtures are LFS
stubs), so these vectors come from the C oracle AFTER it was cross-checked against the independent Python
restatement (oracle/pfv_ref.py) and SURVEY.md Appendix C; they freeze that state so that later edits to the
oracle or the kernels cannot drift silently.  Run:  python tests/golden/make_golden.py
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import pfvo  # noqa: E402
from pretty_fast_video_b200.synth import SynthVideo  # noqa: E402

# KAT-A input: the literal 8x8 test block of the reference's test_dct_encode (src/lib.rs:65); the reference only
# prints its result, SURVEY Appendix C holds the values derived for it.
KAT_A_PX = [44, 42, 43, 43, 46, 49, 42, 33, 36, 49, 56, 47, 42, 41, 36, 28, 36, 48, 57, 52, 42, 35, 29, 23,
            36, 35, 41, 48, 45, 32, 25, 24, 32, 27, 30, 39, 41, 32, 25, 26, 26, 27, 29, 30, 31, 31, 27, 23,
            29, 27, 27, 27, 30, 31, 26, 20, 35, 23, 19, 27, 34, 30, 22, 16]


def main():
    q = pfvo.make_qtables(5)[0][0]
    px = np.array(KAT_A_PX, np.uint8)
    c = pfvo.encode_subblock(px, q)
    np.savez_compressed(os.path.join(HERE, "kat.npz"), kat_a_px=px, kat_a_coeff=c, kat_a_decoded=pfvo.decode_subblock(c, q))
    # --- stream fixture: 96x64, quality 3, I every 4, one drop frame
    sv = SynthVideo(96, 64, 77)
    enc = pfvo.Encoder(96, 64, 24, 3, nthreads=2)
    n = 10
    seam = []
    for t in range(n):
        y, u, v = sv.frame(t)
        if t % 4 == 0:
            enc.encode_iframe(y, u, v)
            seam.append((1, enc.last_headers(), enc.last_coeffs()))
        elif t == 5:
            enc.encode_dropframe()
        else:
            enc.encode_pframe(y, u, v)
            seam.append((2, enc.last_headers(), enc.last_coeffs()))
    enc.finish()
    data = enc.bytes()
    dec = pfvo.Decoder(data)
    sums, frames = [], []
    while True:
        more, fr = dec.advance_frame()
        if fr is not None:
            sums.append(hashlib.sha256(b"".join(p.tobytes() for p in fr)).hexdigest())
            frames.append(np.concatenate([p.ravel() for p in fr]))
        if not more:
            break
    np.savez_compressed(os.path.join(HERE, "stream_96x64_q3.npz"), stream=np.frombuffer(data, np.uint8), nframes=n,
                        sha256=np.array(sums), kinds=np.array([s[0] for s in seam], np.uint8),
                        headers=np.stack([s[1] for s in seam]), coeffs=np.stack([s[2] for s in seam]),
                        frames=np.stack(frames))
    print("stream:", len(data), "bytes,", len(sums), "frames")


if __name__ == "__main__":
    main()
