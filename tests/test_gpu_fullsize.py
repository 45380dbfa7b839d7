"""The shapes bench.py times, compared with the oracle frame by frame (the timed launch shapes themselves
were only covered by small-size tests): 64 x 1080p key frames in ONE submit (grid.y = 64), 15-frame 1080p GOPs on 8
lanes, a 15-frame 3840x2160 GOP - every frame of every lane, not just the last framebuffer."""
import os

import numpy as np
import pytest

import pfvo
from pretty_fast_video_b200 import PFV_FRAME_I, PFV_FRAME_P, Engine, make_qtables
from pretty_fast_video_b200.engine import DecodeJob, EncodeJob
from pretty_fast_video_b200.synth import SynthVideo

pytestmark = pytest.mark.gpu
NT = min(16, os.cpu_count() or 1)


def test_64_key_frames_1080p_in_one_submit():
    w, h, n = 1920, 1080, 64
    qt, _ = make_qtables(5)
    og = pfvo.geometry_for(w, h)
    sv = SynthVideo(w, h, 0x50465601)
    coeffs, want = [], []
    for i in range(n):
        y, u, v = sv.frame(i)
        prev = pfvo.frame_init(og)
        coeffs.append(pfvo.encode_iframe_coeffs(og, qt, y, u, v, prev, NT))
        want.append(prev)                                   # the oracle's closed-loop reconstruction = its decode
    with Engine(w, h, qt, nslots=n, max_jobs=n) as e:
        e.decode_submit([DecodeJob(PFV_FRAME_I, i, coeffs[i], (0, 1, 1)) for i in range(n)])
        e.sync()
        for i in range(n):
            assert np.array_equal(e.slot_read(i), want[i]), f"job {i}"
        # the same batch flagged dense (the kernel that walks several tiles per warp with cp.async staging)
        for i in range(n):
            e.slot_reset(i)
        e.decode_submit([DecodeJob(PFV_FRAME_I, i, coeffs[i], (0, 1, 1), dense_hint=True) for i in range(n)])
        e.sync()
        for i in range(n):
            assert np.array_equal(e.slot_read(i), want[i]), f"dense-flagged job {i}"
        # the same batch through the encoder: 64 jobs per launch, coefficients and reconstruction
        outs = [np.zeros(og.nb * 256, np.int16) for _ in range(n)]
        e.encode_submit([EncodeJob(PFV_FRAME_I, i, sv.frame(i), outs[i]) for i in range(n)])
        e.sync()
        for i in range(n):
            assert np.array_equal(outs[i], coeffs[i]), f"encode job {i}"
            assert np.array_equal(e.slot_read(i), want[i]), f"encode recon {i}"


def _gop_lanes(w, h, lanes, gop, quality, seed):
    """Per lane: the oracle Encoder's seam data and reconstructed (padded) frame after every frame of one GOP."""
    out = []
    for lane in range(lanes):
        sv = SynthVideo(w, h, seed + lane)
        enc = pfvo.Encoder(w, h, 30, quality, nthreads=NT)
        frames = []
        for t in range(gop):
            y, u, v = sv.frame(t)
            if t == 0:
                enc.encode_iframe(y, u, v)
                frames.append((PFV_FRAME_I, None, enc.last_coeffs().copy(), (0, 1, 1), enc.prev_frame().copy(), (y, u, v)))
            else:
                enc.encode_pframe(y, u, v)
                frames.append((PFV_FRAME_P, enc.last_headers().copy(), enc.last_coeffs().copy(), (2, 3, 3), enc.prev_frame().copy(), (y, u, v)))
        enc.close()
        out.append(frames)
    return out


def _check_gops(w, h, lanes, gop, seed):
    qt, px_err = make_qtables(5)
    og = pfvo.geometry_for(w, h)
    data = _gop_lanes(w, h, lanes, gop, 5, seed)
    with Engine(w, h, qt, nslots=2 * lanes, max_jobs=lanes) as e:
        # decode: frame k of every lane per submit, the launch shape of the P-stream workloads
        cur = [2 * l for l in range(lanes)]
        for k in range(gop):
            jobs = []
            for l in range(lanes):
                kind, hdr, coeff, qidx, _, _ = data[l][k]
                jobs.append(DecodeJob(kind, cur[l] ^ 1, coeff, qidx, ref_slot=cur[l], hdr=hdr))
                cur[l] ^= 1
            e.decode_submit(jobs)
            e.sync()
            for l in range(lanes):
                assert np.array_equal(e.slot_read(cur[l]), data[l][k][4]), f"decode lane {l} frame {k}"
        # encode: the same GOPs (full block search), headers + coefficients + reconstruction of every frame
        for l in range(lanes):
            e.slot_reset(2 * l); e.slot_reset(2 * l + 1)
        cur = [2 * l for l in range(lanes)]
        for k in range(gop):
            jobs, hs, cs = [], [], []
            for l in range(lanes):
                kind = data[l][k][0]
                c = np.zeros(og.nb * 256, np.int16)
                hd = np.zeros((og.nb, 4), np.uint8)
                jobs.append(EncodeJob(kind, cur[l] ^ 1, data[l][k][5], c, ref_slot=cur[l], px_err=px_err, hdr_out=hd))
                cur[l] ^= 1
                hs.append(hd); cs.append(c)
            e.encode_submit(jobs)
            e.sync()
            for l in range(lanes):
                kind, whdr, wc, _, wrec, _ = data[l][k]
                if kind == PFV_FRAME_I:
                    assert np.array_equal(cs[l], wc), f"encode lane {l} frame {k} coefficients"
                else:
                    assert np.array_equal(hs[l], whdr), f"encode lane {l} frame {k} headers"
                    coded = whdr[:, 2] != 0
                    assert np.array_equal(cs[l].reshape(-1, 256)[coded], wc.reshape(-1, 256)[coded]), f"encode lane {l} frame {k} coefficients"
                assert np.array_equal(e.slot_read(cur[l]), wrec), f"encode lane {l} frame {k} reconstruction"


def test_full_gops_1080p_8_lanes_every_frame():
    _check_gops(1920, 1080, 8, 15, 0x50465602)


def test_full_gop_4k_every_frame():
    _check_gops(3840, 2160, 2, 15, 0x50465603)
