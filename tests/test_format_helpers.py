"""Colour / format steps next to the path: RGB <-> YUV 4:2:0 as the reference's test helpers do it
(load_frame / save_frame, src/lib.rs:337-395) with VideoPlane::reduce / double (src/common.rs:523-556).
CPU tier: the C oracle against the numpy-float32 restatement.  GPU tier: the device kernels against the oracle."""
import importlib.util
import os

import numpy as np
import pytest

import pfvo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location("pfv_ref", os.path.join(ROOT, "oracle", "pfv_ref.py"))
ref = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(ref)


def images(rng, w, h):
    yield rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    # saturation corners: pure primaries and their complements drive Y/U/V and R/G/B to the clamps
    img = np.zeros((h, w, 3), np.uint8)
    pal = np.array([[0, 0, 0], [255, 255, 255], [255, 0, 0], [0, 255, 0], [0, 0, 255], [255, 255, 0], [0, 255, 255], [255, 0, 255]], np.uint8)
    img[...] = pal[rng.integers(0, 8, (h, w))]
    yield img


@pytest.mark.parametrize("size", [(16, 16), (50, 38), (322, 242)])
def test_oracle_equals_float32_restatement(size):
    w, h = size
    rng = np.random.default_rng(w * 1000 + h)
    for rgb in images(rng, w, h):
        y, u, v = pfvo.rgb_to_yuv420(rgb)
        ry, ru, rv = ref.rgb_to_yuv420(rgb)
        assert np.array_equal(y, ry) and np.array_equal(u, ru) and np.array_equal(v, rv)
        assert np.array_equal(pfvo.yuv420_to_rgb(y, u, v), ref.yuv420_to_rgb(y, u, v))
    # every (y, u, v) byte combination of a slice that reaches all clamps
    yy = rng.integers(0, 256, (h, w), dtype=np.uint8)
    uu = rng.choice(np.array([0, 1, 127, 128, 129, 254, 255], np.uint8), (h // 2, w // 2))
    vv = rng.choice(np.array([0, 1, 127, 128, 129, 254, 255], np.uint8), (h // 2, w // 2))
    assert np.array_equal(pfvo.yuv420_to_rgb(yy, uu, vv), ref.yuv420_to_rgb(yy, uu, vv))


def test_known_values():
    # JPEG YCbCr constants: pure red -> Y 76, Cb 84, Cr 255 (0.5*255+128 = 255.5 -> 255).  Grey: `as u8` truncates, so
    # the f32 sum 127.99999 for Cr becomes 127 and the grey comes back as (126, 128, 128): the reference's behaviour
    rgb = np.zeros((2, 2, 3), np.uint8)
    rgb[...] = (255, 0, 0)
    y, u, v = pfvo.rgb_to_yuv420(rgb)
    assert y[0, 0] == 76 and u[0, 0] == 84 and v[0, 0] == 255
    g = np.full((2, 2, 3), 128, np.uint8)
    y, u, v = pfvo.rgb_to_yuv420(g)
    assert (y == 128).all() and u[0, 0] == 128 and v[0, 0] == 127
    assert (pfvo.yuv420_to_rgb(y, u, v) == np.array([126, 128, 128], np.uint8)).all()


@pytest.mark.gpu
@pytest.mark.parametrize("size", [(16, 16), (24, 10), (72, 50), (50, 38), (322, 242), (1920, 1080)])   # (widths that are / are not multiples of 8; partly filled last warp)
def test_gpu_rgb_paths_match_oracle(size):
    from pretty_fast_video_b200 import PFV_FRAME_I, PFV_FRAME_P, Engine, make_qtables
    from pretty_fast_video_b200.engine import EncodeJob
    w, h = size
    rng = np.random.default_rng(7 * w + h)
    qt, px_err = make_qtables(4)
    og = pfvo.geometry_for(w, h)
    imgs = list(images(rng, w, h))
    with Engine(w, h, qt, nslots=4, max_jobs=2) as e:
        # (1) encoder fed with RGB: coefficients, headers and reconstruction equal the oracle encoder fed with the
        #     oracle's YUV conversion of the same picture
        prev = pfvo.frame_init(og)
        y0, u0, v0 = pfvo.rgb_to_yuv420(imgs[0])
        want_c0 = pfvo.encode_iframe_coeffs(og, qt, y0, u0, v0, prev)
        rec0 = prev.copy()
        y1, u1, v1 = pfvo.rgb_to_yuv420(imgs[1])
        want_h1, want_c1 = pfvo.encode_pframe_coeffs(og, qt, px_err, y1, u1, v1, prev)
        rec1 = prev.copy()
        c0 = np.zeros(og.nb * 256, np.int16)
        c1 = np.zeros(og.nb * 256, np.int16)
        h1 = np.zeros((og.nb, 4), np.uint8)
        e.encode_submit([EncodeJob(PFV_FRAME_I, 0, None, c0, rgb=imgs[0])])
        e.encode_submit([EncodeJob(PFV_FRAME_P, 1, None, c1, ref_slot=0, px_err=px_err, hdr_out=h1, rgb=imgs[1])])
        e.sync()
        assert np.array_equal(c0, want_c0) and np.array_equal(h1, want_h1)
        coded = want_h1[:, 2] != 0
        assert np.array_equal(c1.reshape(-1, 256)[coded], want_c1.reshape(-1, 256)[coded])
        assert np.array_equal(e.slot_read(0), rec0) and np.array_equal(e.slot_read(1), rec1)
        # (2) a decoded slot read back as RGB equals save_frame of the oracle's crop of the same frame
        for slot, rec in ((0, rec0), (1, rec1)):
            y, u, v = pfvo.crop_frame(og, rec)
            assert np.array_equal(e.slot_read_rgb(slot), pfvo.yuv420_to_rgb(y, u, v))
        # (3) arbitrary planes (all clamps) through a slot
        fr = rng.integers(0, 256, rec0.size).astype(np.uint8)
        e.slot_write(2, fr)
        y, u, v = pfvo.crop_frame(og, fr)
        assert np.array_equal(e.slot_read_rgb(2), pfvo.yuv420_to_rgb(y, u, v))
        # (4) the batched entry point (pfv_slots_convert_rgb, one launch for several slots) against the oracle, slot by slot
        nbytes = w * h * 3
        stride = (nbytes + 255) & ~255
        torch = pytest.importorskip("torch")                         # (only to own a piece of device memory)
        buf = torch.empty(3 * stride, dtype=torch.uint8, device="cuda")
        e.slots_convert_rgb([2, 0, 1], buf.data_ptr(), stride)
        e.sync()
        got = buf.cpu().numpy()
        for i, (slot, rec) in enumerate(((2, fr), (0, rec0), (1, rec1))):
            y, u, v = pfvo.crop_frame(og, rec)
            assert np.array_equal(got[i * stride:i * stride + nbytes].reshape(h, w, 3), pfvo.yuv420_to_rgb(y, u, v)), f"batched picture {i}"
