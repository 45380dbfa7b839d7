"""CPU: the C-ABI library loads, exports every symbol include/pfv_b200.h declares, its host-only entry points
agree with the oracle, and the compute entry points FAIL LOUDLY without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import pfvo
from pretty_fast_video_b200 import _native as N
from pretty_fast_video_b200 import Engine, PfvError, geometry_for, make_qtables

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "pfv_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pfv_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound():
    names = header_functions()
    assert len(names) >= 20
    lib = N.lib()
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/pfv_b200.h but not exported"
        assert n in N.SYMBOLS, f"{n} has no ctypes binding"
    assert lib.pfv_abi_version() == 1


def test_struct_layouts_match_the_header():
    assert C.sizeof(N.MbHdr) == 4
    assert C.sizeof(N.Geometry) == 48
    assert C.sizeof(N.DecodeJob) == 64 and N.DecodeJob.hdr.offset == 24 and N.DecodeJob.out_v.offset == 56
    assert C.sizeof(N.EncodeJob) == 64 and N.EncodeJob.src_y.offset == 24 and N.EncodeJob.coeff_out.offset == 56


@pytest.mark.parametrize("size", [(16, 16), (50, 38), (512, 384), (1920, 1080), (3840, 2160), (65534, 2)])
def test_geometry_matches_oracle(size):
    g, o = geometry_for(*size), pfvo.geometry_for(*size)
    for f in ("width", "height", "cwidth", "cheight", "pw", "ph", "cpw", "cph", "nb_y", "nb_c", "nb"):
        assert getattr(g, f) == getattr(o, f), f
    assert g.frame_bytes == g.pw * g.ph + 2 * g.cpw * g.cph


def test_known_geometries():
    g = geometry_for(1920, 1080)
    assert (g.pw, g.ph, g.cpw, g.cph, g.nb_y, g.nb_c, g.nb) == (1920, 1088, 960, 544, 8160, 2040, 12240)
    g = geometry_for(3840, 2160)
    assert g.nb == 48720
    assert geometry_for(512, 384).nb == 1152


def test_qtables_match_oracle_for_every_quality():
    for q in range(11):
        qt, px = make_qtables(q)
        oqt, opx = pfvo.make_qtables(q)
        assert np.array_equal(qt, oqt) and px == opx
    with pytest.raises(PfvError):
        make_qtables(11)                                    # src/enc.rs:38 assert
    with pytest.raises(PfvError):
        make_qtables(-1)


def test_no_gpu_means_loud_failure_not_fallback(has_gpu):
    if has_gpu:
        pytest.skip("GPU present")
    qt, _ = make_qtables(5)
    with pytest.raises(PfvError) as ei:
        Engine(64, 64, qt)
    assert ei.value.code == N.PFV_ERR_NO_DEVICE


def test_product_does_not_reach_into_the_oracle():
    """The oracle is test infrastructure: nothing under the package, include/ or bench.py's GPU legs may use it."""
    pkg = os.path.join(ROOT, "pretty_fast_video_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "pfv_oracle" not in text and "pfvo" not in text and "oracle/" not in text, os.path.join(dirpath, f)


def test_cpp_example_builds_and_links_against_the_abi():
    """examples/decode_speed.cpp (the reference's test_decode_speed_2, src/lib.rs:310-335, on the C ABI) compiles and
    links; without arguments it prints its usage (no GPU needed for that)."""
    import subprocess
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "examples")], stdout=subprocess.DEVNULL)
    r = subprocess.run([os.path.join(ROOT, "examples", "decode_speed")], capture_output=True, text=True)
    assert r.returncode == 2 and "usage" in r.stderr


def test_cubin_is_sm_100a_and_the_staging_paths_are_tma():
    """The shipped library carries sm_100a SASS only, and the kernels the design describes as TMA-staged really are:
    UTMALDG = cp.async.bulk.tensor (encode-P search window, decode-P predictor window), UBLKCP = cp.async.bulk (decode-I tiles),
    SYNCS = mbarrier completion (B200_PROFILING.md's mnemonics)."""
    import shutil
    import subprocess
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    lib = os.path.join(ROOT, "pretty_fast_video_b200", "libpfv_b200.so")
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    archs = {l.split("=")[1].strip() for l in sass.splitlines() if l.strip().startswith("arch =")}
    assert archs == {"sm_100a"}
    body, name = {}, None
    for l in sass.splitlines():
        if "Function :" in l:
            name = l.split("Function :")[1].strip()
            body[name] = []
        elif name:
            body[name].append(l)
    def has(kernel, mnemonic):
        hits = [k for k in body if kernel in k]
        assert hits, kernel
        return all(any(mnemonic in x for x in body[k]) for k in hits)
    assert has("encode_p_kernel", "UTMALDG") and has("encode_p_kernel", "SYNCS")
    assert has("encode_p2_kernel", "UTMALDG") and has("encode_p2_kernel", "SYNCS")          # the default encode-P kernel
    assert has("decode_p_fused_kernel", "UTMALDG") and has("decode_p_fused_kernel", "UBLKCP")
    assert has("mc_copy4_kernel", "UTMALDG")
    assert has("decode_i_stream_kernel", "UBLKCP") and has("decode_i_stream_kernel", "SYNCS")
    for k in ("tok_scan_kernel", "tok_emit_kernel", "tok_store_kernel", "expand_tokens_kernel", "residual_sb2_kernel",
              "encode_i_kernel", "rgb_to_yuv420_kernel", "yuv420_to_rgb_batch_kernel", "yuv420_to_rgb_batch8_kernel",
              "encode_p2_kernel", "decode_p_fused_kernel", "encode_i_persist_kernel"):
        assert any(k in n for n in body), k
