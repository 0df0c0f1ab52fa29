"""CPU-only tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/vsb200.h declares, and
its HOST-side calibration logic (no device work) matches the oracle and the reference golden vectors.  No compute entry
point is exercised here -- those are the -m gpu tests."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "vsb200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vsb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(vsb):
    lib = vsb.lib()
    names = _declared()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"libvsb200.so does not export {missing}"
    assert sorted(set(vsb.SYMBOLS)) == names, set(vsb.SYMBOLS) ^ set(names)


def test_cpp_adapter_header_compiles():
    """include/vsb200.hpp (the C++ adapter with the reference's class/function names) is self-contained C++11."""
    import shutil
    import subprocess
    import tempfile
    gxx = shutil.which("g++", path="/usr/bin") or shutil.which("g++")
    if not gxx or not os.path.exists(os.path.join(ROOT, "include", "vsb200.hpp")):
        pytest.skip("no g++ or adapter header")
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "t.cpp")
        open(src, "w").write('#include "vsb200.hpp"\nint main() { return 0; }\n')
        subprocess.check_call([gxx, "-std=c++11", "-fsyntax-only", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), src])


def test_c_header_is_plain_c_and_links(vsb):
    """include/vsb200.h is a C ABI: a C99 translation unit includes it, links against libvsb200.so and calls a host-only entry."""
    import shutil
    import subprocess
    import tempfile
    gcc = shutil.which("gcc", path="/usr/bin") or shutil.which("gcc")
    if not gcc:
        pytest.skip("no gcc")
    libdir = os.path.dirname(vsb.LIB_PATH)
    with tempfile.TemporaryDirectory() as d:
        src, exe = os.path.join(d, "t.c"), os.path.join(d, "t")
        open(src, "w").write('#include "vsb200.h"\nint main(void) { return vsb_consumer_image_height(3839, 627, 4096, 2048, 1) == 669 ? 0 : 1; }\n')
        subprocess.check_call([gcc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), src, "-L", libdir, "-lvsb200",
                               "-Wl,-rpath," + libdir, "-o", exe])
        assert subprocess.call([exe]) == 0


def test_no_device_fails_loudly(vsb):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    assert vsb.lib().vsb_device_count() == 0
    with pytest.raises(vsb.VsbError) as e:
        vsb.Stitcher(2, 3, True, 1)
    assert "vsb error -2" in str(e.value) or "no CUDA device" in str(e.value)


def test_bad_arguments_are_rejected_without_a_device(vsb):
    lib = vsb.lib()
    roi = (C.c_int * 4)()
    K = (C.c_float * 9)(*([1, 0, 0, 0, 1, 0, 0, 0, 1]))
    assert lib.vsb_warp_roi(7, C.c_float(100.0), K, K, 10, 10, roi) == -1
    assert b"warp_roi" in lib.vsb_last_error()
    assert lib.vsb_warp_roi(0, C.c_float(-1.0), K, K, 10, 10, roi) == -1
    assert lib.vsb_rig_camera(0, 0, 10, 10, C.c_double(90.0), K, K) == -1
    assert lib.vsb_create(None, None) == -1
    assert lib.vsb_destroy(None) == 0


@pytest.mark.parametrize("rig", [(6, 1920, 1080, 3840), (4, 320, 240, 1024), (12, 3840, 2160, 15360), (5, 640, 480, 2000)])
def test_host_geometry_matches_oracle_and_reference(vsb, og, rig):
    n, sw, sh, pano = rig
    gold = np.load(os.path.join(ROOT, "tests", "golden", "reference_cpu.npz"))["warp_roi"]
    scale = np.float32(pano / (2.0 * 3.1415926535897932384626))
    for i in range(n):
        K, R = vsb.rig_camera(n, i, sw, sh)
        Ko, Ro = og.rig_camera(n, i, sw, sh)
        assert np.array_equal(np.float32(K), Ko.reshape(9)) and np.array_equal(np.float32(R), Ro.reshape(9))
        for proj in (0, 1):
            roi = vsb.warp_roi(proj, float(scale), K, R, sw, sh)
            assert roi == og.warp_roi(proj, scale, Ko, Ro, sw, sh)
            row = gold[(gold[:, 0] == n) & (gold[:, 3] == pano) & (gold[:, 4] == proj) & (gold[:, 5] == i)]
            assert len(row) == 1 and roi == tuple(int(v) for v in row[0, 6:])   # = the reference's warpRoi


def test_compose_scale_sizes_and_cameras(vsb, og):
    """compose_scale != 1 (A/calibration.cpp:137-205): the sizes the reference derives -- cvRound for the frame and the blender, (int)
    for the maps, no resize within 0.1 of 1 -- and the scaled cameras, host side (no device needed), against the oracle's restatement."""
    assert vsb.compose_size(1920, 1080, 1.0) == ((1920, 1080), (1920, 1080), False)
    assert vsb.compose_size(1920, 1080, 0.5) == ((960, 540), (960, 540), True)
    # COMPOSE_MEGAPIX = 1.4 on 1080p frames, the reference's default (A/defs.h:53): cvRound and (int) disagree in the width
    cs = min(1.0, (1.4e6 / (1920 * 1080)) ** 0.5)
    assert vsb.compose_size(1920, 1080, cs) == ((1578, 887), (1577, 887), True)
    assert vsb.compose_size(1920, 1080, 0.95) == ((1920, 1080), (1824, 1026), False)    # cameras scaled, frames not (A/timed.cpp:75)
    assert vsb.compose_size(101, 51, 0.5) == ((50, 26), (50, 25), True)                   # cvRound is round-half-even: 50.5 -> 50, 25.5 -> 26
    lib = vsb.lib()
    two = (C.c_int * 2)()
    assert lib.vsb_compose_size(0, 10, C.c_double(0.5), two, two, None) == -1
    assert lib.vsb_compose_size(10, 10, C.c_double(0.0), two, two, None) == -1
    assert lib.vsb_compose_size(10, 10, C.c_double(0.5), None, two, None) == -1
    assert lib.vsb_set_compose_scale(None, C.c_double(0.5), 10, 10) == -1
    for (n, sw, sh, c) in ((6, 1920, 1080, cs), (4, 320, 240, 0.75), (12, 3840, 2160, 0.5)):
        for i in range(n):
            K, R = vsb.rig_camera_scaled(n, i, sw, sh, 90.0, c)
            Ko, Ro = og.rig_camera_scaled(n, i, sw, sh, 90.0, c)
            assert np.array_equal(np.float32(K), Ko.reshape(9)) and np.array_equal(np.float32(R), Ro.reshape(9))
            K1, _ = vsb.rig_camera(n, i, sw, sh)
            assert abs(K[0] - K1[0] * c) <= 1e-3 and abs(K[2] - K1[2] * c) <= 1e-3 and K[8] == 1.0


@pytest.mark.parametrize("rig,views", [((6, 1920, 1080, 3840), 7), ((6, 3840, 2160, 7680), 7), ((12, 3840, 2160, 15360), 15), ((2, 1280, 720, 4021), 3)])
def test_split_plan_of_the_baseline_configs(vsb, og, rig, views):
    """vsb_split_plan (host only) on BASELINE.json's configs: which cameras wrap around +-pi, where they are cut, and that afterwards no
    view is panorama-wide (the point of the split: config 4 has three panorama-wide views, 15360 columns each, without it)."""
    n, sw, sh, pano = rig
    plan = vsb.split_plan(0, pano, n, sw, sh, 5)
    assert len(plan) == views and sorted(set(c for c, _, _ in plan)) == list(range(n))
    scale = np.float32(pano / (2.0 * 3.1415926535897932384626))
    rois = [og.warp_roi(0, scale, *og.rig_camera(n, i, sw, sh), sw, sh) for i in range(n)]
    W = max(r[0] + r[2] for r in rois) - min(r[0] for r in rois)
    whole = [c for c, x0, w in plan if (x0, w) == (0, rois[c][2])]
    parts = [(c, x0, w) for c, x0, w in plan if (x0, w) != (0, rois[c][2])]
    assert len(parts) == 2 * (views - n) and all(rois[c][2] >= W - 1 for c, _, _ in parts) and all(rois[c][2] < W - 1 for c in whole)
    for c, x0, w in parts:
        assert x0 % 32 == 0 and x0 + w <= rois[c][2] and w < 0.3 * W + 240, (c, x0, w)
    for c in set(c for c, _, _ in parts):
        (_, a0, aw), (_, b0, bw) = [p for p in parts if p[0] == c]
        assert a0 == 0 and b0 + bw == rois[c][2] and b0 - aw > 4 * (3 * 32 + 8), "the two windows are the two ends of the image, far apart"
    lib = vsb.lib()
    nv = C.c_int()
    assert lib.vsb_split_plan(0, pano, 0, sw, sh, C.c_double(90.0), 5, C.byref(nv), None, None, None) == -1
    assert lib.vsb_split_plan(0, pano, n, sw, sh, C.c_double(90.0), 9, C.byref(nv), None, None, None) == -1
    assert lib.vsb_calibrate_rig_split(None, 0, pano, n, sw, sh, C.c_double(90.0), None, 0) == -1


def test_reference_scales_and_work_scale_cameras(vsb, og):
    """vsb_ref_scales / vsb_rig_camera_work (host only): stitch_calib's work_scale and compose_scale from its MEGAPIX constants, and the
    cameras of calibrateCameras at a work scale, against the oracle's restatement; the special cases reduce to the existing functions."""
    from oracle import pipeline as op
    assert vsb.ref_scales(1920, 1080) == op.ref_scales(1920, 1080) == ((0.6e6 / (1920 * 1080)) ** 0.5, (1.4e6 / (1920 * 1080)) ** 0.5)
    assert vsb.ref_scales(640, 480, -1.0, -1.0) == (1.0, 1.0) and vsb.ref_scales(320, 240) == (1.0, 1.0)          # small frames: min(1, .)
    assert vsb.lib().vsb_ref_scales(0, 10, C.c_double(0.6), C.c_double(1.4), None, None) == -1
    ws, cs = vsb.ref_scales(1920, 1080)
    for (n, sw, sh) in ((6, 1920, 1080), (4, 320, 240), (12, 3840, 2160)):
        for i in range(n):
            for (w_, a_) in ((1.0, 1.0), (ws, 1.0), (ws, cs / ws), (1.0, 0.75)):
                K, R = vsb.rig_camera_work(n, i, sw, sh, 90.0, w_, a_)
                Ko, Ro = og.rig_camera_work(n, i, sw, sh, 90.0, w_, a_)
                assert np.array_equal(np.float32(K), Ko.reshape(9)) and np.array_equal(np.float32(R), Ro.reshape(9))
            assert vsb.rig_camera_work(n, i, sw, sh, 90.0, 1.0, 1.0) == vsb.rig_camera(n, i, sw, sh)
            assert vsb.rig_camera_work(n, i, sw, sh, 90.0, 1.0, 0.75) == vsb.rig_camera_scaled(n, i, sw, sh, 90.0, 0.75)
    # the reference's default panorama of 6 x 1080p (WORK 0.6, COMPOSE 1.4): sphere radius = the work-scale focal length x compose_work_aspect
    rig = op.OracleRig(6, 1920, 1080, 0, num_bands=5, compose_scale=cs, work_scale=ws)
    assert abs(float(rig.scale) - 960.0 * cs) < 0.05 and (rig.comp_w, rig.comp_h) == (1578, 887) and rig.map_src == (1577, 887)
    assert rig.roi_final[2] in range(4950, 4960) and rig.num_bands == 5


def test_host_voronoi_matches_reference(vsb, og):
    from tests.golden import make_golden as G
    gold = np.load(os.path.join(ROOT, "tests", "golden", "reference_cpu.npz"))
    n, sw, sh, pano, proj = G.SEAM_RIGS[0]
    masks, corners, sizes = G.seam_inputs(og, n, sw, sh, pano, proj)
    s = np.ascontiguousarray(np.array(sizes, np.int32).reshape(-1))
    c = np.ascontiguousarray(np.array(corners, np.int32).reshape(-1))
    ptrs = (C.c_void_p * n)(*[m.ctypes.data for m in masks])
    vsb.check(vsb.lib().vsb_voronoi_seams(n, s.ctypes.data_as(C.POINTER(C.c_int)), c.ctypes.data_as(C.POINTER(C.c_int)), ptrs))
    for i, m in enumerate(masks):
        assert np.array_equal(np.packbits(m > 0, axis=1), gold[f"voronoi_{n}_{pano}_{proj}_{i}"])


def test_consumer_geometry_matches_oracle(vsb, og):
    """Host-only part of the consumer epilogue: the image height of 360_stitcher/timed.cpp:254-270 (no device needed)."""
    lib = vsb.lib()
    for (w, h, ow, oh) in ((3839, 627, 4096, 2048), (7678, 1253, 4096, 2048), (383, 63, 512, 256), (100, 400, 64, 64), (1, 1, 2, 2)):
        for keep in (0, 1):
            assert lib.vsb_consumer_image_height(w, h, ow, oh, keep) == og.consumer_image_height(w, h, ow, oh, bool(keep))


def test_unit_weight_normalisation_shortcut_is_exact():
    """k_blend / k_coarse replace trunc(acc / (1 + 1e-5f)) by acc - sign(acc) when the weight sum is exactly 1
    (vsb_blend_kernels.cuh normalize_s16): exhaustive over the 16-bit accumulator range, in fp32."""
    a = np.arange(-32768, 32768, dtype=np.int32)
    den = np.float32(1.0) + np.float32(1e-5)
    q = np.trunc(a.astype(np.float32) / den).astype(np.int32)
    assert np.array_equal(q, a - np.sign(a))


def test_product_never_touches_the_oracle():
    """The product path must not import, link or execute anything under oracle/ (nor fall back to the CPU)."""
    pkg = os.path.join(ROOT, "video-stitcher_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in text.replace("no oracle dependency", ""), f"{f} mentions the oracle"
    for f in ("vsb200.h", "vsb200.hpp"):
        p = os.path.join(ROOT, "include", f)
        if os.path.exists(p):
            assert "oracle/" not in open(p).read()
