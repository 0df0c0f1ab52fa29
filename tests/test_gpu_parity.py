"""-m gpu parity tests: the CUDA path (through the C ABI) against oracle-G on identical seeded inputs.
Bar: bit-exact (every stage is integer/byte work or fp32 with a pinned operation order)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _eq(a, b, what):
    a = np.asarray(a); b = np.asarray(b)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    if a.dtype.kind == "f":
        same = (a == b) | (np.isnan(a) & np.isnan(b))
    else:
        same = a == b
    assert same.all(), f"{what}: {np.count_nonzero(~same)} of {a.size} samples differ (max |d| = " \
                       f"{np.nanmax(np.abs(a.astype(np.float64) - b.astype(np.float64)))})"


# ------------------------------------------------------------------------------------------ B7 primitives
@pytest.mark.parametrize("shape,dshape", [((120, 160), (90, 130)), ((33, 47), (64, 64)), ((1080, 1920), (200, 333))])
def test_remap_linear_matches_oracle(cuda, og, vsb, shape, dshape):
    # recipe of sources/modules/cudawarping/test/test_remap.cpp:158-177 (random image, maps reaching outside the image)
    from tests.gpu_util import dev, host, stream
    rng = np.random.default_rng(1)
    src = rng.integers(0, 256, shape + (3,), dtype=np.uint8)
    xm = (rng.random(dshape) * (shape[1] + 10) - 5).astype(np.float32)
    ym = (rng.random(dshape) * (shape[0] + 10) - 5).astype(np.float32)
    xm[0, 0] = np.nan; ym[1, 1] = np.nan; xm[2, 2] = -1; ym[2, 2] = -1; xm[3, 3] = 1e12
    # every border class of the tap loader: exactly on / just inside / just outside each edge, infinities
    edge_x = [-1.0, -0.5, -1.0000001, 0.0, shape[1] - 2.0, shape[1] - 1.0, shape[1] - 0.5, float(shape[1]), np.inf, -np.inf, -1e12]
    edge_y = [-1.0, -0.25, 0.0, shape[0] - 1.0, shape[0] - 0.75, float(shape[0]), shape[0] - 2.0, 0.5, 3.0, 2.0, 1.0]
    for j, (ex, ey) in enumerate(zip(edge_x, edge_y)):
        xm[5, j] = ex; ym[5, j] = 7.25
        xm[6, j] = 9.5; ym[6, j] = ey
        xm[7, j] = ex; ym[7, j] = ey
    xm[8, :] = shape[1] - 1 - rng.random(dshape[1]).astype(np.float32) * 3   # last columns of the last rows (tail loads)
    ym[8, :] = shape[0] - 1 - rng.random(dshape[1]).astype(np.float32) * 2
    want = og.remap_linear_u8(src, xm, ym)
    d_src, d_x, d_y = dev(src), dev(xm), dev(ym)
    d_dst = cuda.zeros(dshape + (3,), dtype=cuda.uint8, device="cuda")
    vsb.check(vsb.lib().vsb_remap_linear_u8c3(vsb._vp(d_src.data_ptr()), shape[1], shape[0], vsb.C.c_size_t(shape[1] * 3),
                                              vsb._vp(d_x.data_ptr()), vsb._vp(d_y.data_ptr()), vsb.C.c_size_t(dshape[1] * 4),
                                              vsb._vp(d_dst.data_ptr()), dshape[1], dshape[0], vsb.C.c_size_t(dshape[1] * 3), vsb._vp(stream())))
    _eq(host(d_dst), want, "remap")


@pytest.mark.parametrize("shape", [(64, 96), (160, 224), (33, 47), (20, 37), (640, 1184)])
def test_pyramids_match_oracle(cuda, og, vsb, shape):
    # sizes of sources/modules/cudawarping/test/test_pyramids.cpp plus the blender's CV_16SC3 type it omits
    from tests.gpu_util import dev, host, stream
    rng = np.random.default_rng(2)
    a = rng.integers(-32768, 32768, shape + (3,)).astype(np.int16)
    h, w = shape
    d_a = dev(a)
    dh, dw = (h + 1) // 2, (w + 1) // 2
    d_dn = cuda.zeros((dh, dw, 3), dtype=cuda.int16, device="cuda")
    vsb.check(vsb.lib().vsb_pyr_down_s16c3(vsb._vp(d_a.data_ptr()), w, h, vsb.C.c_size_t(w * 6), vsb._vp(d_dn.data_ptr()), vsb.C.c_size_t(dw * 6), vsb._vp(stream())))
    _eq(host(d_dn), og.pyr_down_s16(a), "pyrDown s16c3")
    d_up = cuda.zeros((2 * h, 2 * w, 3), dtype=cuda.int16, device="cuda")
    vsb.check(vsb.lib().vsb_pyr_up_s16c3(vsb._vp(d_a.data_ptr()), w, h, vsb.C.c_size_t(w * 6), vsb._vp(d_up.data_ptr()), vsb.C.c_size_t(2 * w * 6), vsb._vp(stream())))
    _eq(host(d_up), og.pyr_up_s16(a), "pyrUp s16c3")
    f = rng.random(shape).astype(np.float32)
    d_f = dev(f)
    d_fd = cuda.zeros((dh, dw), dtype=cuda.float32, device="cuda")
    vsb.check(vsb.lib().vsb_pyr_down_f32(vsb._vp(d_f.data_ptr()), w, h, vsb.C.c_size_t(w * 4), vsb._vp(d_fd.data_ptr()), vsb.C.c_size_t(dw * 4), vsb._vp(stream())))
    _eq(host(d_fd), og.pyr_down_f32(f), "pyrDown f32")


def test_border_gain_resize_weighted_add_match_oracle(cuda, og, vsb):
    from tests.gpu_util import dev, host, stream
    rng = np.random.default_rng(3)
    L = vsb.lib(); vp = vsb._vp; sz = vsb.C.c_size_t
    img = rng.integers(0, 256, (50, 70, 3), dtype=np.uint8)
    d_img = dev(img)
    for (t, b, l, r) in [(0, 13, 127, 97), (5, 0, 0, 1), (60, 3, 80, 2)]:
        d = cuda.zeros((50 + t + b, 70 + l + r, 3), dtype=cuda.int16, device="cuda")
        vsb.check(L.vsb_border_reflect_u8c3_to_s16c3(vp(d_img.data_ptr()), 70, 50, sz(210), t, b, l, r, vp(d.data_ptr()), sz((70 + l + r) * 6), vp(stream())))
        _eq(host(d), og.border_reflect_u8c3_to_s16(img, t, b, l, r), "copyMakeBorder REFLECT")
    for g in (0.97, 1.0, 1.03, 1.7):
        d = dev(img)
        vsb.check(L.vsb_gain_u8(vp(d.data_ptr()), 210, 50, sz(210), vsb.C.c_float(g), vp(stream())))
        _eq(host(d), og.gain_u8(img, np.float32(g)), "gain")
    m = rng.random((10, 10)).astype(np.float32) * 100
    d = cuda.zeros((627, 961), dtype=cuda.float32, device="cuda")
    d_m = dev(m)
    vsb.check(L.vsb_custom_resize(vp(d_m.data_ptr()), 10, 10, sz(40), vp(d.data_ptr()), 961, 627, sz(961 * 4), vp(stream())))
    _eq(host(d), og.custom_resize(m, 961, 627), "custom_resize")
    # addSrcWeight / normalize: replay against the oracle blender arithmetic on one level
    src = rng.integers(-300, 300, (40, 64, 3)).astype(np.int16)
    w = rng.random((40, 64)).astype(np.float32)
    w[rng.random((40, 64)) < 0.3] = 0
    dst = rng.integers(-300, 300, (40, 64, 3)).astype(np.int16)
    dw = rng.random((40, 64)).astype(np.float32)
    d_dst, d_dw, d_s, d_w = dev(dst), dev(dw), dev(src), dev(w)
    vsb.check(L.vsb_add_src_weight_32f(vp(d_s.data_ptr()), sz(64 * 6), vp(d_w.data_ptr()), sz(256), vp(d_dst.data_ptr()), sz(64 * 6), vp(d_dw.data_ptr()), sz(256), 64, 40, vp(stream())))
    want = (dst.astype(np.int32) + np.trunc(src.astype(np.float32) * w[..., None]).astype(np.int32)).astype(np.int16)
    _eq(host(d_dst), want, "addSrcWeight32F")
    _eq(host(d_dw), dw + w, "addSrcWeight32F weights")
    vsb.check(L.vsb_normalize_32f(vp(d_dw.data_ptr()), sz(256), vp(d_dst.data_ptr()), sz(64 * 6), 64, 40, vp(stream())))
    # static_cast<short>(float) on the device is cvt.rzi.s16.f32: truncates and saturates
    want2 = np.clip(np.trunc(want.astype(np.float32) / ((dw + w) + np.float32(1e-5))[..., None]), -32768, 32767).astype(np.int16)
    _eq(host(d_dst), want2, "normalize32F")


@pytest.mark.parametrize("proj,n,sw,sh,pano", [(0, 6, 640, 360, 1280), (1, 5, 320, 240, 900), (0, 6, 1920, 1080, 3840)])
def test_build_maps_and_warp_match_oracle(cuda, og, vsb, proj, n, sw, sh, pano):
    """B3: {Spherical,Cylindrical}WarperGpu::buildMaps (S/src/warpers_cuda.cpp:210-277, S/src/cuda/build_warp_maps.cu:88-152) and
    ::warp (:279-298) on the device.  ROI exact; maps within 2e-3 px of oracle-G's (device sinf / cosf vs libm, the same bound the
    oracle's own maps keep against the reference's CPU projector, tests/test_oracle_pin.py); the warp itself is checked on the
    DEVICE maps, bit-exact (NEAREST / LINEAR x CONSTANT / REFLECT, CV_8UC1 / CV_8UC3: the modes of A/calibration.cpp:118,122,227)."""
    from tests.gpu_util import dev, host, stream
    C = vsb.C
    L = vsb.lib()
    scale = np.float32(pano / (2.0 * 3.1415926535897932384626))
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (sh, sw, 3), dtype=np.uint8)
    mask = np.full((sh, sw), 255, np.uint8)
    d_img, d_mask = dev(img), dev(mask)
    for i in (0, n // 2, n - 1):
        K, R = og.rig_camera(n, i, sw, sh)
        want_roi = og.warp_roi(proj, scale, K, R, sw, sh)
        assert vsb.warp_roi(proj, float(scale), K.reshape(9), R.reshape(9), sw, sh) == want_roi
        w, h = want_roi[2], want_roi[3]
        pitch = (w * 4 + 255) // 256 * 256
        d_x = cuda.full((h, pitch // 4), -7.0, dtype=cuda.float32, device="cuda")
        d_y = cuda.full((h, pitch // 4), -7.0, dtype=cuda.float32, device="cuda")
        roi = (C.c_int * 4)()
        vsb.check(L.vsb_build_maps(proj, C.c_float(scale), vsb._fp9(K.reshape(9)), vsb._fp9(R.reshape(9)), sw, sh, vsb._vp(d_x.data_ptr()),
                                   vsb._vp(d_y.data_ptr()), C.c_size_t(pitch), roi, vsb._vp(stream())))
        assert tuple(roi) == want_roi
        gx, gy = host(d_x), host(d_y)
        assert (gx[:, w:] == -7.0).all() and (gy[:, w:] == -7.0).all(), "row padding must stay untouched"
        gx, gy = np.ascontiguousarray(gx[:, :w]), np.ascontiguousarray(gy[:, :w])
        wx, wy = og.build_maps(proj, scale, K, R, *want_roi)
        behind = (wx == -1) & (wy == -1)
        assert np.array_equal(behind, (gx == -1) & (gy == -1)), "rays behind the camera map to (-1, -1)"
        inside = ~behind & (wx > -2) & (wx < sw + 1) & (wy > -2) & (wy < sh + 1)   # the part of the map that addresses the image
        assert inside.any()
        assert np.abs(gx - wx)[inside].max() <= 2e-3 and np.abs(gy - wy)[inside].max() <= 2e-3
        for (src, d_src, cn, interp, border) in ((img, d_img, 3, 1, 2), (mask, d_mask, 1, 0, 0), (img, d_img, 3, 1, 0), (img[..., 1].copy(), None, 1, 1, 2)):
            if d_src is None:
                d_src = dev(src)
            dp = w * cn + 3
            d_dst = cuda.full((h, dp), 9, dtype=cuda.uint8, device="cuda")
            vsb.check(L.vsb_warp(proj, C.c_float(scale), vsb._fp9(K.reshape(9)), vsb._fp9(R.reshape(9)), vsb._vp(d_src.data_ptr()), sw, sh,
                                 C.c_size_t(sw * cn), cn, interp, border, vsb._vp(d_dst.data_ptr()), C.c_size_t(dp), roi, vsb._vp(stream())))
            got = host(d_dst)
            assert tuple(roi) == want_roi and (got[:, w * cn:] == 9).all()
            got = got[:, :w * cn].reshape((h, w, 3) if cn == 3 else (h, w))
            _eq(got, og.remap_u8(src, gx, gy, interp, border), f"warp view {i} cn {cn} interp {interp} border {border}")


# ------------------------------------------------------------------------------------------ calibration on the device (8f row 4)
def test_device_voronoi_dilate_resize_gain_primitives(cuda, og, vsb):
    """VoronoiSeamFinder on device masks against the reference's golden seams (bit-exact), MORPH_DILATE / cuda::resize against
    oracle-G, GainCompensator::feed on device images against the reference's own gains (float64, bit for bit)."""
    import torch
    from tests.gpu_util import dev, host, stream
    from tests.golden import make_golden as G
    import os
    gold = np.load(os.path.join(os.path.dirname(G.__file__), "reference_cpu.npz"))
    C, L = vsb.C, vsb.lib()
    for (n, sw, sh, pano, proj) in G.SEAM_RIGS:
        masks, corners, sizes = G.seam_inputs(og, n, sw, sh, pano, proj)
        d_masks = [dev(m) for m in masks]
        ptrs = (C.c_void_p * n)(*[t.data_ptr() for t in d_masks])
        sz = (C.c_int * (2 * n))(*[int(v) for p in sizes for v in p])
        co = (C.c_int * (2 * n))(*[int(v) for p in corners for v in p])
        vsb.check(L.vsb_voronoi_seams_device(n, sz, co, ptrs, vsb._vp(stream())))
        for i, t in enumerate(d_masks):
            got = host(t)
            assert np.array_equal(np.packbits(got > 0, axis=1), gold[f"voronoi_{n}_{pano}_{proj}_{i}"]), f"rig {n}/{pano}/{proj} view {i}"
            assert set(np.unique(got)) <= {0, 255}
    rng = np.random.default_rng(9)
    sm = (rng.random((57, 101)) > 0.5).astype(np.uint8) * 255
    d_sm, d_dil = dev(sm), cuda.zeros((57, 101), dtype=cuda.uint8, device="cuda")
    vsb.check(L.vsb_dilate3x3_u8(vsb._vp(d_sm.data_ptr()), 101, 57, vsb._vp(d_dil.data_ptr()), vsb._vp(stream())))
    _eq(host(d_dil), og.dilate3x3_u8c1(sm), "dilate 3x3")
    d_big = cuda.zeros((627, 961), dtype=cuda.uint8, device="cuda")
    vsb.check(L.vsb_resize_linear_u8(vsb._vp(d_dil.data_ptr()), 101, 57, C.c_size_t(101), 1, vsb._vp(d_big.data_ptr()), 961, 627, C.c_size_t(961),
                                     C.c_double(0.0), C.c_double(0.0), vsb._vp(stream())))
    _eq(host(d_big), og.resize_linear_u8c1(host(d_dil), 961, 627), "cuda::resize CV_8UC1 (sizes)")
    img = rng.integers(0, 256, (270, 480, 3), dtype=np.uint8)
    ss = min(1.0, (0.01e6 / (480 * 270)) ** 0.5)
    dw, dh = int(np.rint(480 * ss)), int(np.rint(270 * ss))
    d_img, d_small = dev(img), cuda.zeros((dh, dw, 3), dtype=cuda.uint8, device="cuda")
    vsb.check(L.vsb_resize_linear_u8(vsb._vp(d_img.data_ptr()), 480, 270, C.c_size_t(480 * 3), 3, vsb._vp(d_small.data_ptr()), dw, dh, C.c_size_t(dw * 3),
                                     C.c_double(ss), C.c_double(ss), vsb._vp(stream())))
    _eq(host(d_small), og.cuda_resize_linear_u8(img, dw, dh, ss, ss), "cuda::resize CV_8UC3 (explicit scale)")
    imgs, masks, corners, sizes = G.gain_input()
    n = len(imgs)
    d_i, d_m = [dev(a) for a in imgs], [dev(a) for a in masks]
    ip = (C.c_void_p * n)(*[t.data_ptr() for t in d_i]); mp = (C.c_void_p * n)(*[t.data_ptr() for t in d_m])
    sz = (C.c_int * (2 * n))(*[int(v) for p in sizes for v in p]); co = (C.c_int * (2 * n))(*[int(v) for p in corners for v in p])
    g = (C.c_double * n)()
    vsb.check(L.vsb_gain_compensator_feed(n, ip, mp, sz, co, g, vsb._vp(stream())))
    assert np.array_equal(np.array(list(g)), gold["gain_compensator"]), "device gain estimation vs the reference's GainCompensator"


@pytest.mark.parametrize("case", ["small6_nolocal", "cfg2"])
def test_device_calibration_products_and_compose(cuda, og, case):
    """vsb_calibrate_rig_device: ROIs and blender geometry equal the oracle's; projection maps within 2e-3 px (device sinf / cosf);
    seam masks equal except where a map difference moves a boundary sample (< 0.2 % of the samples); and, with the DEVICE's own
    products injected into oracle-G, the composed panorama is bit-exact -- everything downstream of the maps is exact.  Gains
    estimated at run time (vsb_estimate_gains) agree with the oracle's pipeline on its libm maps to 1e-4."""
    import vsb200
    from oracle import pipeline as op
    from tests.gpu_util import GpuRig, dev, stream
    kw = dict(CASES[case])
    gains = vsb200.synth.gains(kw["n_views"])
    ref = op.OracleRig(gains=gains, **kw)
    grig = GpuRig(gains=gains, device_calibration=True, **kw)
    n = kw["n_views"]
    assert grig.roi_final == ref.roi_final and grig.roi_padded == ref.roi_padded and grig.num_bands == ref.num_bands
    xmaps, ymaps, masks = [], [], []
    for i in range(n):
        assert grig.geom[i] == ref.blender.view_geom(i) and grig.sizes[i] == tuple(ref.sizes[i]) and grig.corners[i] == tuple(ref.corners[i])
        gx, gy = grig.proj_map(i, 0), grig.proj_map(i, 1)
        inside = (ref.xmaps[i] > -2) & (ref.xmaps[i] < kw["src_w"] + 1) & (ref.ymaps[i] > -2) & (ref.ymaps[i] < kw["src_h"] + 1) & ~((ref.xmaps[i] == -1) & (ref.ymaps[i] == -1))
        assert np.abs(gx - ref.xmaps[i])[inside].max() <= 2e-3 and np.abs(gy - ref.ymaps[i])[inside].max() <= 2e-3
        g = grig.geom[i]
        w0 = grig.weight(i, 0)[g["top"]:g["top"] + grig.sizes[i][1], g["left"]:g["left"] + grig.sizes[i][0]]
        m = np.rint(w0 * 255).astype(np.uint8)
        assert np.count_nonzero(m != ref.masks[i]) <= 0.002 * m.size, f"seam mask view {i}"
        xmaps.append(gx); ymaps.append(gy); masks.append(m)
    mine = op.OracleRig.from_products(kw["src_w"], kw["src_h"], grig.corners, grig.sizes, xmaps, ymaps, masks, kw["num_bands"], kw["enable_local"], gains)
    if kw["enable_local"]:
        for i in range(n):
            mx, my = vsb200.synth.mesh(*grig.sizes[i])
            mine.set_mesh(i, mx, my); grig.set_mesh(i, mx, my)
    frames = [vsb200.synth.frame(i, 0, kw["src_w"], kw["src_h"]) for i in range(n)]
    _eq(grig.compose([frames])[0], mine.compose(frames)[0], "panorama from the device's calibration products")
    # run-time gain refresh: darken two cameras, estimate, install, and the estimate follows
    dim = [np.clip(f.astype(np.float32) * (0.8 if i in (1, 4 % n) else 1.0), 0, 255).astype(np.uint8) for i, f in enumerate(frames)]
    srcs = [dev(f) for f in dim]
    got = np.array(grig.st.estimate_gains([t.data_ptr() for t in srcs], kw["src_w"] * 3, apply=True, stream=stream()))
    want = ref.estimate_gains(dim)
    assert np.abs(got - want).max() <= 1e-4 * np.abs(want).max(), (got, want)
    assert got[1] > got[0], "the darkened camera gets the larger gain"
    mine.gains = [float(np.float32(v)) for v in got]
    _eq(grig.compose([dim])[0], mine.compose(dim)[0], "panorama after the run-time gain refresh")


# ------------------------------------------------------------------------------------------ whole path
CASES = {
    "small4": dict(n_views=4, src_w=320, src_h=240, pano_width=1024, num_bands=3, enable_local=True),
    "small6_nolocal": dict(n_views=6, src_w=320, src_h=180, pano_width=960, num_bands=5, enable_local=False),
    "cyl5": dict(n_views=5, src_w=256, src_h=192, pano_width=800, num_bands=4, enable_local=True, projection=1),
    "cfg2": dict(n_views=6, src_w=1920, src_h=1080, pano_width=3840, num_bands=5, enable_local=True),
    "bands2": dict(n_views=4, src_w=320, src_h=240, pano_width=1024, num_bands=2, enable_local=True),   # generic per-level path
    "bands6": dict(n_views=3, src_w=400, src_h=300, pano_width=1100, num_bands=6, enable_local=True),
    "bands7": dict(n_views=5, src_w=320, src_h=200, pano_width=1300, num_bands=7, enable_local=False),
    "wide2": dict(n_views=2, src_w=640, src_h=360, pano_width=2011, num_bands=5, enable_local=True),    # BASELINE config 1 shape (2 cams)
}


def _rigs(case, inject, max_batch=1):
    import vsb200
    from oracle import pipeline as op
    from tests.gpu_util import GpuRig
    kw = dict(CASES[case])
    gains = vsb200.synth.gains(kw["n_views"])
    orig = op.OracleRig(gains=gains, **kw)
    grig = GpuRig(gains=gains, max_batch=max_batch, oracle_rig=orig if inject else None, **kw)
    if kw["enable_local"]:
        for i in range(kw["n_views"]):
            mx, my = vsb200.synth.mesh(*orig.sizes[i])
            orig.set_mesh(i, mx, my)
            grig.set_mesh(i, mx, my)
    return orig, grig, kw


@pytest.mark.parametrize("case", ["small4", "small6_nolocal", "cyl5", "cfg2"])
def test_calibration_products_match_oracle(cuda, og, case):
    orig, grig, kw = _rigs(case, inject=False)
    assert grig.roi_final == orig.roi_final and grig.roi_padded == orig.roi_padded and grig.num_bands == orig.num_bands
    for i in range(kw["n_views"]):
        assert grig.geom[i] == orig.blender.view_geom(i), f"view {i} border geometry"
        assert grig.sizes[i] == tuple(orig.sizes[i]) and grig.corners[i] == tuple(orig.corners[i])
        for k in range(orig.num_bands + 1):
            _eq(grig.weight(i, k), orig.blender.view_weight(i, k), f"weight pyramid view {i} level {k}")
        # level-0 weight inside the un-bordered rect is mask/255: seam masks bit-exact
        g = grig.geom[i]
        w0 = grig.weight(i, 0)[g["top"]:g["top"] + orig.sizes[i][1], g["left"]:g["left"] + orig.sizes[i][0]]
        _eq(np.rint(w0 * 255).astype(np.uint8), orig.masks[i], f"seam mask view {i}")
        if kw["enable_local"]:
            _eq(grig.mesh_map(i, 0), orig.mesh_maps[i][0], f"x mesh map view {i}")
            _eq(grig.mesh_map(i, 1), orig.mesh_maps[i][1], f"y mesh map view {i}")


@pytest.mark.parametrize("case,inject", [("small4", True), ("small4", False), ("small6_nolocal", False), ("cyl5", False),
                                          ("cfg2", True), ("cfg2", False), ("bands2", False), ("bands6", False), ("bands7", False),
                                          ("wide2", False)])
def test_compose_matches_oracle(cuda, og, case, inject):
    import vsb200
    orig, grig, kw = _rigs(case, inject)
    frames = [vsb200.synth.frame(i, 0, kw["src_w"], kw["src_h"]) for i in range(kw["n_views"])]
    want, want_mask = orig.compose(frames)
    got = grig.compose([frames])[0]
    # intermediates first so a failure names the stage
    skipped = 0.0
    for i in range(kw["n_views"]):
        # remap #2 only computes the bordered tiles a later kernel reads (static table): compare there
        g = grig.geom[i]
        done0 = grig.g0_computed(i).astype(bool)
        skipped += 1.0 - done0.mean()
        crop = done0[g["top"]:g["top"] + grig.sizes[i][1], g["left"]:g["left"] + grig.sizes[i][0]]
        _eq(grig.warped(i)[crop], orig.warp_view(i, frames[i])[crop], f"warped view {i}")
    if case == "cfg2":
        assert skipped / kw["n_views"] > 0.05, "tile skipping should drop the parts of the views no band reads"
    for i in range(kw["n_views"]):
        # oracle src_level holds Laplacians after feed; rebuild the Gaussian pyramid from the warped view
        g = orig.blender.view_geom(i)
        gk = og.border_reflect_u8c3_to_s16(orig.warp_view(i, frames[i]), g["top"], g["bottom"], g["left"], g["right"])
        done0 = grig.g0_computed(i).astype(bool)
        fast = orig.num_bands >= 3  # fast path materialises levels 0 and 2 only, level 2 only in the tiles something reads
        for k in range(orig.num_bands + 1):
            if not fast:
                assert done0.all()
                _eq(grig.gauss_level(i, k), gk, f"gaussian level {k} view {i}")
            elif k == 0:
                _eq(grig.gauss_level(i, 0)[done0], gk[done0], f"gaussian level 0 view {i}")
            elif k == 2:
                done = grig.g2_computed(i).astype(bool)
                assert done.any()
                _eq(grig.gauss_level(i, 2)[done], gk[done], f"gaussian level 2 view {i}")
            gk = og.pyr_down_s16(gk)
    _eq(got, want, "composed panorama (CV_16SC3)")
    _eq((np.abs(got).sum(axis=2) > 0) | (want_mask > 0), want_mask > 0, "output mask support")
    assert grig.st.last_launch_count() == (7 if orig.num_bands >= 3 else 3 + orig.num_bands)  # K1 K2 down2 down_tail coarse blend_seam blend_int
    # B4/B5 per-view entry points give the same frame
    _eq(grig.feed_blend(frames), want, "feed + blend")


def test_full_size_config4_compose(cuda, og):
    """BASELINE.json configs[3] at full size: 12 x 3840x2160 -> 15360-wide spherical panorama, CPW on, 5 bands (maximum sizes:
    ~125 MPx of warped views per frame).  One frame, bit-exact against oracle-G."""
    import vsb200
    kw = dict(n_views=12, src_w=3840, src_h=2160, pano_width=15360, num_bands=5, enable_local=True)
    CASES["cfg4"] = kw
    orig, grig, _ = _rigs("cfg4", inject=False)
    assert grig.roi_final == orig.roi_final and grig.roi_final[2] >= 15359
    frames = [vsb200.synth.frame(i, 0, kw["src_w"], kw["src_h"]) for i in range(kw["n_views"])]
    want, want_mask = orig.compose(frames)
    got = grig.compose([frames])[0]
    _eq(got, want, "config 4 panorama (CV_16SC3)")
    assert grig.st.last_launch_count() == 8  # K1 K2 down2 down1(L3: too large for shared memory) down_tail coarse blend_seam blend_int
    # The 8-rank view-sharded split of this rig (bench.py --gpus 8) walked through on this one handle, rank after rank: the shard
    # plan, every packed per-peer message (pack as the owner, unpack as the reader), each rank's front half over its own tile
    # lists and each rank's strip blend.  The strips add up to the same panorama.
    import torch
    from tests.gpu_util import dev, host, stream
    world, n = 8, kw["n_views"]
    st = grig.st
    owners, strips = [None] * n, []
    for r in range(world):
        st.shard_set(r, world)
        x0, x1, owned = st.shard_info()
        strips.append((x0, x1, owned))
        for v in owned:
            assert owners[v] is None
            owners[v] = r
    assert None not in owners, owners
    d_src = [dev(f) for f in frames]
    bufs = {}
    for r in range(world):
        st.shard_set(r, world)
        st.shard_plan(owners)
        for p in range(world):
            sb = st.shard_peer_bytes(p)[0] if p != r else 0
            if sb:
                assert sb % 16 == 0
                bufs[(r, p)] = torch.zeros(sb, dtype=torch.uint8, device="cuda")
                st.shard_pack(p, 1, bufs[(r, p)].data_ptr(), stream())
    W, H = grig.roi_final[2], grig.roi_final[3]
    total = torch.zeros((H, W, 3), dtype=torch.int32, device="cuda")
    for r in range(world):
        st.shard_set(r, world)
        st.shard_plan(owners)
        for v0, v1 in vsb200.dist.contiguous_runs(strips[r][2]):
            st.feed_batch(v0, v1, 1, [d_src[v].data_ptr() for v in range(v0, v1)], kw["src_w"] * 3, stream())
        for p in range(world):
            rb = st.shard_peer_bytes(p)[1] if p != r else 0
            assert rb == (bufs[(p, r)].numel() if (p, r) in bufs else 0), (p, r, rb)
            if rb:
                st.shard_unpack(p, 1, bufs[(p, r)].data_ptr(), stream())
        out = torch.zeros((H, W, 3), dtype=torch.int16, device="cuda")
        st.blend_batch([out.data_ptr()], W * 6, stream())
        total += out.to(torch.int32)
        del out
    _eq(host(total).astype(np.int16), want, "config 4, sum of the 8 ranks' strips")
    assert len(bufs) >= world  # every rank exchanges with at least its neighbours


@pytest.mark.parametrize("name,kw", [
    # BASELINE.json configs[2] at full size (6 x 1080p -> 7680-wide spherical panorama, CPW on, 5 bands)
    ("cfg3", dict(n_views=6, src_w=1920, src_h=1080, pano_width=7680, num_bands=5, enable_local=True)),
    # BASELINE.json configs[0] at its true shape (2 x 1280x720 -> 4021-wide spherical panorama, enable_local = false, 5 bands)
    ("cfg1", dict(n_views=2, src_w=1280, src_h=720, pano_width=4021, num_bands=5, enable_local=False)),
])
def test_full_size_config3_and_config1_compose(cuda, og, name, kw):
    import vsb200
    CASES[name] = kw
    orig, grig, _ = _rigs(name, inject=False)
    assert grig.roi_final == orig.roi_final and grig.roi_padded == orig.roi_padded and grig.num_bands == orig.num_bands
    for i in range(kw["n_views"]):
        assert grig.geom[i] == orig.blender.view_geom(i), f"view {i} border geometry"
    frames = [vsb200.synth.frame(i, 1, kw["src_w"], kw["src_h"]) for i in range(kw["n_views"])]
    want, want_mask = orig.compose(frames)
    got = grig.compose([frames])[0]
    _eq(got, want, f"{name} panorama (CV_16SC3)")
    _eq((np.abs(got).sum(axis=2) > 0) | (want_mask > 0), want_mask > 0, "output mask support")


def test_batched_compose_and_mesh_swap(cuda, og):
    import vsb200
    orig, grig, kw = _rigs("small4", inject=False, max_batch=5)
    fr = [[vsb200.synth.frame(i, f, kw["src_w"], kw["src_h"]) for i in range(kw["n_views"])] for f in range(5)]
    want = [orig.compose(f)[0] for f in fr]
    got = grig.compose(fr)  # 5 frames: one remap launch pair, then two sub-batches (2 + 3) on the handle's internal streams
    for f in range(5):
        _eq(got[f], want[f], f"batched frame {f}")
    assert grig.st.last_launch_count() == 12  # K1 K2 + 2 x (down2 down_tail coarse blend_seam blend_int)
    got3 = grig.compose(fr[1:4])  # 3 frames: one submission on the caller's stream
    for f in range(3):
        _eq(got3[f], want[1 + f], f"3-frame batch, frame {f}")
    assert grig.st.last_launch_count() == 7
    # install a different mesh (recalibration, config 5) and compose again
    for i in range(kw["n_views"]):
        mx, my = vsb200.synth.mesh(*orig.sizes[i], phase=0.7)
        orig.set_mesh(i, mx, my)
        grig.set_mesh(i, mx, my)
    _eq(grig.compose([fr[1]])[0], orig.compose(fr[1])[0], "after mesh swap")
    # and swap twice without a compose in between (re-uses the unpublished buffer)
    for ph in (0.2, 1.1):
        for i in range(kw["n_views"]):
            mx, my = vsb200.synth.mesh(*orig.sizes[i], phase=ph)
            orig.set_mesh(i, mx, my)
            grig.set_mesh(i, mx, my)
    _eq(grig.compose([fr[2]])[0], orig.compose(fr[2])[0], "after double mesh swap")


@pytest.mark.parametrize("variant,pad", [(-1, 0), (0, 0), (1, 0), (2, 0), (3, 0), (0, 1), (1, 4), (2, 4), (3, 4), (3, 16)])
def test_remap_kernel_variants(cuda, og, tmp_path, variant, pad):
    """Every form of the remap kernels (coordinate-driven; table-driven with consecutive / lane-interleaved pixels, 4 or 2 per thread;
    3 = shared-memory staged source footprints, the default) gives the oracle's frames; pad = 1 makes the caller's rows unaligned
    (falls back to the coordinate-driven kernels), pad = 4 keeps them word but not 16-byte aligned (the staged form falls back to
    the lane-interleaved one), pad = 16 gives 16-byte aligned rows longer than the image (staged, boxes may not run past the last row)."""
    import os
    import subprocess
    import sys
    import vsb200
    orig, _, kw = _rigs("small4", inject=False)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = str(tmp_path / "v.npz")
    env = dict(os.environ, VSB_REMAP_VARIANT=str(variant))
    subprocess.run([sys.executable, os.path.join(root, "tests", "variant_gpu_worker.py"), out, str(pad)], check=True, env=env, timeout=600)
    got = np.load(out)
    for f in range(2):
        frames = [vsb200.synth.frame(i, f, kw["src_w"], kw["src_h"]) for i in range(kw["n_views"])]
        _eq(got[f"pano{f}"], orig.compose(frames)[0], f"variant {variant} pad {pad} frame {f}")


def test_nv12_primitive_matches_oracle(cuda, og, vsb):
    """B-row 'wire format': vsb_nv12_to_bgr = cv::cvtColor(CV_YUV2BGR_NV12), bit-exact, aligned and unaligned pitches."""
    import ctypes as C
    import torch
    import vsb200
    from tests.gpu_util import host, stream
    L = vsb.lib()
    for (w, h, pitch) in ((320, 240, 320), (322, 242, 331), (1920, 1080, 2048)):
        nv = vsb200.synth.frame_nv12(1, 0, w, h)
        buf = torch.zeros((h * 3 // 2, pitch), dtype=torch.uint8, device="cuda")
        buf[:, :w] = torch.from_numpy(nv).cuda()
        out = torch.full((h, w * 3 + 5), 7, dtype=torch.uint8, device="cuda")
        vsb.check(L.vsb_nv12_to_bgr(C.c_void_p(buf.data_ptr()), w, h, C.c_size_t(pitch), C.c_void_p(out.data_ptr()), C.c_size_t(w * 3 + 5), C.c_void_p(stream())))
        got = host(out)
        _eq(got[:, :w * 3].reshape(h, w, 3), og.nv12_to_bgr(nv, w, h), f"nv12 -> bgr {w}x{h} pitch {pitch}")
        assert (got[:, w * 3:] == 7).all()


@pytest.mark.parametrize("case", ["small4", "bands6"])
def test_nv12_input_and_u8_output(cuda, og, case):
    """SURVEY 8f rows 2-3 fused into the path: NV12 frames in (cvtColor on the device), CV_8UC3 panorama out (convertTo fused
    into the blend store); device and host entry points, batched."""
    import torch
    import vsb200
    from tests.gpu_util import dev, host, stream
    B = vsb200.binding
    NF = 4 if case == "small4" else 2   # 4 frames: the split (two-stream) submission with NV12 staging
    orig, grig, kw = _rigs(case, inject=False, max_batch=NF)
    n, sw, sh = kw["n_views"], kw["src_w"], kw["src_h"]
    nv = [[vsb200.synth.frame_nv12(i, f, sw, sh) for i in range(n)] for f in range(NF)]
    want16 = [orig.compose([og.nv12_to_bgr(a, sw, sh) for a in fr])[0] for fr in nv]
    want8 = [og.s16_to_u8(w) for w in want16]
    W, H = grig.roi_final[2], grig.roi_final[3]
    srcs = [dev(a) for fr in nv for a in fr]
    # NV12 in, CV_16SC3 out
    grig.st.set_formats(B.IN_NV12, B.OUT_S16C3)
    outs = [torch.full((H, W, 3), -12345, dtype=torch.int16, device="cuda") for _ in range(NF)]
    grig.st.compose([t.data_ptr() for t in srcs], sw, [o.data_ptr() for o in outs], W * 6, stream())
    for f in range(NF):
        _eq(host(outs[f]), want16[f], f"NV12 in, frame {f}")
    # the conversion runs inside remap #1's tap fetch (k_remap_stage1_nv12): no separate nv12_to_bgr launch
    assert grig.st.last_launch_count() == (12 if NF >= 4 else 7)
    # an ODD row pitch cannot take the fused form (16-bit chroma loads): the BGR staging path (k_nv12_to_bgr) gives the same frames
    odd = sw + 1
    padded = []
    for t in srcs:
        buf = torch.zeros((sh * 3 // 2, odd), dtype=torch.uint8, device="cuda")
        buf[:, :sw] = t
        padded.append(buf)
    outs_p = [torch.full((H, W, 3), -12345, dtype=torch.int16, device="cuda") for _ in range(NF)]
    grig.st.compose([t.data_ptr() for t in padded], odd, [o.data_ptr() for o in outs_p], W * 6, stream())
    for f in range(NF):
        _eq(host(outs_p[f]), want16[f], f"NV12 in (odd pitch, staged conversion), frame {f}")
    assert grig.st.last_launch_count() == (13 if NF >= 4 else 8)
    # NV12 in, CV_8UC3 out, pitched output
    grig.st.set_formats(B.IN_NV12, B.OUT_U8C3)
    pitch = (W * 3 + 63) // 64 * 64
    outs8 = [torch.full((H, pitch), 99, dtype=torch.uint8, device="cuda") for _ in range(NF)]
    grig.st.compose([t.data_ptr() for t in srcs], sw, [o.data_ptr() for o in outs8], pitch, stream())
    for f in range(NF):
        got = host(outs8[f])
        _eq(got[:, :W * 3].reshape(H, W, 3), want8[f], f"NV12 in, CV_8UC3 out, frame {f}")
        assert (got[:, W * 3:] == 99).all(), "row padding must stay untouched"
    # host entry point with both formats
    houts = [np.zeros((H, W, 3), np.uint8) for _ in range(NF)]
    grig.st.compose_host([a.ctypes.data for fr in nv for a in fr], sw, [o.ctypes.data for o in houts], W * 3)
    for f in range(NF):
        _eq(houts[f], want8[f], f"host path NV12 -> CV_8UC3 frame {f}")
    # BGR in, CV_8UC3 out through feed + blend
    grig.st.set_formats(B.IN_BGR8, B.OUT_U8C3)
    bgr = [og.nv12_to_bgr(a, sw, sh) for a in nv[1]]
    d_bgr = [dev(a) for a in bgr]
    out1 = torch.zeros((H, W, 3), dtype=torch.uint8, device="cuda")
    for i, t_ in enumerate(d_bgr):
        grig.st.feed(i, t_.data_ptr(), sw * 3, stream())
    grig.st.blend(out1.data_ptr(), W * 3, stream())
    _eq(host(out1), want8[1], "feed + blend, CV_8UC3 out")


@pytest.mark.parametrize("case,out_w,out_h", [("small4", 1280, 640), ("cfg2", 4096, 2048), ("wide2", 512, 300)])
def test_consumer_epilogue(cuda, og, case, out_w, out_h):
    """SURVEY 8f row 2 complete: CV_8UC3 panorama (convertTo fused) -> cv::resize + COLOR_BGR2RGB, and the letter-boxed I420 frame
    the encoder is fed (A/timed.cpp:250-315) -- on the device, bit-exact against the restated CPU consumer."""
    import torch
    import vsb200
    from tests.gpu_util import dev, host, stream
    B = vsb200.binding
    orig, grig, kw = _rigs(case, inject=False)
    frames = [vsb200.synth.frame(i, 2, kw["src_w"], kw["src_h"]) for i in range(kw["n_views"])]
    pano8 = og.s16_to_u8(orig.compose(frames)[0])
    W, H = grig.roi_final[2], grig.roi_final[3]
    grig.st.set_formats(B.IN_BGR8, B.OUT_U8C3)
    pitch = (W * 3 + 15) // 16 * 16
    d_pano = torch.zeros((H, pitch), dtype=torch.uint8, device="cuda")
    srcs = [dev(f) for f in frames]
    grig.st.compose([t.data_ptr() for t in srcs], kw["src_w"] * 3, [d_pano.data_ptr()], pitch, stream())
    _eq(host(d_pano)[:, :W * 3].reshape(H, W, 3), pano8, "CV_8UC3 panorama")
    ih = og.consumer_image_height(W, H, out_w, out_h)
    rgb = torch.full((ih, out_w * 3 + 7), 5, dtype=torch.uint8, device="cuda")
    grig.st.consume(d_pano.data_ptr(), pitch, out_w, out_h, B.CONSUME_RGB, rgb.data_ptr(), out_w * 3 + 7, stream=stream())
    got = host(rgb)
    _eq(got[:, :out_w * 3].reshape(ih, out_w, 3), og.consume(pano8, out_w, out_h, 0), "resize + BGR2RGB")
    assert (got[:, out_w * 3:] == 5).all()
    yuv = torch.zeros(out_w * out_h * 3 // 2, dtype=torch.uint8, device="cuda")
    grig.st.consume(d_pano.data_ptr(), pitch, out_w, out_h, B.CONSUME_I420, yuv.data_ptr(), out_w, stream=stream())
    _eq(host(yuv), og.consume(pano8, out_w, out_h, 1), "letter-boxed I420 frame")


def test_recalibration_thread_concurrent_with_compose(cuda, og):
    """BASELINE config 5: a second host thread keeps publishing new CPW meshes (vsb_set_mesh) while the main thread submits
    batches.  Every composed frame must be exactly the panorama of ONE of the published meshes (never a mixture within a view,
    never a half-written map), and the last publication must be the one in effect afterwards."""
    import threading
    import vsb200
    orig, grig, kw = _rigs("small4", inject=False, max_batch=4)
    n = kw["n_views"]
    fr = [[vsb200.synth.frame(i, f, kw["src_w"], kw["src_h"]) for i in range(n)] for f in range(4)]
    V = 1  # the view whose mesh alternates (the others keep theirs, so a frame has exactly two possible panoramas)
    meshes = [vsb200.synth.mesh(*orig.sizes[V], phase=ph) for ph in (0.0, 0.9)]
    want = []
    for m in meshes:
        orig.set_mesh(V, *m)
        want.append([orig.compose(f)[0] for f in fr])
    stop, installs, errors = threading.Event(), [0], []

    def recalibrate():
        k = 1
        try:
            while not stop.is_set():
                grig.set_mesh(V, *meshes[k & 1])
                installs[0] += 1
                k += 1
        except Exception as e:  # surfaces in the main thread's assert
            errors.append(e)

    th = threading.Thread(target=recalibrate, daemon=True)
    th.start()
    seen = [0, 0]
    try:
        for it in range(25):
            got = grig.compose(fr)
            for f in range(4):
                which = [np.array_equal(got[f], want[k][f]) for k in range(2)]
                assert any(which), f"iteration {it} frame {f}: panorama matches neither published mesh"
                seen[which.index(True)] += 1
    finally:
        stop.set()
        th.join()
    assert not errors, errors
    assert installs[0] >= 2
    grig.set_mesh(V, *meshes[0])
    got = grig.compose(fr)
    for f in range(4):
        _eq(got[f], want[0][f], f"after the last publication, frame {f}")


def test_shard_compose_single_rank(cuda, og):
    """The native view-sharded entry point with world = 1 (no peer): front / exchange / back on the handle's three internal streams,
    submissions alternating between the two halves of the frame slots and between two caller streams -- every frame bit-exact.
    (The N > 1 exchange itself is covered by tests/test_gpu_shard.py on >= 2 GPUs and by bench.py's in-run parity check.)"""
    import torch
    import vsb200
    from tests.gpu_util import dev, host
    B = vsb200.binding
    orig, grig, kw = _rigs("small4", inject=False, max_batch=4)
    n = kw["n_views"]
    fr = [[vsb200.synth.frame(i, f, kw["src_w"], kw["src_h"]) for i in range(n)] for f in range(6)]
    want = [orig.compose(f)[0] for f in fr]
    grig.st.shard_init(0, 1, B.shard_unique_id())
    assert grig.st.shard_info()[2] == list(range(n))
    W, H = grig.roi_final[2], grig.roi_final[3]
    srcs = [[dev(a) for a in one] for one in fr]
    outs = [torch.full((H, W, 3), -12345, dtype=torch.int16, device="cuda") for _ in range(6)]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    torch.cuda.synchronize()
    for k in range(3):  # three submissions of two frames: slot halves 0, 1, 0
        st_ = streams[k & 1]
        grig.st.shard_compose([t.data_ptr() for f in (2 * k, 2 * k + 1) for t in srcs[f]], kw["src_w"] * 3,
                              [outs[2 * k].data_ptr(), outs[2 * k + 1].data_ptr()], W * 6, st_.cuda_stream)
    torch.cuda.synchronize()
    for f in range(6):
        _eq(host(outs[f]), want[f], f"shard_compose (world 1) frame {f}")
    # a 4-frame submission (no second slot half: fully serialised) on the default stream
    outs4 = [torch.full((H, W, 3), -12345, dtype=torch.int16, device="cuda") for _ in range(4)]
    grig.st.shard_compose([t.data_ptr() for f in range(4) for t in srcs[f]], kw["src_w"] * 3, [o.data_ptr() for o in outs4], W * 6,
                          torch.cuda.current_stream().cuda_stream)
    for f in range(4):
        _eq(host(outs4[f]), want[f], f"shard_compose (world 1, 4 frames) frame {f}")


def test_compose_host_roundtrip(cuda, og):
    import vsb200
    orig, grig, kw = _rigs("small4", inject=False, max_batch=2)
    fr = [[vsb200.synth.frame(i, f, kw["src_w"], kw["src_h"]) for i in range(kw["n_views"])] for f in range(2)]
    W, H = grig.roi_final[2], grig.roi_final[3]
    outs = [np.zeros((H, W, 3), np.int16) for _ in range(2)]
    grig.st.compose_host([a.ctypes.data for f in fr for a in f], kw["src_w"] * 3, [o.ctypes.data for o in outs], W * 6)
    for f in range(2):
        _eq(outs[f], orig.compose(fr[f])[0], f"host-buffer compose frame {f}")


def test_feed_online_boundary_and_pitched_output(cuda, og):
    """B4 inner boundary: MultiBandBlender::feed_online takes the WARPED view (the caller keeps its own remaps); and a pitched
    (16-byte aligned) output buffer takes the vector-store path of k_blend."""
    import torch
    import vsb200
    from tests.gpu_util import dev, host, stream
    orig, grig, kw = _rigs("small4", inject=False)
    frames = [vsb200.synth.frame(i, 3, kw["src_w"], kw["src_h"]) for i in range(kw["n_views"])]
    want, _ = orig.compose(frames)
    W, H = grig.roi_final[2], grig.roi_final[3]
    pitch = (W * 6 + 255) // 256 * 256
    out = torch.full((H, pitch // 2), -12345, dtype=torch.int16, device="cuda")
    warped = [dev(orig.warp_view(i, frames[i])) for i in range(kw["n_views"])]
    for i, t in enumerate(warped):
        grig.st.feed_warped(i, t.data_ptr(), t.shape[1] * 3, stream())
    grig.st.blend(out.data_ptr(), pitch, stream())
    got = host(out)
    _eq(got[:, :W * 3].reshape(H, W, 3), want, "feed_online(warped) + blend, pitched output")
    assert (got[:, W * 3:] == -12345).all(), "row padding must stay untouched"
    # batched compose into pitched outputs
    srcs = [dev(f) for f in frames]
    out2 = torch.full((H, pitch // 2), -12345, dtype=torch.int16, device="cuda")
    grig.st.compose([t.data_ptr() for t in srcs], kw["src_w"] * 3, [out2.data_ptr()], pitch, stream())
    _eq(host(out2)[:, :W * 3].reshape(H, W, 3), want, "compose, pitched output")


def test_cpp_host_demo_matches_oracle(cuda, og, tmp_path):
    """Host code stays C++: examples/stitch_demo.cpp (timed.cpp-shaped, on include/vsb200.hpp) gives the oracle's panoramas."""
    import math
    import os
    import subprocess
    import vsb200
    from oracle import pipeline as op
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "examples", "stitch_demo")
    if not os.path.exists(exe):
        import __graft_entry__ as ge
        ge.build_examples()
    n, sw, sh, pano, nb, nf = 4, 320, 240, 1024, 3, 2
    gains = [1.0 + 0.03 * ((i % 3) - 1) for i in range(n)]
    orig = op.OracleRig(n, sw, sh, pano, num_bands=nb, enable_local=True, gains=[float(np.float32(g)) for g in gains])
    for v in range(n):
        W, H = orig.sizes[v]
        mx, my = vsb200.synth.identity_mesh(W, H)
        for i in range(10):
            for j in range(10):  # libm sin/cos in double, like the C++ demo
                mx[i, j] = np.float32(mx[i, j] + np.float32(6.0 * math.sin(math.pi * i / 9) * math.cos(math.pi * j / 9)))
                my[i, j] = np.float32(my[i, j] + np.float32(4.0 * math.sin(math.pi * j / 9)))
        orig.set_mesh(v, mx, my)
    frames = [[vsb200.synth.frame(i, f, sw, sh) for i in range(n)] for f in range(nf)]
    fin, fout = tmp_path / "frames.bin", tmp_path / "out.bin"
    with open(fin, "wb") as fh:
        for fr in frames:
            for a in fr:
                fh.write(a.tobytes())
    r = subprocess.run([exe, str(n), str(sw), str(sh), str(pano), str(nb), str(nf), str(fin), str(fout)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    W, H = orig.roi_final[2], orig.roi_final[3]
    got = np.fromfile(fout, np.int16).reshape(nf, H, W, 3)
    for f in range(nf):
        _eq(got[f], orig.compose(frames[f])[0], f"C++ demo frame {f}")
    # wire-to-wire variant: NV12 frames in, letter-boxed I420 frames out (setFormats + consume through the C++ adapter)
    ow, oh = 1280, 640
    nv = [[vsb200.synth.frame_nv12(i, f, sw, sh) for i in range(n)] for f in range(nf)]
    with open(fin, "wb") as fh:
        for fr in nv:
            for a in fr:
                fh.write(a.tobytes())
    r = subprocess.run([exe, str(n), str(sw), str(sh), str(pano), str(nb), str(nf), str(fin), str(fout), str(ow), str(oh)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    got = np.fromfile(fout, np.uint8).reshape(nf, ow * oh * 3 // 2)
    for f in range(nf):
        pano8 = og.s16_to_u8(orig.compose([og.nv12_to_bgr(a, sw, sh) for a in nv[f]])[0])
        _eq(got[f], og.consume(pano8, ow, oh, 1), f"C++ demo, NV12 -> I420, frame {f}")


def test_error_behaviour(cuda, vsb):
    # call-order and argument errors surface as status codes + text, never as crashes (the reference CV_Asserts)
    st = vsb.Stitcher(2, 5, True, 1)
    with pytest.raises(vsb.VsbError):
        st.get_roi()
    st.prepare([(0, 0), (50, 0)], [(64, 64), (64, 64)])
    m = np.full((64, 64), 255, np.uint8)
    with pytest.raises(vsb.VsbError):
        st.init_view(1, m.ctypes.data, 64, 64, 64, (50, 0))  # out of order
    st.init_view(0, m.ctypes.data, 64, 64, 64, (0, 0))
    st.init_view(1, m.ctypes.data, 64, 64, 64, (50, 0))
    with pytest.raises(vsb.VsbError):
        st.feed(0, 1, 192)  # no maps yet
    with pytest.raises(vsb.VsbError):
        vsb.Stitcher(0)
    assert b"num_views" in vsb.lib().vsb_last_error()
