"""-m gpu: compose_scale != 1 (A/calibration.cpp:137-205: COMPOSE_MEGAPIX; A/timed.cpp:74-81: the per-frame cuda::resize in front of
remap #1) through the C ABI, bit-exact against oracle-G (oracle/pipeline.py OracleRig(compose_scale=...)).

The same path is run WITHOUT a GPU by tests/test_emulated_pipeline.py::test_product_library_compose_scale_on_the_emulated_runtime
(the shipped library on the emulated runtime); this file is its hardware twin, at sizes the emulation cannot reach."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _eq(a, b, what):
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    bad = int(np.count_nonzero(a != b))
    assert bad == 0, f"{what}: {bad} of {a.size} samples differ (max |d| = {np.abs(a.astype(np.int64) - b.astype(np.int64)).max()})"


REF_DEFAULT = min(1.0, (1.4e6 / (1920 * 1080)) ** 0.5)   # COMPOSE_MEGAPIX = 1.4 on 1080p (A/defs.h:53): 1578 x 887 frame, 1577 x 887 maps

CASES = {
    "small4": dict(n_views=4, src_w=320, src_h=240, pano_width=1024, num_bands=3, enable_local=True, compose_scale=0.75),
    "mismatch6": dict(n_views=6, src_w=322, src_h=182, pano_width=960, num_bands=4, enable_local=True, compose_scale=0.8),      # 257.6 x 145.6: cvRound != (int)
    "cyl5_nolocal": dict(n_views=5, src_w=256, src_h=192, pano_width=800, num_bands=4, enable_local=False, projection=1, compose_scale=0.5),
    "near_one": dict(n_views=4, src_w=320, src_h=240, pano_width=1024, num_bands=3, enable_local=True, compose_scale=0.95),     # cameras scaled, frames not
    "ref_default_1080p": dict(n_views=6, src_w=1920, src_h=1080, pano_width=3840, num_bands=5, enable_local=True, compose_scale=REF_DEFAULT),
}


def _rigs(case, max_batch=1, device_calibration=False, inject=False):
    import vsb200
    from oracle import pipeline as op
    from tests.gpu_util import GpuRig
    kw = dict(CASES[case])
    gains = vsb200.synth.gains(kw["n_views"])
    orig = op.OracleRig(gains=gains, **kw)
    grig = GpuRig(gains=gains, max_batch=max_batch, device_calibration=device_calibration, oracle_rig=orig if inject else None, **kw)
    return orig, grig, kw


def _meshes(orig, grig, kw):
    import vsb200
    if kw["enable_local"]:
        for i in range(kw["n_views"]):
            mx, my = vsb200.synth.mesh(*grig.sizes[i])
            orig.set_mesh(i, mx, my)
            grig.set_mesh(i, mx, my)


@pytest.mark.parametrize("case", ["small4", "mismatch6", "cyl5_nolocal", "near_one", "ref_default_1080p"])
def test_scaled_calibration_and_compose_match_oracle(cuda, og, case):
    """vsb_calibrate_rig_scaled (host): geometry, maps, seam masks and weight pyramids equal the oracle's; composed frames bit-exact,
    through vsb_compose (two frames per submission) and through vsb_feed + vsb_blend."""
    import vsb200
    orig, grig, kw = _rigs(case, max_batch=2)
    assert grig.roi_final == orig.roi_final and grig.roi_padded == orig.roi_padded and grig.num_bands == orig.num_bands
    frame_sz, map_src, resized = vsb200.binding.compose_size(kw["src_w"], kw["src_h"], kw["compose_scale"])
    assert frame_sz == (orig.comp_w, orig.comp_h) and map_src == tuple(orig.map_src) and resized == orig.scaled
    if case in ("mismatch6", "ref_default_1080p"):
        assert frame_sz != map_src, "this case is meant to hit the reference's cvRound / (int) mismatch"
    for i in range(kw["n_views"]):
        assert grig.geom[i] == orig.blender.view_geom(i), f"view {i} border geometry"
        assert grig.sizes[i] == tuple(orig.sizes[i]) and grig.corners[i] == tuple(orig.corners[i])
        _eq(grig.proj_map(i, 0), orig.xmaps[i], f"x projection map view {i}")
        _eq(grig.proj_map(i, 1), orig.ymaps[i], f"y projection map view {i}")
        for k in range(orig.num_bands + 1):
            _eq(grig.weight(i, k), orig.blender.view_weight(i, k), f"weight pyramid view {i} level {k}")
    _meshes(orig, grig, kw)
    frames = [[vsb200.synth.frame(i, f, kw["src_w"], kw["src_h"]) for i in range(kw["n_views"])] for f in range(2)]
    got = grig.compose(frames)
    for i in range(kw["n_views"]):
        g = grig.geom[i]
        done0 = grig.g0_computed(i).astype(bool)
        crop = done0[g["top"]:g["top"] + grig.sizes[i][1], g["left"]:g["left"] + grig.sizes[i][0]]
        _eq(grig.warped(i)[crop], orig.warp_view(i, frames[0][i])[crop], f"warped view {i}")
    for f in range(2):
        _eq(got[f], orig.compose(frames[f])[0], f"composed panorama, frame {f}")
    assert np.count_nonzero(got[0]) > got[0].size // 2
    _eq(grig.feed_blend(frames[1]), got[1], "vsb_feed + vsb_blend")


def test_scaled_device_calibration_and_low_level_route(cuda, og):
    """The same on the device calibration (maps within 2e-3 px, everything downstream exact: the device's products injected into
    oracle-G give the device's panorama) and through the low-level route a host with its own calibration takes (vsb_prepare /
    vsb_init_view / vsb_set_maps of the resized frame + vsb_set_compose_scale)."""
    import vsb200
    from oracle import pipeline as op
    orig, grig, kw = _rigs("small4", device_calibration=True)
    n = kw["n_views"]
    assert grig.roi_final == orig.roi_final
    xmaps, ymaps, masks = [], [], []
    for i in range(n):
        assert grig.geom[i] == orig.blender.view_geom(i) and grig.sizes[i] == tuple(orig.sizes[i]) and grig.corners[i] == tuple(orig.corners[i])
        gx, gy = grig.proj_map(i, 0), grig.proj_map(i, 1)
        inside = (orig.xmaps[i] > -2) & (orig.xmaps[i] < orig.comp_w + 1) & (orig.ymaps[i] > -2) & (orig.ymaps[i] < orig.comp_h + 1) & ~((orig.xmaps[i] == -1) & (orig.ymaps[i] == -1))
        assert np.abs(gx - orig.xmaps[i])[inside].max() <= 2e-3 and np.abs(gy - orig.ymaps[i])[inside].max() <= 2e-3
        g = grig.geom[i]
        w0 = grig.weight(i, 0)[g["top"]:g["top"] + grig.sizes[i][1], g["left"]:g["left"] + grig.sizes[i][0]]
        m = np.rint(w0 * 255).astype(np.uint8)
        # (a sanity bound: one seam-scale sample that the device's sinf / cosf put on the other side of a mask boundary is a few dozen
        # compose-scale pixels, a seam-scale column a few hundred; the exact statement is the panorama below)
        assert np.count_nonzero(m != orig.masks[i]) <= 0.02 * m.size, f"seam mask view {i}"
        xmaps.append(gx); ymaps.append(gy); masks.append(m)
    mine = op.OracleRig.from_products(kw["src_w"], kw["src_h"], grig.corners, orig.prep_sizes, xmaps, ymaps, masks, kw["num_bands"], True, orig.gains)
    mine.sizes = [tuple(z) for z in orig.sizes]
    mine.scaled, mine.compose_scale, mine.comp_w, mine.comp_h = True, orig.compose_scale, orig.comp_w, orig.comp_h
    _meshes(mine, grig, kw)
    frames = [vsb200.synth.frame(i, 3, kw["src_w"], kw["src_h"]) for i in range(n)]
    _eq(grig.compose([frames])[0], mine.compose(frames)[0], "panorama from the device's scaled calibration")
    # low-level route with the oracle's products
    orig2, grig2, _ = _rigs("small4", inject=True)
    _meshes(orig2, grig2, kw)
    _eq(grig2.compose([frames])[0], orig2.compose(frames)[0], "panorama through vsb_set_maps + vsb_set_compose_scale")
    # maps that address another frame size than the one the resize produces are refused at submission time
    grig2.st.set_compose_scale(0.5, kw["src_w"], kw["src_h"])
    with pytest.raises(vsb200.binding.VsbError):
        grig2.compose([frames])
    grig2.st.set_compose_scale(0.75, kw["src_w"], kw["src_h"])
    _eq(grig2.compose([frames])[0], orig2.compose(frames)[0], "panorama after restoring the scale")


def test_scaled_wire_formats_and_host_path(cuda, og):
    """NV12 frames in (converted at full size, then resized: the order of the reference's capture thread and stitch_online) and
    CV_8UC3 panoramas out; and the host-buffer path (vsb_compose_host takes the FULL-size frames)."""
    import torch
    import vsb200
    from tests.gpu_util import dev, host, stream
    B, S = vsb200.binding, vsb200.synth
    orig, grig, kw = _rigs("small4", max_batch=2)
    _meshes(orig, grig, kw)
    n, sw, sh = kw["n_views"], kw["src_w"], kw["src_h"]
    W, H = grig.roi_final[2], grig.roi_final[3]
    # host path, BGR in / s16 out
    frames = [[S.frame(i, f, sw, sh) for i in range(n)] for f in range(2)]
    srcs = [np.ascontiguousarray(a) for fr in frames for a in fr]
    outs = [np.full((H, W, 3), -1, np.int16) for _ in range(2)]
    grig.st.compose_host([a.ctypes.data for a in srcs], sw * 3, [o.ctypes.data for o in outs], W * 6)
    for f in range(2):
        _eq(outs[f], orig.compose(frames[f])[0], f"vsb_compose_host, frame {f}")
    # NV12 in, u8 out, device buffers
    grig.st.set_formats(B.IN_NV12, B.OUT_U8C3)
    nv = [S.frame_nv12(i, 5, sw, sh) for i in range(n)]
    bgr = [og.nv12_to_bgr(a, sw, sh) for a in nv]
    d_nv = [dev(a) for a in nv]
    d_out = torch.full((H, W, 3), 0xAB, dtype=torch.uint8, device="cuda")
    grig.st.compose([t.data_ptr() for t in d_nv], sw, [d_out.data_ptr()], W * 3, stream())
    _eq(host(d_out), og.s16_to_u8(orig.compose(bgr)[0]), "NV12 in -> resize -> compose -> CV_8UC3 out")


REF_RIG, REF_MEGAPIX = (6, 1920, 1080, 5), (0.6, 1.4)   # (cameras, frame width, frame height, bands), (WORK_MEGAPIX, COMPOSE_MEGAPIX) of A/defs.h:51-53


@pytest.mark.parametrize("on_device", [False, True])
def test_reference_default_scales_at_1080p(cuda, og, on_device):
    """vsb_calibrate_rig_megapix(0.6, 1.4): stitch_calib's defaults on 6 x 1080p -- work_scale 0.538, compose_scale 0.822, sphere radius
    = the work-scale focal length x compose_work_aspect (a 4955-wide panorama), 1578 x 887 frames against 1577 x 887 maps.  Host
    calibration: bit-exact against oracle-G.  Device calibration: the same geometry, panorama equal to oracle-G's from the device's own
    products (as for every device calibration; its maps differ from libm's by ulps)."""
    import torch
    import vsb200
    from oracle import pipeline as op
    from tests.gpu_util import dev, host, stream
    B, S = vsb200.binding, vsb200.synth
    n, sw, sh, nb = REF_RIG
    gains = S.gains(n)
    ws, cs = B.ref_scales(sw, sh, *REF_MEGAPIX)
    orig = op.OracleRig(n, sw, sh, 0, num_bands=nb, enable_local=True, gains=gains, compose_scale=cs, work_scale=ws)
    st = B.Stitcher(n, nb, True, 1)
    st.calibrate_rig_megapix(0, sw, sh, REF_MEGAPIX[0], REF_MEGAPIX[1], 90.0, gains, on_device=on_device)
    roi, rpad, bands = st.get_roi()
    assert tuple(roi) == tuple(orig.roi_final) and tuple(rpad) == tuple(orig.roi_padded) and bands == orig.num_bands
    info = st.rig_info()
    assert abs(info.scale - float(orig.scale)) == 0.0
    xmaps, ymaps, masks = [], [], []
    for i in range(n):
        w, h = info.view_roi[i][2], info.view_roi[i][3]
        assert (w, h) == tuple(orig.sizes[i]) and (info.view_roi[i][0], info.view_roi[i][1]) == tuple(orig.corners[i])
        assert st.view_geometry(i) == orig.blender.view_geom(i)
        a = np.empty((h, w), np.float32); st.debug_read(8, i, 0, 0, a.ctypes.data, a.nbytes)
        b = np.empty((h, w), np.float32); st.debug_read(9, i, 0, 0, b.ctypes.data, b.nbytes)
        g = st.view_geometry(i)
        bw, bh = g["x_br"] - g["x_tl"], g["y_br"] - g["y_tl"]
        w0 = np.empty((bh, bw), np.float32); st.debug_read(2, i, 0, 0, w0.ctypes.data, w0.nbytes)
        m = np.rint(w0[g["top"]:g["top"] + h, g["left"]:g["left"] + w] * 255).astype(np.uint8)
        if on_device:
            inside = (orig.xmaps[i] > -2) & (orig.xmaps[i] < orig.comp_w + 1) & (orig.ymaps[i] > -2) & (orig.ymaps[i] < orig.comp_h + 1) & ~((orig.xmaps[i] == -1) & (orig.ymaps[i] == -1))
            assert np.abs(a - orig.xmaps[i])[inside].max() <= 2e-3 and np.abs(b - orig.ymaps[i])[inside].max() <= 2e-3
        else:
            _eq(a, orig.xmaps[i], f"x projection map view {i}")
            _eq(b, orig.ymaps[i], f"y projection map view {i}")
            _eq(m, orig.masks[i], f"seam mask view {i}")
        xmaps.append(a); ymaps.append(b); masks.append(m)
    ref = orig
    if on_device:   # oracle-G on the device's own products: everything downstream of the maps is exact
        ref = op.OracleRig.from_products(sw, sh, orig.corners, orig.prep_sizes, xmaps, ymaps, masks, nb, True, orig.gains)
        ref.sizes = [tuple(z) for z in orig.sizes]
        ref.scaled, ref.compose_scale, ref.comp_w, ref.comp_h = True, orig.compose_scale, orig.comp_w, orig.comp_h
    for i in range(n):
        mx, my = S.mesh(*orig.sizes[i])
        ref.set_mesh(i, mx, my)
        st.set_mesh(i, mx.ctypes.data, my.ctypes.data, mx.shape[0], mx.shape[1])
    frames = [S.frame(i, 0, sw, sh) for i in range(n)]
    srcs = [dev(f) for f in frames]
    W, H = roi[2], roi[3]
    out = torch.full((H, W, 3), -12345, dtype=torch.int16, device="cuda")
    st.compose([t.data_ptr() for t in srcs], sw * 3, [out.data_ptr()], W * 6, stream())
    _eq(host(out), ref.compose(frames)[0], "panorama at the reference's default scales")
