"""-m gpu: the split calibration (modular wrap-around ROI; wrapAround, A/defs.h:25) through the C ABI.

vsb_calibrate_rig_split installs every camera that looks across +-pi as TWO views (column windows of its warped image), so no buffer of
the handle is panorama-wide.  The panorama must equal the UNSPLIT oracle's, bit for bit.  The same path runs without a GPU in
tests/test_emulated_pipeline.py::test_product_library_split_calibration_on_the_emulated_runtime; the planner is pinned on the oracle at
the sizes of configs 2 and 4 by tests/test_oracle_wrap_split.py.  (Written after the round's GPU budget was spent.)"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES = {
    "small4": dict(n_views=4, src_w=320, src_h=240, pano_width=1024, num_bands=3, enable_local=True),
    "cyl5": dict(n_views=5, src_w=256, src_h=192, pano_width=800, num_bands=4, enable_local=True, projection=1),      # two cameras wrap
    "nolocal6": dict(n_views=6, src_w=320, src_h=180, pano_width=960, num_bands=5, enable_local=False),
    "cfg2": dict(n_views=6, src_w=1920, src_h=1080, pano_width=3840, num_bands=5, enable_local=True),                 # BASELINE config 2 at full size
}


def _eq(a, b, what):
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    bad = int(np.count_nonzero(a != b))
    assert bad == 0, f"{what}: {bad} of {a.size} samples differ (max |d| = {np.abs(a.astype(np.int64) - b.astype(np.int64)).max()})"


@pytest.mark.parametrize("case", ["small4", "cyl5", "nolocal6", "cfg2"])
def test_split_calibration_composes_the_unsplit_panorama(cuda, og, case):
    import vsb200
    from oracle import pipeline as op
    from tests.gpu_util import GpuSplitRig
    kw = dict(CASES[case])
    n = kw["n_views"]
    gains = vsb200.synth.gains(n)
    orig = op.OracleRig(gains=gains, **kw)
    grig = GpuSplitRig(gains=gains, max_batch=2, **kw)
    assert grig.n > n and grig.roi_final == orig.roi_final and grig.roi_padded == orig.roi_padded and grig.num_bands == orig.num_bands
    W_pano = orig.roi_final[2]
    for k, (cam, x0, fw) in enumerate(grig.win):
        w, h = grig.sizes[k]
        assert (cam, x0, w) == tuple(grig.plan[k]) and fw == orig.sizes[cam][0] and h == orig.sizes[cam][1]
        assert grig.corners[k] == (orig.corners[cam][0] + x0, orig.corners[cam][1])
        assert w <= W_pano // 2 + 2, "no view is panorama-wide"
        _eq(grig.proj_map(k, 0), np.ascontiguousarray(orig.xmaps[cam][:, x0:x0 + w]), f"x projection map view {k}")
        g = grig.geom[k]
        w0 = grig.weight(k, 0)[g["top"]:g["top"] + h, g["left"]:g["left"] + w]
        _eq(np.rint(w0 * 255).astype(np.uint8), np.ascontiguousarray(orig.masks[cam][:, x0:x0 + w]), f"seam mask view {k}")
    if kw["enable_local"]:
        for c in range(n):
            mx, my = vsb200.synth.mesh(*orig.sizes[c])
            orig.set_mesh(c, mx, my)
            grig.set_camera_mesh(c, mx, my)
        for k, (cam, x0, fw) in enumerate(grig.win):
            w = grig.sizes[k][0]
            want = np.ascontiguousarray(orig.mesh_maps[cam][0][:, x0:x0 + w] - np.float32(x0))
            got = grig.mesh_map(k, 0)
            same = (got.view(np.uint32) == want.view(np.uint32)) | (np.isnan(got) & np.isnan(want))
            assert same.all(), f"x mesh map view {k}: {int((~same).sum())} samples differ"
    frames = [[vsb200.synth.frame(i, f, kw["src_w"], kw["src_h"]) for i in range(n)] for f in range(2)]
    got = grig.compose(frames)
    for k, (cam, x0, fw) in enumerate(grig.win):
        g, (w, h) = grig.geom[k], grig.sizes[k]
        done0 = grig.g0_computed(k).astype(bool)
        crop = done0[g["top"]:g["top"] + h, g["left"]:g["left"] + w]
        _eq(grig.warped(k)[crop], np.ascontiguousarray(orig.warp_view(cam, frames[0][cam])[:, x0:x0 + w])[crop], f"warped view {k}")
    for f in range(2):
        _eq(got[f], orig.compose(frames[f])[0], f"composed panorama, frame {f}")
    assert np.count_nonzero(got[0]) > got[0].size // 2


def test_split_wire_formats_and_wrong_view_count(cuda, og):
    """NV12 in (converted inside remap #1's tap fetch; the two views of a camera read the same frame) / CV_8UC3 out; and a handle
    sized for the cameras instead of the views is refused with the number it needs."""
    import torch
    import vsb200
    from oracle import pipeline as op
    from tests.gpu_util import GpuSplitRig, dev, host, stream
    B, S = vsb200.binding, vsb200.synth
    kw = dict(CASES["small4"])
    n, sw, sh = kw["n_views"], kw["src_w"], kw["src_h"]
    gains = S.gains(n)
    orig = op.OracleRig(gains=gains, **kw)
    grig = GpuSplitRig(gains=gains, **kw)
    for c in range(n):
        mx, my = S.mesh(*orig.sizes[c])
        orig.set_mesh(c, mx, my)
        grig.set_camera_mesh(c, mx, my)
    grig.st.set_formats(B.IN_NV12, B.OUT_U8C3)
    nv = [S.frame_nv12(i, 2, sw, sh) for i in range(n)]
    bgr = [og.nv12_to_bgr(a, sw, sh) for a in nv]
    d_nv = [dev(a) for a in nv]
    W, H = grig.roi_final[2], grig.roi_final[3]
    d_out = torch.full((H, W, 3), 0xAB, dtype=torch.uint8, device="cuda")
    grig.st.compose([t.data_ptr() for t in grig.per_view(d_nv)], sw, [d_out.data_ptr()], W * 3, stream())
    _eq(host(d_out), og.s16_to_u8(orig.compose(bgr)[0]), "split rig, NV12 in -> CV_8UC3 out")
    st = B.Stitcher(n, kw["num_bands"], True, 1)
    with pytest.raises(B.VsbError) as e:
        st.calibrate_rig_split(0, kw["pano_width"], n, sw, sh, 90.0, gains)
    assert f"needs {grig.n} views" in str(e.value)


def test_split_on_the_device_calibration_and_gain_refresh(cuda, og):
    """vsb_calibrate_rig_split(on_device = 1): the device calibration's products installed as windows.  Its panorama equals the UNSPLIT
    device calibration's bit for bit (both build the same maps and seams with the device's sinf / cosf), and vsb_estimate_gains --
    one gain per CAMERA -- installs each gain on both views of a split camera."""
    import vsb200
    from tests.gpu_util import GpuRig, GpuSplitRig, dev, stream
    S = vsb200.synth
    kw = dict(CASES["small4"])
    n, sw, sh = kw["n_views"], kw["src_w"], kw["src_h"]
    gains = S.gains(n)
    whole = GpuRig(gains=gains, device_calibration=True, **kw)
    split = GpuSplitRig(gains=gains, device_calibration=True, **kw)
    assert split.n == n + 1 and split.roi_final == whole.roi_final
    for c in range(n):
        mx, my = S.mesh(*whole.sizes[c])
        whole.set_mesh(c, mx, my)
        split.set_camera_mesh(c, mx, my)
    frames = [S.frame(i, 1, sw, sh) for i in range(n)]
    _eq(split.compose([frames])[0], whole.compose([frames])[0], "split vs unsplit device calibration")
    dim = [np.clip(f.astype(np.float32) * (0.8 if i in (1, 2) else 1.0), 0, 255).astype(np.uint8) for i, f in enumerate(frames)]
    srcs = [dev(f) for f in dim]
    ptrs = [t.data_ptr() for t in srcs]
    g_whole = whole.st.estimate_gains(ptrs, sw * 3, apply=True, stream=stream())[:n]
    g_split = split.st.estimate_gains(ptrs, sw * 3, apply=True, stream=stream())[:n]
    assert g_whole == g_split and g_whole[1] > g_whole[0], (g_whole, g_split)
    _eq(split.compose([dim])[0], whole.compose([dim])[0], "after the run-time gain refresh (camera 2 is the split one)")
