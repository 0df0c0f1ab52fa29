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
