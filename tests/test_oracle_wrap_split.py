"""CPU, oracle only: the design check behind DESIGN.md section 9 (modular wrap-around ROI).

A view that straddles +-pi has the reference's full-panorama-width ROI (detectResultRoi of the spherical warper): its two parts
sit at the two ends of one image with zeros between them.  Claim: feeding the two parts as two sub-views -- inner edges
zero-extended by a margin of 3 * 2^num_bands (the REFLECT gap), the right part's origin moved by a multiple of 2^num_bands -- gives
the same panorama bit for bit,
because Gaussian, Laplacian and weight levels of the sub-views equal the full-width view's wherever a weight is non-zero.
(The product still allocates the full-width ROI and skips its empty tiles; this test pins the conditions a split has to meet.)"""
import os

import numpy as np
import pytest

import vsb200
from oracle import oracle as og
from oracle import pipeline as op


def _split_rig(rig, margin_units):
    """The rig's blender inputs with every full-width view replaced by its two parts."""
    nb = rig.num_bands
    unit = 1 << nb
    W_pano = rig.roi_final[2]
    parts = []  # (source view, x0, x1) in view columns
    for i in range(rig.n):
        w, h = rig.sizes[i]
        if w < W_pano - 1:
            parts.append((i, 0, w))
            continue
        # columns the projection maps fill from the camera image (what remap #1 can make non-zero; remap #2 moves content by a few px)
        valid = og.remap_nearest_u8c1(np.full((rig.src_h, rig.src_w), 255, np.uint8), rig.xmaps[i], rig.ymaps[i])
        cols = np.nonzero(valid.any(axis=0))[0]
        gaps = np.diff(cols)
        k = int(np.argmax(gaps))
        m = margin_units * unit + 8
        assert gaps[k] > 4 * m, "the two parts must be far apart"
        left_end, right_start = int(cols[k]) + 1, int(cols[k + 1])
        parts.append((i, 0, left_end + m))
        parts.append((i, (right_start - m) // unit * unit, w))  # origin moved by a multiple of 2^nb
    corners = [(rig.corners[i][0] + x0, rig.corners[i][1]) for i, x0, x1 in parts]
    sizes = [(x1 - x0, rig.sizes[i][1]) for i, x0, x1 in parts]
    return parts, corners, sizes


RIG6 = (6, 480, 270, 1536)
RIG4 = (4, 320, 240, 1024)


@pytest.mark.parametrize("rig_kw,nb,margin_units,exact", [(RIG6, 4, 3, True), (RIG6, 5, 3, True), (RIG4, 3, 3, True), (RIG6, 5, 0, False), (RIG4, 3, 0, False)])
def test_wrapped_view_split_into_two_sub_views(rig_kw, nb, margin_units, exact):
    og.build()
    S = vsb200.synth
    n, sw, sh, pano = rig_kw
    rig = op.OracleRig(n, sw, sh, pano, num_bands=nb, enable_local=True, gains=S.gains(n))
    for i in range(n):
        rig.set_mesh(i, *S.mesh(*rig.sizes[i]))
    frames = [S.frame(i, 0, sw, sh) for i in range(n)]
    warped = [rig.warp_view(i, frames[i]) for i in range(n)]
    for i in range(n):
        rig.blender.feed_online(i, warped[i])
    want, want_mask = rig.blender.blend()
    parts, corners, sizes = _split_rig(rig, margin_units)
    assert len(parts) == n + 1, "exactly one view of this rig wraps"
    b = og.Blender(nb)
    b.prepare(corners, sizes)
    assert b.dst_roi() == rig.blender.dst_roi()
    for i, x0, x1 in parts:
        b.init_view(np.ascontiguousarray(rig.masks[i][:, x0:x1]), (rig.corners[i][0] + x0, rig.corners[i][1]))
    for k, (i, x0, x1) in enumerate(parts):
        b.feed_online(k, np.ascontiguousarray(warped[i][:, x0:x1]))
    got, got_mask = b.blend()
    same = np.array_equal(got, want) and np.array_equal(got_mask, want_mask)
    if exact:
        assert same, int(np.count_nonzero(got != want))
    else:
        # cut at the content edge itself, the REFLECT border of the inner edge mirrors image content where the full-width view has
        # zeros: the split is NOT exact then (which makes the margin rule part of the design, not an implementation detail)
        assert not same


FULL_SIZE = [((12, 3840, 2160, 15360), 5, 0)] if os.environ.get("VSB_ORACLE_FULL") else []   # config 4: 40 s on 16 threads, exact (opt-in)


@pytest.mark.parametrize("rig_kw,nb,proj", [(RIG6, 5, 0), (RIG4, 3, 0), ((5, 256, 192, 800), 4, 1), ((2, 640, 360, 2011), 5, 0),
                                            ((6, 1920, 1080, 3840), 5, 0)] + FULL_SIZE)   # the last one: BASELINE config 2 at full size
def test_product_split_plan_is_exact_on_the_oracle(rig_kw, nb, proj):
    """The plan the PRODUCT makes (vsb_split_plan, host code of vsb_calibrate_rig_split: which cameras to split, where to cut, the
    margin, the origin of the right part) fed to the oracle's blender as crops of the full-width views: same panorama, bit for bit.
    (The shipped library runs the same plan end to end in tests/test_emulated_pipeline.py and tests/test_gpu_vsb_wrap_split.py.)"""
    og.build()
    S, B = vsb200.synth, vsb200.binding
    n, sw, sh, pano = rig_kw
    rig = op.OracleRig(n, sw, sh, pano, projection=proj, num_bands=nb, enable_local=True, gains=S.gains(n))
    plan = B.split_plan(proj, pano, n, sw, sh, nb)
    assert len(plan) > n, "at least one camera of these rigs wraps"
    W_pano = rig.roi_final[2]
    for c, x0, w in plan:
        assert x0 % (1 << rig.num_bands) == 0 and 0 < w <= rig.sizes[c][0] and x0 + w <= rig.sizes[c][0]
        assert w <= W_pano // 2 + 2 or rig.sizes[c][0] < W_pano - 1, "no view of a split camera is panorama-wide any more"
    for i in range(n):
        rig.set_mesh(i, *S.mesh(*rig.sizes[i]))
    frames = [S.frame(i, 0, sw, sh) for i in range(n)]
    warped = [rig.warp_view(i, frames[i]) for i in range(n)]
    for i in range(n):
        rig.blender.feed_online(i, warped[i])
    want, want_mask = rig.blender.blend()
    b = og.Blender(nb)
    b.prepare([(rig.corners[c][0] + x0, rig.corners[c][1]) for c, x0, w in plan], [(w, rig.sizes[c][1]) for c, x0, w in plan])
    assert b.dst_roi() == rig.blender.dst_roi()
    for c, x0, w in plan:
        b.init_view(np.ascontiguousarray(rig.masks[c][:, x0:x0 + w]), (rig.corners[c][0] + x0, rig.corners[c][1]))
    for k, (c, x0, w) in enumerate(plan):
        # what the product's remap #2 makes of the window: the mesh map re-based to the window, reading the window of P (zeros outside)
        p = og.gain_u8(og.remap_linear_u8(frames[c], rig.xmaps[c][:, x0:x0 + w], rig.ymaps[c][:, x0:x0 + w]), np.float32(rig.gains[c]))
        mx = np.ascontiguousarray(rig.mesh_maps[c][0][:, x0:x0 + w] - np.float32(x0))
        my = np.ascontiguousarray(rig.mesh_maps[c][1][:, x0:x0 + w])
        q = og.remap_linear_u8(p, mx, my)
        assert np.array_equal(q, warped[c][:, x0:x0 + w]), f"view {k}: remap #2 on the window"
        b.feed_online(k, q)
    got, got_mask = b.blend()
    assert np.array_equal(got, want) and np.array_equal(got_mask, want_mask), int(np.count_nonzero(got != want))
