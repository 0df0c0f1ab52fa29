#!/usr/bin/env python
"""Generates tests/golden/*.npz from the REFERENCE'S OWN CODE run in this container.

Source of truth: oracle/_ref/libvsref.so = the vendored OpenCV 3.4.0 CPU sources of ultravideo/video-stitcher compiled in
place from /root/reference/sources by oracle/ref.mk (plus the restated CPU branch of MultiBandBlender in
oracle/ref_shim.cpp).  /root/reference does not exist on the GPU box, so the outputs are committed as small fixtures and
tests/test_oracle_pin.py replays the (seeded) inputs through oracle-G.

    python tests/golden/make_golden.py        # needs /root/reference (builds oracle/_ref if missing)

Inputs are regenerated from seeds by `inputs()` below (shared with the test), only outputs are stored.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PYR_SHAPES = [(33, 47), (64, 96), (20, 37)]
RIGS = [  # (n_views, src_w, src_h, pano_width)
    (6, 1920, 1080, 3840), (4, 320, 240, 1024), (6, 1920, 1080, 7680), (12, 3840, 2160, 15360), (2, 1280, 720, 4021),
    (5, 640, 480, 2000),
]
MAP_CASES = [(0, 6, 1, 1920, 1080, 3840), (1, 6, 2, 1920, 1080, 3840), (0, 4, 0, 320, 240, 1024), (0, 6, 3, 1920, 1080, 3840)]
MAP_STRIDE = 13
SEAM_RIGS = [(6, 1920, 1080, 3840, 0), (4, 320, 240, 1024, 0), (5, 640, 480, 2000, 1)]


def pyr_input(shape, seed=7):
    rng = np.random.default_rng(seed + shape[0] * 1000 + shape[1])
    a = rng.integers(-2000, 2000, shape + (3,), dtype=np.int16)
    a[0, 0] = (32767, -32768, 32767)   # saturation corners
    a[-1, -1] = (-32768, -32768, 32767)
    return a


def weight_input(shape, seed=11):
    rng = np.random.default_rng(seed + shape[0])
    return rng.random(shape).astype(np.float32)


def remap_input(seed=3):
    rng = np.random.default_rng(seed)
    src = rng.integers(0, 256, (60, 80, 3), dtype=np.uint8)
    yy, xx = np.mgrid[0:50, 0:70].astype(np.float32)
    # the 45-degree rotation recipe of sources/modules/cudawarping/test/test_remap.cpp:158-165, reaching outside the image
    c, s = np.float32(np.cos(np.pi / 4)), np.float32(np.sin(np.pi / 4))
    xm = (c * xx - s * yy + 20).astype(np.float32)
    ym = (s * xx + c * yy - 15).astype(np.float32)
    return src, xm, ym


RESIZE_COEFFS = (0.3, 0.5, 1.5, 2.0)  # CW/test/test_resize.cpp:160


def resize_test_recipe(channels, seed=7):
    """CW/test/test_resize.cpp:142-153: a randomMat CV_8UC1 / CV_8UC3 of one of DIFFERENT_SIZES (113 x 113), resized by each coefficient."""
    rng = np.random.default_rng(seed + channels)
    return rng.integers(0, 256, (113, 113) if channels == 1 else (113, 113, channels), dtype=np.uint8)


def remap_test_recipe(size=(128, 128), seed=5):
    """CW/test/test_remap.cpp:127-148 (SetUp): maps of a 45-degree rotation, M = [[cos, -sin, w/2], [sin, cos, 0]], on a
    randomMat CV_8UC3 source of the same size."""
    h, w = size
    rng = np.random.default_rng(seed)
    src = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    a = np.pi / 4
    xm = (np.cos(a) * xx - np.sin(a) * yy + w / 2.0).astype(np.float32)
    ym = (np.sin(a) * xx + np.cos(a) * yy).astype(np.float32)
    return src, xm, ym


def gain_input(seed=23, n=6):
    """Warped-image-like inputs of GainCompensator::feed: n overlapping CV_8UC3 images + CV_8U masks (255 = valid) on a ring."""
    rng = np.random.default_rng(seed)
    sizes = [(40 + 3 * i, 30 + i) for i in range(n)]
    corners = [(25 * i, 2 * (i % 3)) for i in range(n)]
    imgs = [np.clip(rng.integers(0, 256, (h, w, 3)) * (0.8 + 0.08 * i), 0, 255).astype(np.uint8) for i, (w, h) in enumerate(sizes)]
    masks = [np.where(rng.random((h, w)) < 0.9, 255, 0).astype(np.uint8) for (w, h) in sizes]
    return imgs, masks, corners, sizes


def nv12_input(w=64, h=48, seed=17):
    """Random NV12 frame (Y plane then interleaved UV) covering the full byte range, incl. the saturating branches."""
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, (h * 3 // 2, w), dtype=np.uint8), w, h


def s16_input(seed=19):
    rng = np.random.default_rng(seed)
    a = rng.integers(-600, 900, (24, 40, 3)).astype(np.int16)
    a[0, 0] = (-32768, 32767, 255); a[0, 1] = (256, -1, 0)
    return a


def consume_input(seed=23):
    """A panorama-shaped CV_8UC3 image (odd sizes like dst_roi_final) for the consumer epilogue, and its output frame sizes."""
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, (63, 383, 3), dtype=np.uint8), 512, 256


def blend_recipe(size=128, seed=5):
    """Upstream MultiBandBlender.CanBlendTwoImages recipe (sources/modules/stitching/test/test_blenders.cpp:57-72):
    two images, left/right half masks, 5 bands -- on seeded synthetic images because baboon/lena are not vendored."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:size, 0:size].astype(np.float32)
    imgs = []
    for k in range(2):
        base = 128 + 90 * np.sin(xx / (7.0 + 4 * k) + k) * np.cos(yy / (5.0 + 3 * k))
        img = base[..., None] + rng.uniform(-30, 30, (size, size, 3))
        imgs.append(np.clip(np.rint(img), 0, 255).astype(np.uint8))
    m1 = np.zeros((size, size), np.uint8); m1[:, :size // 2] = 255
    m2 = np.zeros((size, size), np.uint8); m2[:, size // 2:] = 255
    return imgs, [m1, m2], [(0, 0), (0, 0)]


def offset_recipe(seed=9):
    """Two 300x200 views at tl = (-7, 13) and (193, 13) with 20-px ramp masks: non-zero borders, padding, 2^nb alignment."""
    rng = np.random.default_rng(seed)
    imgs = [rng.integers(0, 256, (200, 300, 3), dtype=np.uint8) for _ in range(2)]
    ramp = np.clip((np.arange(300, dtype=np.float32) - 140) / 20.0, 0, 1)
    m1 = np.broadcast_to(np.rint(255 * (1 - ramp)).astype(np.uint8), (200, 300)).copy()
    m2 = np.broadcast_to(np.rint(255 * np.clip((np.arange(300, dtype=np.float32) - 80) / 20.0, 0, 1)).astype(np.uint8), (200, 300)).copy()
    return imgs, [m1, m2], [(-7, 13), (193, 13)]


def seam_inputs(og, n, sw, sh, pano, proj):
    """Seam-scale all-255 masks warped NEAREST, as 360_stitcher/calibration.cpp:92-126 does (oracle-G geometry; the ROIs are
    pinned separately against the reference's warpRoi)."""
    import math
    scale = np.float32(pano / (2.0 * 3.1415926535897932384626))
    seam_scale = min(1.0, math.sqrt(0.01 * 1e6 / (sw * sh)))
    seam_w, seam_h = int(np.rint(sw * seam_scale)), int(np.rint(sh * seam_scale))
    sws = np.float32(float(scale) * seam_scale)
    swa = np.float32(seam_scale)
    masks, corners, sizes = [], [], []
    ones = np.full((seam_h, seam_w), 255, np.uint8)
    for i in range(n):
        K, R = og.rig_camera(n, i, sw, sh)
        Ks = K.copy(); Ks[0, 0] *= swa; Ks[0, 2] *= swa; Ks[1, 1] *= swa; Ks[1, 2] *= swa
        roi = og.warp_roi(proj, sws, Ks, R, seam_w, seam_h)
        xm, ym = og.build_maps(proj, sws, Ks, R, *roi)
        masks.append(og.remap_nearest_u8c1(ones, xm, ym))
        corners.append(roi[:2]); sizes.append(roi[2:])
    return masks, corners, sizes


def main():
    from oracle import oracle as og
    from oracle import ref as vr
    if not vr.available() and not vr.build():
        raise SystemExit("oracle/_ref/libvsref.so cannot be built here (no /root/reference)")
    vr.set_num_threads(1)
    out = {}
    # 1. pyramids (cv::pyrDown / cv::pyrUp, sources/modules/imgproc/src/pyramids.cpp)
    for sh in PYR_SHAPES:
        a = pyr_input(sh)
        out[f"pyr_down_s16_{sh[0]}x{sh[1]}"] = vr.pyr_down(a, vr.T_S16C3)
        out[f"pyr_up_s16_{sh[0]}x{sh[1]}"] = vr.pyr_up(a, vr.T_S16C3)
        out[f"pyr_down_f32_{sh[0]}x{sh[1]}"] = vr.pyr_down(weight_input(sh), vr.T_F32C1)
    # 2. copyMakeBorder REFLECT (sources/modules/core/src/copy.cpp)
    img = pyr_input((20, 37)).astype(np.uint8)
    out["border_reflect"] = vr.copy_make_border(img, vr.T_U8C3, 17, 19, 30, 3)  # borders < image size: beyond that the CUDA BrdReflect formula (single fold) and the CPU borderInterpolate loop differ
    # 3. warpRoi for every view of several rigs, both projections (warpers_inl.hpp:150-210, warpers.cpp:277-318)
    rois = []
    for (n, sw, sh, pano) in RIGS:
        scale = np.float32(pano / (2.0 * 3.1415926535897932384626))
        for proj in (0, 1):
            for i in range(n):
                K, R = og.rig_camera(n, i, sw, sh)
                rois.append((n, sw, sh, pano, proj, i) + vr.warp_roi(proj, scale, K, R, sw, sh))
    out["warp_roi"] = np.array(rois, np.int32)
    # 4. buildMaps, subsampled (CPU projectors, warpers_inl.hpp:59-110,263-312)
    for (proj, n, i, sw, sh, pano) in MAP_CASES:
        scale = np.float32(pano / (2.0 * 3.1415926535897932384626))
        K, R = og.rig_camera(n, i, sw, sh)
        xm, ym, roi = vr.build_maps(proj, scale, K, R, sw, sh)
        out[f"maps_{proj}_{n}_{i}_{pano}"] = np.stack([xm[::MAP_STRIDE, ::MAP_STRIDE], ym[::MAP_STRIDE, ::MAP_STRIDE]])
        out[f"maps_roi_{proj}_{n}_{i}_{pano}"] = np.array(roi, np.int32)
    # 5. VoronoiSeamFinder::find (seam_finders.cpp:72-162) + distanceTransform
    for (n, sw, sh, pano, proj) in SEAM_RIGS:
        masks, corners, sizes = seam_inputs(og, n, sw, sh, pano, proj)
        vr.voronoi_find(sizes, corners, masks)
        for i, m in enumerate(masks):
            out[f"voronoi_{n}_{pano}_{proj}_{i}"] = np.packbits(m > 0, axis=1)
    m = (weight_input((40, 55)) > 0.2).astype(np.uint8) * 255
    out["dist_l1"] = vr.distance_l1(m)
    # 6. mask post-processing twins: dilate 3x3, linear resize (calibration.cpp:232-236 run these on the GPU; CPU twins here)
    sm = (weight_input((24, 31), 13) > 0.5).astype(np.uint8) * 255
    out["dilate3x3"] = vr.dilate3x3_u8c1(sm)
    # 7. cv::remap LINEAR/CONSTANT on the rotation recipe (fixed-point CPU arithmetic: documented tolerance, not exact)
    src, xm, ym = remap_input()
    out["remap_linear"] = vr.remap_u8(src, xm, ym)
    out["remap_nearest"] = vr.remap_u8(src[..., 0].copy(), xm, ym, nearest=True)
    # 7b. the reference's FLOAT gold of cuda::remap (CW/test/interpolation.hpp:66-84) on the same recipe and on the recipe at
    #     the size / map of CW/test/test_remap.cpp:127-148 (128x128 source, 45-degree rotation about (w/2, 0)): this is the
    #     arithmetic oracle-G restates (floor, 4 fp32 taps, saturate_cast), unlike the fixed-point CPU cv::remap above
    out["remap_gold_linear"] = vr.remap_gold_u8(src, xm, ym)
    src2, xm2, ym2 = remap_test_recipe()
    out["remap_gold_recipe"] = vr.remap_gold_u8(src2, xm2, ym2)
    # 7c. the reference's float gold of cuda::resize INTER_LINEAR (CW/test/test_resize.cpp:54-74) on its own recipe
    for cn in (1, 3):
        for c in RESIZE_COEFFS:
            out[f"resize_gold_c{cn}_{c}"] = vr.resize_gold_u8(resize_test_recipe(cn), c, c)
    # 8. gain convertTo
    out["gain_1.03"] = vr.gain_u8(np.arange(256, dtype=np.uint8), 1.03)
    out["gain_0.97"] = vr.gain_u8(np.arange(256, dtype=np.uint8), 0.97)
    # 9. oracle-C blender on the upstream recipe and on the offset recipe
    for name, (imgs, masks, tls) in (("recipe", blend_recipe()), ("offset", offset_recipe())):
        b = vr.BlenderC(5)
        b.prepare(tls, [(im.shape[1], im.shape[0]) for im in imgs])
        for i in range(2):
            b.add_view(masks[i], tls[i])
        for i in range(2):
            b.feed(i, imgs[i])
        res, mask = b.blend()
        out[f"blend_{name}_out"] = res
        out[f"blend_{name}_mask"] = np.packbits(mask > 0, axis=1)
        out[f"blend_{name}_geom"] = np.array([[g[k] for k in ("top", "bottom", "left", "right", "x_tl", "y_tl", "x_br", "y_br")] for g in b.geom], np.int32)
        out[f"blend_{name}_roi"] = np.array(b.dst_roi(), np.int32)
        for i in range(2):
            for k in range(b.num_bands + 1):
                out[f"blend_{name}_w_{i}_{k}"] = b.view_weight(i, k)
    # 9b. the reference's own GainCompensator::feed (S/src/exposure_compensate.cpp:71-142, compiled from its source) and
    #     cuda::resize's arithmetic twin is pinned through the mask pipeline above; the gains are float64, compared bit for bit
    imgs, masks, corners, sizes = gain_input()
    out["gain_compensator"] = vr.gain_compensator_feed(imgs, masks, corners, sizes)
    # 10. wire / consumer formats: cvtColor(CV_YUV2BGR_NV12) (A/networking.cpp:46), convertTo(CV_8U) (A/timed.cpp:250)
    nv, w, h = nv12_input()
    out["nv12_bgr"] = vr.cvt_nv12_bgr(nv, w, h)
    out["s16_to_u8"] = vr.convert_s16_u8(s16_input())
    # 11. consumer epilogue: resize INTER_LINEAR + BGR2RGB / letter-boxed BGR2YUV_I420 (A/timed.cpp:254-315)
    pano, ow, oh = consume_input()
    out["consume_rgb"] = vr.consume(pano, ow, oh, 0)
    out["consume_i420"] = vr.consume(pano, ow, oh, 1)
    path = os.path.join(HERE, "reference_cpu.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
