"""Oracle-side restatement of the WHOLE path (calibration + per-frame compose), built from oracle-G primitives.

TEST INFRASTRUCTURE ONLY (see oracle_g.h).  Mirrors, in reference order:
  calibration : 360_stitcher/calibration.cpp:28-249  (calibrateCameras, warpImages) with work_scale = 1; compose_scale = 1 unless
                the rig is built with compose_scale= (then calibration.cpp:137-205 and the per-frame cuda::resize, timed.cpp:74-81)
  per frame   : 360_stitcher/timed.cpp:56-152        (stitch_online x N, stitch_one)
"""
import math

import numpy as np

from . import oracle as og

SEAM_MEGAPIX = 0.01  # 360_stitcher/defs.h:52


def identity_mesh(W, H, rows=10, cols=10):
    """MeshWarper::createMesh initial grid, 360_stitcher/meshwarper.cpp:74-79 (float arithmetic as written there)."""
    mx = np.empty((rows, cols), np.float32)
    my = np.empty((rows, cols), np.float32)
    for i in range(rows):
        for j in range(cols):
            mx[i, j] = np.float32(np.float32(j) * np.float32(W)) / np.float32(cols - 1)
            my[i, j] = np.float32(np.float32(i) * np.float32(H)) / np.float32(rows - 1)
    return mx, my


def synthetic_mesh(W, H, rows=10, cols=10, phase=0.0):
    """identity grid + (6 sin(pi i/9) cos(pi j/9), 4 sin(pi j/9)) px displacement."""
    mx, my = identity_mesh(W, H, rows, cols)
    i = np.arange(rows, dtype=np.float64)[:, None]
    j = np.arange(cols, dtype=np.float64)[None, :]
    dx = 6.0 * np.sin(math.pi * i / 9 + phase) * np.cos(math.pi * j / 9)
    dy = 4.0 * np.sin(math.pi * j / 9 + phase) * np.ones_like(i)
    return (mx + dx.astype(np.float32)).astype(np.float32), (my + dy.astype(np.float32)).astype(np.float32)


def ref_scales(src_w, src_h, work_megapix=0.6, compose_megapix=1.4):
    """work_scale and compose_scale as stitch_calib / warpImages derive them (calibration.cpp:270-277,140-143; defaults A/defs.h:51-53):
    min(1, sqrt(MEGAPIX * 1e6 / area)), a negative MEGAPIX meaning scale 1."""
    area = src_w * src_h
    ws = 1.0 if work_megapix < 0 else min(1.0, math.sqrt(work_megapix * 1e6 / area))
    cs = 1.0 if compose_megapix <= 0 else min(1.0, math.sqrt(compose_megapix * 1e6 / area))
    return ws, cs


class OracleRig:
    def __init__(self, n_views, src_w, src_h, pano_width, projection=og.PROJ_SPHERICAL, num_bands=5,
                 enable_local=True, gains=None, hfov_deg=90.0, compose_scale=1.0, work_scale=1.0):
        """pano_width > 0: the sphere radius is pano_width / 2 pi (this repository's parametrisation; work_scale must be 1).
        pano_width = 0: the reference's own -- warped_image_scale = (float)cameras[0].focal at work scale (stitch_calib,
        calibration.cpp:283-289), which with work_scale / compose_scale from WORK_MEGAPIX / COMPOSE_MEGAPIX (ref_scales below) is
        stitch_calib's default geometry."""
        self.n, self.src_w, self.src_h = n_views, src_w, src_h
        self.work_scale = float(work_scale)
        assert pano_width > 0 or pano_width == 0, pano_width
        # compose_scale (calibration.cpp:137-205, timed.cpp:74-81), followed literally: when it is more than 0.1 away from 1 the frames
        # are cuda::resize'd per frame to cvRound(full * scale) (:159-160 = the dsize cuda::resize computes) and the blender is sized
        # from that; the maps and masks are ALWAYS built for (int)(full * scale) (:204); the cameras and the warper are always scaled.
        # src_w x src_h stays the size of the caller's frames (full_img_size).
        self.compose_scale = float(compose_scale)
        self.scaled = abs(self.compose_scale - 1) > 1e-1          # the reference's own test, timed.cpp:75 / calibration.cpp:157
        self.comp_w, self.comp_h = src_w, src_h                   # the frame remap #1 reads
        if self.scaled:
            self.comp_w, self.comp_h = int(np.rint(src_w * self.compose_scale)), int(np.rint(src_h * self.compose_scale))
        self.map_src = (int(src_w * self.compose_scale), int(src_h * self.compose_scale))   # img_size of buildMaps / the mask warp
        self.projection, self.enable_local = projection, enable_local
        self.gains = [1.0] * n_views if gains is None else [float(g) for g in gains]
        # cameras at work scale (calibrateCameras, calibration.cpp:28-68); warped_image_scale (:283)
        self.K, self.R = zip(*[og.rig_camera_work(n_views, i, src_w, src_h, hfov_deg, self.work_scale, 1.0) for i in range(n_views)])
        self.scale = np.float32(pano_width / (2.0 * 3.1415926535897932384626)) if pano_width > 0 else np.float32(self.K[0][0, 0])

        # ---- seam scale: warp all-255 masks (NEAREST / CONSTANT), Voronoi (calibration.cpp:92-135)
        seam_scale = min(1.0, math.sqrt(SEAM_MEGAPIX * 1e6 / (src_w * src_h)))
        seam_work_aspect = seam_scale / self.work_scale                                  # :280
        seam_w, seam_h = int(np.rint(src_w * seam_scale)), int(np.rint(src_h * seam_scale))
        seam_warp_scale = np.float32(float(self.scale) * seam_work_aspect)                # static_cast<float>(warped_image_scale * seam_work_aspect), :103
        swa = np.float32(seam_work_aspect)
        seam_masks, seam_corners, seam_sizes = [], [], []
        self.seam_scale, self.seam_size, self.seam_maps = seam_scale, (seam_w, seam_h), []
        ones = np.full((seam_h, seam_w), 255, np.uint8)
        for i in range(n_views):
            Ks = self.K[i].copy()
            Ks[0, 0] *= swa; Ks[0, 2] *= swa; Ks[1, 1] *= swa; Ks[1, 2] *= swa
            roi = og.warp_roi(projection, seam_warp_scale, Ks, self.R[i], seam_w, seam_h)
            xm, ym = og.build_maps(projection, seam_warp_scale, Ks, self.R[i], *roi)
            seam_masks.append(og.remap_nearest_u8c1(ones, xm, ym))
            seam_corners.append(roi[:2]); seam_sizes.append(roi[2:])
            self.seam_maps.append((xm, ym))
        self.seam_warped_masks = [m.copy() for m in seam_masks]  # what the exposure compensator gets (before the seam finder)
        og.voronoi_find(seam_sizes, seam_corners, seam_masks)
        self.seam_masks, self.seam_corners, self.seam_sizes = seam_masks, seam_corners, seam_sizes

        # ---- compose scale: ROIs, prepare, maps, masks, init_gpu (calibration.cpp:137-246)
        compose_work_aspect = self.compose_scale / self.work_scale                        # :148
        if compose_work_aspect != 1.0:
            # warper scale: warped_image_scale * static_cast<float>(compose_work_aspect) (:151); cameras: focal, ppx, ppy *= aspect (:168-172)
            self.scale = np.float32(self.scale * np.float32(compose_work_aspect))
            self.K, self.R = zip(*[og.rig_camera_work(n_views, i, src_w, src_h, hfov_deg, self.work_scale, compose_work_aspect) for i in range(n_views)])
        prep = [og.warp_roi(projection, self.scale, self.K[i], self.R[i], self.comp_w, self.comp_h) for i in range(n_views)]   # :176-178
        self.rois = [og.warp_roi(projection, self.scale, self.K[i], self.R[i], *self.map_src) for i in range(n_views)]         # buildMaps' own roi
        self.corners = [r[:2] for r in prep]                 # where the blender puts view i
        self.prep_sizes = [r[2:] for r in prep]              # what prepare() sizes the panorama from
        self.sizes = [r[2:] for r in self.rois]              # size of the maps, the masks and the warped views (= prep_sizes unless round != trunc)
        self.blender = og.Blender(num_bands)
        self.blender.prepare(self.corners, self.prep_sizes)
        self.xmaps, self.ymaps, self.masks = [], [], []
        full = np.full((self.map_src[1], self.map_src[0]), 255, np.uint8)
        for i in range(n_views):
            xm, ym = og.build_maps(projection, self.scale, self.K[i], self.R[i], *self.rois[i])
            warped = og.remap_nearest_u8c1(full, xm, ym)
            seam = og.dilate3x3_u8c1(seam_masks[i]) if enable_local else seam_masks[i]
            seam = og.resize_linear_u8c1(seam, warped.shape[1], warped.shape[0])
            mask = np.bitwise_and(seam, warped)
            self.blender.init_view(mask, self.corners[i])
            self.xmaps.append(xm); self.ymaps.append(ym); self.masks.append(mask)
        self.mesh_maps = [None] * n_views
        self.roi_final, self.roi_padded = self.blender.dst_roi()
        self.num_bands = self.blender.num_bands

    @classmethod
    def from_products(cls, src_w, src_h, corners, sizes, xmaps, ymaps, masks, num_bands=5, enable_local=True, gains=None):
        """A rig whose static inputs come from elsewhere (e.g. read back from vsb_calibrate_rig_device): only the blender geometry
        and weight pyramids are derived here (prepare + init_gpu)."""
        self = cls.__new__(cls)
        n = len(corners)
        self.n, self.src_w, self.src_h = n, src_w, src_h
        self.enable_local = enable_local
        self.gains = [1.0] * n if gains is None else [float(g) for g in gains]
        self.corners, self.sizes = [tuple(c) for c in corners], [tuple(z) for z in sizes]
        self.blender = og.Blender(num_bands)
        self.blender.prepare(self.corners, self.sizes)
        self.xmaps, self.ymaps, self.masks = list(xmaps), list(ymaps), list(masks)
        for i in range(n):
            self.blender.init_view(self.masks[i], self.corners[i])
        self.mesh_maps = [None] * n
        self.roi_final, self.roi_padded = self.blender.dst_roi()
        self.num_bands = self.blender.num_bands
        return self

    def estimate_gains(self, frames, seam_maps=None):
        """GainCompensator::feed as the application drives it (calibration.cpp:95,118,131): frames -> cuda::resize to seam scale ->
        warp LINEAR / BORDER_REFLECT with the seam-scale maps -> feed with the warped all-255 masks.  Returns float64 gains."""
        seam_w, seam_h = self.seam_size
        maps = self.seam_maps if seam_maps is None else seam_maps
        imgs = []
        for i in range(self.n):
            small = og.cuda_resize_linear_u8(frames[i], seam_w, seam_h, self.seam_scale, self.seam_scale)
            imgs.append(og.remap_u8(small, maps[i][0], maps[i][1], og.INTER_LINEAR, og.BORDER_REFLECT))
        return og.gain_compensator_feed(imgs, self.seam_warped_masks, self.seam_corners, self.seam_sizes)

    def set_mesh(self, i, mesh_x, mesh_y):
        w, h = self.sizes[i]
        self.mesh_maps[i] = og.mesh_to_map(mesh_x, mesh_y, w, h)

    def warp_view(self, i, img):
        """stitch_online up to (not including) feed_online: remap#1 -> gain -> remap#2 (timed.cpp:84-108)."""
        if getattr(self, "scaled", False):   # cuda::resize(full, img, Size(), compose_scale, compose_scale, INTER_LINEAR), timed.cpp:74-77
            img = og.cuda_resize_linear_u8(img, self.comp_w, self.comp_h, self.compose_scale, self.compose_scale)
        p = og.remap_linear_u8(img, self.xmaps[i], self.ymaps[i])
        p = og.gain_u8(p, np.float32(self.gains[i]))
        if self.enable_local:
            assert self.mesh_maps[i] is not None, "set_mesh first"
            p = og.remap_linear_u8(p, self.mesh_maps[i][0], self.mesh_maps[i][1])
        return p

    def compose(self, frames):
        """stitch_one (timed.cpp:123-152): returns (CV_16SC3 pano of roi_final size, mask)."""
        for i in range(self.n):
            self.blender.feed_online(i, self.warp_view(i, frames[i]))
        return self.blender.blend()
