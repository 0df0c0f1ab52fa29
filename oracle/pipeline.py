"""Oracle-side restatement of the WHOLE path (calibration + per-frame compose), built from oracle-G primitives.

TEST INFRASTRUCTURE ONLY (see oracle_g.h).  Mirrors, in reference order:
  calibration : 360_stitcher/calibration.cpp:28-249  (calibrateCameras, warpImages) with work_scale = compose_scale = 1
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


class OracleRig:
    def __init__(self, n_views, src_w, src_h, pano_width, projection=og.PROJ_SPHERICAL, num_bands=5,
                 enable_local=True, gains=None, hfov_deg=90.0):
        self.n, self.src_w, self.src_h = n_views, src_w, src_h
        self.projection, self.enable_local = projection, enable_local
        self.scale = np.float32(pano_width / (2.0 * 3.1415926535897932384626))
        self.gains = [1.0] * n_views if gains is None else [float(g) for g in gains]
        self.K, self.R = zip(*[og.rig_camera(n_views, i, src_w, src_h, hfov_deg) for i in range(n_views)])

        # ---- seam scale: warp all-255 masks (NEAREST / CONSTANT), Voronoi (calibration.cpp:92-135)
        seam_scale = min(1.0, math.sqrt(SEAM_MEGAPIX * 1e6 / (src_w * src_h)))
        seam_w, seam_h = int(np.rint(src_w * seam_scale)), int(np.rint(src_h * seam_scale))
        seam_warp_scale = np.float32(float(self.scale) * seam_scale)
        swa = np.float32(seam_scale)
        seam_masks, seam_corners, seam_sizes = [], [], []
        ones = np.full((seam_h, seam_w), 255, np.uint8)
        for i in range(n_views):
            Ks = self.K[i].copy()
            Ks[0, 0] *= swa; Ks[0, 2] *= swa; Ks[1, 1] *= swa; Ks[1, 2] *= swa
            roi = og.warp_roi(projection, seam_warp_scale, Ks, self.R[i], seam_w, seam_h)
            xm, ym = og.build_maps(projection, seam_warp_scale, Ks, self.R[i], *roi)
            seam_masks.append(og.remap_nearest_u8c1(ones, xm, ym))
            seam_corners.append(roi[:2]); seam_sizes.append(roi[2:])
        og.voronoi_find(seam_sizes, seam_corners, seam_masks)
        self.seam_masks, self.seam_corners, self.seam_sizes = seam_masks, seam_corners, seam_sizes

        # ---- compose scale: ROIs, prepare, maps, masks, init_gpu (calibration.cpp:137-246)
        self.rois = [og.warp_roi(projection, self.scale, self.K[i], self.R[i], src_w, src_h) for i in range(n_views)]
        self.corners = [r[:2] for r in self.rois]
        self.sizes = [r[2:] for r in self.rois]
        self.blender = og.Blender(num_bands)
        self.blender.prepare(self.corners, self.sizes)
        self.xmaps, self.ymaps, self.masks = [], [], []
        full = np.full((src_h, src_w), 255, np.uint8)
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

    def set_mesh(self, i, mesh_x, mesh_y):
        w, h = self.sizes[i]
        self.mesh_maps[i] = og.mesh_to_map(mesh_x, mesh_y, w, h)

    def warp_view(self, i, img):
        """stitch_online up to (not including) feed_online: remap#1 -> gain -> remap#2 (timed.cpp:84-108)."""
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
