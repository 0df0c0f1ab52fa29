"""Helpers for the -m gpu parity tests: device buffers come from torch, compute from libvsb200 via ctypes."""
import ctypes as C

import numpy as np
import torch

import vsb200

B = vsb200.binding


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def stream():
    return torch.cuda.current_stream().cuda_stream


def host(t):
    torch.cuda.synchronize()
    return t.cpu().numpy()


def pitch_of(t):
    return t.stride(0) * t.element_size()


class GpuRig:
    """Drives a vsb_stitcher for a rig; static inputs either from the oracle (inject) or from vsb_calibrate_rig."""

    def __init__(self, n_views, src_w, src_h, pano_width, projection=0, num_bands=5, enable_local=True, gains=None,
                 max_batch=1, oracle_rig=None, device_calibration=False, compose_scale=1.0):
        self.n, self.src_w, self.src_h = n_views, src_w, src_h
        self.st = B.Stitcher(n_views, num_bands, enable_local, max_batch)
        if compose_scale != 1.0 and oracle_rig is None:
            self.st.calibrate_rig_scaled(projection, pano_width, src_w, src_h, compose_scale, 90.0, gains, on_device=device_calibration)
        elif device_calibration:
            self.st.calibrate_rig_device(projection, pano_width, src_w, src_h, 90.0, gains)
        elif oracle_rig is None:
            self.st.calibrate_rig(projection, pano_width, src_w, src_h, 90.0, gains)
        else:
            r = oracle_rig
            self.st.prepare(r.corners, getattr(r, "prep_sizes", r.sizes))
            for i in range(n_views):
                m = np.ascontiguousarray(r.masks[i])
                self.st.init_view(i, m.ctypes.data, m.shape[1], m.shape[0], m.shape[1], r.corners[i])
                xm, ym = np.ascontiguousarray(r.xmaps[i]), np.ascontiguousarray(r.ymaps[i])
                self.st.set_maps(i, xm.ctypes.data, ym.ctypes.data, xm.shape[1], xm.shape[0], xm.shape[1] * 4,
                                 getattr(r, "comp_w", src_w), getattr(r, "comp_h", src_h))
                self.st.set_gain(i, r.gains[i])
            if compose_scale != 1.0:   # the low-level route: maps of the resized frame + vsb_set_compose_scale
                self.st.set_compose_scale(compose_scale, src_w, src_h)
        self.roi_final, self.roi_padded, self.num_bands = self.st.get_roi()
        self.geom = [self.st.view_geometry(i) for i in range(n_views)]
        info = self.st.rig_info()
        self.sizes = [(info.view_roi[i][2], info.view_roi[i][3]) for i in range(n_views)]
        self.corners = [(info.view_roi[i][0], info.view_roi[i][1]) for i in range(n_views)]

    def set_mesh(self, i, mx, my):
        mx = np.ascontiguousarray(mx, np.float32)
        my = np.ascontiguousarray(my, np.float32)
        self.st.set_mesh(i, mx.ctypes.data, my.ctypes.data, mx.shape[0], mx.shape[1])

    def compose(self, frames_per_frame):
        """frames_per_frame: list (n_frames) of list (n_views) of HxWx3 uint8 arrays -> list of (H,W,3) int16."""
        nf = len(frames_per_frame)
        W, H = self.roi_final[2], self.roi_final[3]
        srcs = [dev(f) for fr in frames_per_frame for f in fr]
        outs = [torch.full((H, W, 3), -12345, dtype=torch.int16, device="cuda") for _ in range(nf)]
        self.st.compose([t.data_ptr() for t in srcs], self.src_w * 3, [o.data_ptr() for o in outs], W * 6, stream())
        return [host(o) for o in outs]

    def feed_blend(self, frames):
        W, H = self.roi_final[2], self.roi_final[3]
        srcs = [dev(f) for f in frames]
        out = torch.full((H, W, 3), -12345, dtype=torch.int16, device="cuda")
        for i, t in enumerate(srcs):
            self.st.feed(i, t.data_ptr(), self.src_w * 3, stream())
        self.st.blend(out.data_ptr(), W * 6, stream())
        return host(out)

    def read(self, what, view, level, shape, dtype, frame=0):
        a = np.empty(shape, dtype)
        self.st.debug_read(what, view, level, frame, a.ctypes.data, a.nbytes)
        return a

    def warped(self, i, frame=0):
        w, h = self.sizes[i]
        return self.read(0, i, 0, (h, w, 3), np.uint8, frame)

    def gauss_level(self, i, k, frame=0):
        g = self.geom[i]
        bw, bh = (g["x_br"] - g["x_tl"]) >> k, (g["y_br"] - g["y_tl"]) >> k
        return self.read(1, i, k, (bh, bw, 3), np.int16, frame)

    def g0_computed(self, i):
        """1 where k_remap_stage2 computes the bordered level-0 sample, 0 where its tile is skipped (nothing reads it)."""
        g = self.geom[i]
        bw, bh = g["x_br"] - g["x_tl"], g["y_br"] - g["y_tl"]
        return self.read(7, i, 0, (bh, bw), np.uint8)

    def g2_computed(self, i):
        """1 where k_down2 computes the level-2 sample, 0 where its tile is skipped (nothing reads it)."""
        g = self.geom[i]
        bw, bh = (g["x_br"] - g["x_tl"]) >> 2, (g["y_br"] - g["y_tl"]) >> 2
        return self.read(6, i, 0, (bh, bw), np.uint8)

    def weight(self, i, k):
        g = self.geom[i]
        bw, bh = (g["x_br"] - g["x_tl"]) >> k, (g["y_br"] - g["y_tl"]) >> k
        return self.read(2, i, k, (bh, bw), np.float32)

    def proj_map(self, i, which):
        w, h = self.sizes[i]
        return self.read(8 + which, i, 0, (h, w), np.float32)

    def mesh_map(self, i, which):
        w, h = self.sizes[i]
        return self.read(4 + which, i, 0, (h, w), np.float32)


class GpuSplitRig(GpuRig):
    """vsb_calibrate_rig_split: cameras that wrap around +-pi are installed as two views (column windows of their warped image).
    The interface stays per CAMERA: frames and meshes are given per camera and handed to every view of that camera."""

    def __init__(self, n_views, src_w, src_h, pano_width, projection=0, num_bands=5, enable_local=True, gains=None, max_batch=1,
                 device_calibration=False):
        self.n_cameras, self.src_w, self.src_h = n_views, src_w, src_h
        self.plan = B.split_plan(projection, pano_width, n_views, src_w, src_h, num_bands)
        self.n = len(self.plan)
        self.st = B.Stitcher(self.n, num_bands, enable_local, max_batch)
        self.st.calibrate_rig_split(projection, pano_width, n_views, src_w, src_h, 90.0, gains, on_device=device_calibration)
        self.roi_final, self.roi_padded, self.num_bands = self.st.get_roi()
        self.geom = [self.st.view_geometry(k) for k in range(self.n)]
        info = self.st.rig_info()
        self.sizes = [(info.view_roi[k][2], info.view_roi[k][3]) for k in range(self.n)]
        self.corners = [(info.view_roi[k][0], info.view_roi[k][1]) for k in range(self.n)]
        self.win = [self.st.view_window(k) for k in range(self.n)]   # (camera, x0, width of the camera's warped image)

    def set_camera_mesh(self, cam, mx, my):
        for k in range(self.n):
            if self.win[k][0] == cam:
                self.set_mesh(k, mx, my)

    def per_view(self, per_camera):
        return [per_camera[self.win[k][0]] for k in range(self.n)]

    def compose(self, frames_per_frame):
        return super().compose([self.per_view(fr) for fr in frames_per_frame])
