"""ctypes binding of oracle-G (oracle/oracle_g.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product path (video-stitcher_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle_g.so")

PROJ_SPHERICAL = 0
PROJ_CYLINDRICAL = 1


def build(force=False):
    src = os.path.join(_HERE, "oracle_g.c")
    if (not force and os.path.exists(_SO)
            and os.path.getmtime(_SO) >= max(os.path.getmtime(src), os.path.getmtime(os.path.join(_HERE, "oracle_g.h")))):
        return _SO
    os.makedirs(os.path.dirname(_SO), exist_ok=True)
    subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.og_blender_create.restype = C.c_void_p
        _lib.og_blender_view_weight.restype = C.POINTER(C.c_float)
        _lib.og_blender_dst_weight.restype = C.POINTER(C.c_float)
        _lib.og_blender_dst_level.restype = C.POINTER(C.c_int16)
        _lib.og_blender_src_level.restype = C.POINTER(C.c_int16)
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def set_num_threads(n):
    lib().og_set_num_threads(int(n))


# ---------------------------------------------------------------- geometry
def rig_camera(n_views, i, src_w, src_h, hfov_deg=90.0):
    K = np.zeros(9, np.float32)
    R = np.zeros(9, np.float32)
    lib().og_rig_camera(n_views, i, src_w, src_h, C.c_double(hfov_deg), _p(K, C.c_float), _p(R, C.c_float))
    return K.reshape(3, 3), R.reshape(3, 3)


def rig_camera_scaled(n_views, i, src_w, src_h, hfov_deg, compose_work_aspect):
    """the camera after calibration.cpp:168-172 scaled focal / ppx / ppy by compose_work_aspect (doubles), as float K"""
    K = np.zeros(9, np.float32)
    R = np.zeros(9, np.float32)
    lib().og_rig_camera_scaled(n_views, i, src_w, src_h, C.c_double(hfov_deg), C.c_double(compose_work_aspect), _p(K, C.c_float), _p(R, C.c_float))
    return K.reshape(3, 3), R.reshape(3, 3)


def rig_camera_work(n_views, i, src_w, src_h, hfov_deg, work_scale, aspect):
    """calibrateCameras at a work scale, then focal / ppx / ppy *= aspect (doubles), as float K"""
    K = np.zeros(9, np.float32)
    R = np.zeros(9, np.float32)
    lib().og_rig_camera_work(n_views, i, src_w, src_h, C.c_double(hfov_deg), C.c_double(work_scale), C.c_double(aspect), _p(K, C.c_float), _p(R, C.c_float))
    return K.reshape(3, 3), R.reshape(3, 3)


def projector(K, R):
    K = _f32(K).reshape(9)
    R = _f32(R).reshape(9)
    a, b, c = (np.zeros(9, np.float32) for _ in range(3))
    lib().og_projector(_p(K, C.c_float), _p(R, C.c_float), _p(a, C.c_float), _p(b, C.c_float), _p(c, C.c_float))
    return a.reshape(3, 3), b.reshape(3, 3), c.reshape(3, 3)  # k_rinv, r_kinv, rinv


def warp_roi(proj, scale, K, R, src_w, src_h):
    K = _f32(K).reshape(9)
    R = _f32(R).reshape(9)
    roi = (C.c_int * 4)()
    lib().og_warp_roi(proj, C.c_float(scale), _p(K, C.c_float), _p(R, C.c_float), src_w, src_h, roi)
    return tuple(roi)


def build_maps(proj, scale, K, R, tl_x, tl_y, w, h):
    K = _f32(K).reshape(9)
    R = _f32(R).reshape(9)
    xm = np.empty((h, w), np.float32)
    ym = np.empty((h, w), np.float32)
    lib().og_build_maps(proj, C.c_float(scale), _p(K, C.c_float), _p(R, C.c_float), tl_x, tl_y, w, h,
                        _p(xm, C.c_float), _p(ym, C.c_float))
    return xm, ym


# ---------------------------------------------------------------- remap & co
def remap_linear_u8(src, xmap, ymap):
    src = np.ascontiguousarray(src, np.uint8)
    cn = 1 if src.ndim == 2 else src.shape[2]
    sh, sw = src.shape[:2]
    xmap, ymap = _f32(xmap), _f32(ymap)
    dh, dw = xmap.shape
    dst = np.empty((dh, dw) if src.ndim == 2 else (dh, dw, cn), np.uint8)
    lib().og_remap_linear_u8(_p(src, C.c_uint8), sw, sh, cn, C.c_size_t(sw * cn), _p(xmap, C.c_float), _p(ymap, C.c_float),
                             C.c_size_t(dw), _p(dst, C.c_uint8), dw, dh, C.c_size_t(dw * cn))
    return dst


def remap_nearest_u8c1(src, xmap, ymap):
    src = np.ascontiguousarray(src, np.uint8)
    sh, sw = src.shape
    xmap, ymap = _f32(xmap), _f32(ymap)
    dh, dw = xmap.shape
    dst = np.empty((dh, dw), np.uint8)
    lib().og_remap_nearest_u8c1(_p(src, C.c_uint8), sw, sh, C.c_size_t(sw), _p(xmap, C.c_float), _p(ymap, C.c_float),
                                C.c_size_t(dw), _p(dst, C.c_uint8), dw, dh, C.c_size_t(dw))
    return dst


INTER_NEAREST, INTER_LINEAR = 0, 1
BORDER_CONSTANT, BORDER_REFLECT = 0, 2


def remap_u8(src, xmap, ymap, interp=INTER_LINEAR, border=BORDER_CONSTANT):
    """cuda::remap as RotationWarperGpu::warp calls it: NEAREST / LINEAR x BORDER_CONSTANT(0) / BORDER_REFLECT, 1 or 3 channels."""
    src = np.ascontiguousarray(src, np.uint8)
    cn = 1 if src.ndim == 2 else src.shape[2]
    sh, sw = src.shape[:2]
    xmap, ymap = _f32(xmap), _f32(ymap)
    dh, dw = xmap.shape
    dst = np.empty((dh, dw) if src.ndim == 2 else (dh, dw, cn), np.uint8)
    lib().og_remap_u8_border(_p(src, C.c_uint8), sw, sh, cn, C.c_size_t(sw * cn), _p(xmap, C.c_float), _p(ymap, C.c_float),
                             C.c_size_t(dw), _p(dst, C.c_uint8), dw, dh, C.c_size_t(dw * cn), int(interp), int(border))
    return dst


def cuda_resize_linear_u8(src, dw, dh, fx=0.0, fy=0.0):
    """cuda::resize INTER_LINEAR (CV_8UC1 / CV_8UC3); fx = fy = 0: factors from the sizes, else the explicit-scale form."""
    src = np.ascontiguousarray(src, np.uint8)
    cn = 1 if src.ndim == 2 else src.shape[2]
    sh, sw = src.shape[:2]
    dst = np.empty((dh, dw) if src.ndim == 2 else (dh, dw, cn), np.uint8)
    lib().og_cuda_resize_linear_u8(_p(src, C.c_uint8), sw, sh, cn, _p(dst, C.c_uint8), dw, dh, C.c_double(fx), C.c_double(fy))
    return dst


def gain_compensator_feed(imgs, masks, corners_xy, sizes_wh):
    """GainCompensator::feed on warped CV_8UC3 images + CV_8U masks: returns the gains (float64, n >= 4)."""
    n = len(imgs)
    imgs = [np.ascontiguousarray(a, np.uint8) for a in imgs]
    masks = [np.ascontiguousarray(a, np.uint8) for a in masks]
    ip = (C.POINTER(C.c_uint8) * n)(*[_p(a, C.c_uint8) for a in imgs])
    mp = (C.POINTER(C.c_uint8) * n)(*[_p(a, C.c_uint8) for a in masks])
    sz = (C.c_int * (2 * n))(*[int(v) for p in sizes_wh for v in p])
    co = (C.c_int * (2 * n))(*[int(v) for p in corners_xy for v in p])
    g = np.zeros(n, np.float64)
    rc = lib().og_gain_compensator_feed(n, ip, mp, sz, co, _p(g, C.c_double))
    if rc != 0:
        raise ValueError("gain_compensator_feed: singular system or fewer than 4 views")
    return g


def gain_u8(img, gain):
    out = np.ascontiguousarray(img, np.uint8).copy()
    lib().og_gain_u8(_p(out, C.c_uint8), C.c_size_t(out.size), C.c_float(gain))
    return out


def nv12_to_bgr(nv12, w, h):
    """nv12: (h * 3 // 2, w) uint8 (Y plane then interleaved UV) -> (h, w, 3) BGR"""
    nv12 = np.ascontiguousarray(nv12, np.uint8)
    assert nv12.shape == (h * 3 // 2, w) and w % 2 == 0 and h % 2 == 0
    dst = np.empty((h, w, 3), np.uint8)
    lib().og_nv12_to_bgr(_p(nv12, C.c_uint8), w, h, C.c_size_t(w), _p(dst, C.c_uint8), C.c_size_t(w * 3))
    return dst


def s16_to_u8(a):
    a = np.ascontiguousarray(a, np.int16)
    dst = np.empty(a.shape, np.uint8)
    lib().og_s16_to_u8(_p(a, C.c_int16), C.c_size_t(a.size), _p(dst, C.c_uint8))
    return dst


def resize_linear_u8c3(src, dw, dh):
    src = np.ascontiguousarray(src, np.uint8)
    sh, sw, _ = src.shape
    dst = np.empty((dh, dw, 3), np.uint8)
    lib().og_resize_linear_u8c3(_p(src, C.c_uint8), sw, sh, C.c_size_t(sw * 3), _p(dst, C.c_uint8), dw, dh, C.c_size_t(dw * 3))
    return dst


def bgr_to_i420(bgr):
    bgr = np.ascontiguousarray(bgr, np.uint8)
    h, w, _ = bgr.shape
    dst = np.empty(w * h * 3 // 2, np.uint8)
    lib().og_bgr_to_i420(_p(bgr, C.c_uint8), w, h, C.c_size_t(w * 3), _p(dst, C.c_uint8))
    return dst


def consumer_image_height(src_w, src_h, out_w, out_h, keep_aspect=True):
    return int(lib().og_consumer_image_height(src_w, src_h, out_w, out_h, int(keep_aspect)))


def consume(pano_u8, out_w, out_h, fmt, keep_aspect=True):
    """The consumer thread of 360_stitcher/timed.cpp:254-315 after the download: resize, then RGB (fmt 0) or the letter-boxed
    out_w x out_h frame as I420 (fmt 1)."""
    h, w, _ = pano_u8.shape
    ih = consumer_image_height(w, h, out_w, out_h, keep_aspect)
    img = resize_linear_u8c3(pano_u8, out_w, ih)
    if fmt == 0:
        return np.ascontiguousarray(img[..., ::-1])
    canvas = np.zeros((out_h, out_w, 3), np.uint8)
    row0 = out_h // 2 - ih // 2
    canvas[row0:row0 + ih] = img
    return bgr_to_i420(canvas)


def resize_linear_u8c1(src, dw, dh):
    src = np.ascontiguousarray(src, np.uint8)
    sh, sw = src.shape
    dst = np.empty((dh, dw), np.uint8)
    lib().og_resize_linear_u8c1(_p(src, C.c_uint8), sw, sh, _p(dst, C.c_uint8), dw, dh)
    return dst


def dilate3x3_u8c1(src):
    src = np.ascontiguousarray(src, np.uint8)
    h, w = src.shape
    dst = np.empty_like(src)
    lib().og_dilate3x3_u8c1(_p(src, C.c_uint8), w, h, _p(dst, C.c_uint8))
    return dst


# ---------------------------------------------------------------- mesh
def add_src_weight_32f(src, weight, dst, dst_weight):
    """addSrcWeightKernel32F on CV_16SC3 src / dst and CV_32F weights: returns (dst, dst_weight) after the accumulation."""
    src = np.ascontiguousarray(src, np.int16)
    weight = np.ascontiguousarray(weight, np.float32)
    dst = np.array(dst, np.int16, order="C")
    dw = np.array(dst_weight, np.float32, order="C")
    rows, cols = weight.shape
    lib().og_add_src_weight_32f(_p(src, C.c_int16), _p(weight, C.c_float), _p(dst, C.c_int16), _p(dw, C.c_float), rows, cols)
    return dst, dw


def normalize_32f(weight, src):
    """normalizeUsingWeightKernel32F: returns the normalised CV_16SC3 array."""
    weight = np.ascontiguousarray(weight, np.float32)
    out = np.array(src, np.int16, order="C")
    rows, cols = weight.shape
    lib().og_normalize_32f(_p(weight, C.c_float), _p(out, C.c_int16), rows, cols)
    return out


def custom_resize(inp, tx, ty):
    inp = _f32(inp)
    rows, cols = inp.shape
    out = np.empty((ty, tx), np.float32)
    lib().og_custom_resize(_p(inp, C.c_float), cols, rows, _p(out, C.c_float), tx, ty)
    return out


def mesh_to_half_table(mesh_x, mesh_y, W, H):
    mesh_x, mesh_y = _f32(mesh_x), _f32(mesh_y)
    rows, cols = mesh_x.shape
    wx = np.empty((H // 2, W // 2), np.float32)
    wy = np.empty((H // 2, W // 2), np.float32)
    lib().og_mesh_to_half_table(_p(mesh_x, C.c_float), _p(mesh_y, C.c_float), rows, cols, W, H,
                                _p(wx, C.c_float), _p(wy, C.c_float))
    return wx, wy


def mesh_to_map(mesh_x, mesh_y, W, H):
    mesh_x, mesh_y = _f32(mesh_x), _f32(mesh_y)
    rows, cols = mesh_x.shape
    mx = np.empty((H, W), np.float32)
    my = np.empty((H, W), np.float32)
    lib().og_mesh_to_map(_p(mesh_x, C.c_float), _p(mesh_y, C.c_float), rows, cols, W, H, _p(mx, C.c_float), _p(my, C.c_float))
    return mx, my


# ---------------------------------------------------------------- pyramids
def border_reflect_u8c3_to_s16(img, top, bottom, left, right):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape[:2]
    dst = np.empty((h + top + bottom, w + left + right, 3), np.int16)
    lib().og_border_reflect_u8c3_to_s16(_p(img, C.c_uint8), w, h, C.c_size_t(w * 3), top, bottom, left, right, _p(dst, C.c_int16))
    return dst


def _pyr(fn, src, up):
    src = np.ascontiguousarray(src, np.int16)
    cn = 1 if src.ndim == 2 else src.shape[2]
    h, w = src.shape[:2]
    dh, dw = (2 * h, 2 * w) if up else ((h + 1) // 2, (w + 1) // 2)
    dst = np.empty((dh, dw) if src.ndim == 2 else (dh, dw, cn), np.int16)
    fn(_p(src, C.c_int16), w, h, cn, _p(dst, C.c_int16))
    return dst


def pyr_down_s16(src):
    return _pyr(lib().og_pyr_down_s16, src, False)


def pyr_up_s16(src):
    return _pyr(lib().og_pyr_up_s16, src, True)


def pyr_down_s16_int(src):
    return _pyr(lib().og_pyr_down_s16_int, src, False)


def pyr_up_s16_int(src):
    return _pyr(lib().og_pyr_up_s16_int, src, True)


def pyr_down_s16_halfup(src):
    return _pyr(lib().og_pyr_down_s16_halfup, src, False)


def pyr_up_s16_halfup(src):
    return _pyr(lib().og_pyr_up_s16_halfup, src, True)


def pyr_down_f32(src):
    src = _f32(src)
    h, w = src.shape
    dst = np.empty(((h + 1) // 2, (w + 1) // 2), np.float32)
    lib().og_pyr_down_f32(_p(src, C.c_float), w, h, _p(dst, C.c_float))
    return dst


# ---------------------------------------------------------------- seams
def voronoi_find(sizes_wh, corners_xy, masks):
    """masks: list of (h,w) uint8 arrays, modified in place (like VoronoiSeamFinder::find)."""
    n = len(masks)
    sizes = np.ascontiguousarray(np.array(sizes_wh, np.int32).reshape(-1))
    corners = np.ascontiguousarray(np.array(corners_xy, np.int32).reshape(-1))
    for m in masks:
        assert m.dtype == np.uint8 and m.flags["C_CONTIGUOUS"]
    ptrs = (C.POINTER(C.c_uint8) * n)(*[_p(m, C.c_uint8) for m in masks])
    lib().og_voronoi_find(n, _p(sizes, C.c_int), _p(corners, C.c_int), ptrs)
    return masks


# ---------------------------------------------------------------- blender
class Blender:
    """Restatement of the authors' GPU MultiBandBlender (prepare / init_gpu / feed_online / blend)."""

    def __init__(self, num_bands=5):
        self._h = C.c_void_p(lib().og_blender_create(num_bands))
        self.n_views = 0

    def __del__(self):
        if getattr(self, "_h", None):
            lib().og_blender_destroy(self._h)
            self._h = None

    def prepare(self, corners_xy, sizes_wh):
        c = np.ascontiguousarray(np.array(corners_xy, np.int32).reshape(-1))
        s = np.ascontiguousarray(np.array(sizes_wh, np.int32).reshape(-1))
        lib().og_blender_prepare(self._h, len(c) // 2, _p(c, C.c_int), _p(s, C.c_int))
        self.n_views = 0

    @property
    def num_bands(self):
        return lib().og_blender_num_bands(self._h)

    def dst_roi(self):
        a = (C.c_int * 4)()
        b = (C.c_int * 4)()
        lib().og_blender_dst_roi(self._h, a, b)
        return tuple(a), tuple(b)

    def init_view(self, mask, tl):
        mask = np.ascontiguousarray(mask, np.uint8)
        h, w = mask.shape
        r = lib().og_blender_init_view(self._h, _p(mask, C.c_uint8), w, h, C.c_size_t(w), int(tl[0]), int(tl[1]))
        assert r >= 0
        self.n_views += 1
        return r

    def set_cpu_pyramids(self, on):
        lib().og_blender_set_cpu_pyramids(self._h, int(bool(on)))

    def set_view_weight(self, i, level, w):
        w = _f32(w)
        lib().og_blender_set_view_weight(self._h, i, level, _p(w, C.c_float))

    def view_geom(self, i):
        g = (C.c_int * 8)()
        lib().og_blender_view_geom(self._h, i, g)
        return dict(zip(["top", "bottom", "left", "right", "x_tl", "y_tl", "x_br", "y_br"], g))

    def _arr(self, ptr, w, h, cn, dt):
        n = w.value * h.value * cn
        a = np.ctypeslib.as_array(ptr, shape=(n,)).astype(dt, copy=True)
        return a.reshape(h.value, w.value) if cn == 1 else a.reshape(h.value, w.value, cn)

    def view_weight(self, i, level):
        w, h = C.c_int(), C.c_int()
        p = lib().og_blender_view_weight(self._h, i, level, C.byref(w), C.byref(h))
        return self._arr(p, w, h, 1, np.float32)

    def dst_level(self, level):
        w, h = C.c_int(), C.c_int()
        p = lib().og_blender_dst_level(self._h, level, C.byref(w), C.byref(h))
        return self._arr(p, w, h, 3, np.int16)

    def dst_weight(self, level):
        w, h = C.c_int(), C.c_int()
        p = lib().og_blender_dst_weight(self._h, level, C.byref(w), C.byref(h))
        return self._arr(p, w, h, 1, np.float32)

    def src_level(self, i, level):
        w, h = C.c_int(), C.c_int()
        p = lib().og_blender_src_level(self._h, i, level, C.byref(w), C.byref(h))
        return self._arr(p, w, h, 3, np.int16)

    def feed_online(self, i, img):
        img = np.ascontiguousarray(img, np.uint8)
        h, w = img.shape[:2]
        lib().og_blender_feed_online(self._h, i, _p(img, C.c_uint8), w, h, C.c_size_t(w * 3))

    def blend(self):
        (x, y, W, H), _ = self.dst_roi()
        out = np.empty((H, W, 3), np.int16)
        mask = np.empty((H, W), np.uint8)
        lib().og_blender_blend(self._h, _p(out, C.c_int16), _p(mask, C.c_uint8))
        return out, mask
