"""ctypes binding of oracle/_ref/libvsref.so -- the reference's own vendored OpenCV 3.4.0 CPU code compiled
in place by oracle/ref.mk (plus the restated CPU branch of MultiBandBlender, oracle/ref_shim.cpp).

TEST INFRASTRUCTURE ONLY: used by tests/ (to pin oracle-G and to make tests/golden/), and by bench.py's
cpu_baseline / --impl reference legs.  The product path never imports it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "libvsref.so")

T_U8C3, T_S16C3, T_F32C1, T_U8C1 = 0, 1, 2, 3
_DT = {T_U8C3: (np.uint8, 3), T_S16C3: (np.int16, 3), T_F32C1: (np.float32, 1), T_U8C1: (np.uint8, 1)}


def available():
    return os.path.exists(SO)


def build():
    """Compiles the reference sources where they lie (only possible where /root/reference exists)."""
    if os.path.isdir("/root/reference/sources"):
        subprocess.check_call(["make", "-s", "-j8", "-C", _HERE, "-f", "ref.mk"])
    return available()


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(f"{SO} missing: run `make -C oracle ref` where /root/reference is present")
        _lib = C.CDLL(SO)
        _lib.vr_build_info.restype = C.c_char_p
        for f in ("vr_blender_create", "vr_rig_create"):
            getattr(_lib, f).restype = C.c_void_p
    return _lib


def _p(a, t=C.c_void_p):
    return a.ctypes.data_as(t) if t is not C.c_void_p else C.c_void_p(a.ctypes.data)


def set_num_threads(n):
    lib().vr_set_num_threads(int(n))


def get_num_threads():
    return lib().vr_get_num_threads()


def build_info():
    return lib().vr_build_info().decode()


def _typed(a, t):
    dt, cn = _DT[t]
    a = np.ascontiguousarray(a, dt)
    assert (a.ndim == 2 and cn == 1) or (a.ndim == 3 and a.shape[2] == cn)
    return a


def pyr_down(src, t):
    src = _typed(src, t)
    h, w = src.shape[:2]
    dst = np.empty(((h + 1) // 2, (w + 1) // 2) + src.shape[2:], src.dtype)
    lib().vr_pyr_down(_p(src), w, h, t, _p(dst))
    return dst


def pyr_up(src, t):
    src = _typed(src, t)
    h, w = src.shape[:2]
    dst = np.empty((2 * h, 2 * w) + src.shape[2:], src.dtype)
    lib().vr_pyr_up(_p(src), w, h, t, _p(dst))
    return dst


def remap_u8(src, xmap, ymap, nearest=False):
    src = np.ascontiguousarray(src, np.uint8)
    cn = 1 if src.ndim == 2 else src.shape[2]
    sh, sw = src.shape[:2]
    xmap = np.ascontiguousarray(xmap, np.float32)
    ymap = np.ascontiguousarray(ymap, np.float32)
    dh, dw = xmap.shape
    dst = np.empty((dh, dw) if src.ndim == 2 else (dh, dw, cn), np.uint8)
    lib().vr_remap_u8(_p(src), sw, sh, cn, _p(xmap), _p(ymap), _p(dst), dw, dh, int(nearest))
    return dst


def resize_gold_u8(src, fx, fy):
    """The reference's float gold of cuda::resize INTER_LINEAR (CW/test/test_resize.cpp:54-74): dsize = saturate_cast<int>(size * f)."""
    src = np.ascontiguousarray(src, np.uint8)
    cn = 1 if src.ndim == 2 else src.shape[2]
    sh, sw = src.shape[:2]
    dw, dh = int(np.rint(sw * fx)), int(np.rint(sh * fy))
    dst = np.empty((dh, dw) if src.ndim == 2 else (dh, dw, cn), np.uint8)
    lib().vr_resize_gold_u8(_p(src), sw, sh, cn, C.c_double(fx), C.c_double(fy), _p(dst), dw, dh)
    return dst


def remap_gold_u8(src, xmap, ymap):
    """The reference's float gold of cuda::remap LINEAR / BORDER_CONSTANT(0) (CW/test/interpolation.hpp:66-84)."""
    src = np.ascontiguousarray(src, np.uint8)
    cn = 1 if src.ndim == 2 else src.shape[2]
    sh, sw = src.shape[:2]
    xmap = np.ascontiguousarray(xmap, np.float32)
    ymap = np.ascontiguousarray(ymap, np.float32)
    dh, dw = xmap.shape
    dst = np.empty((dh, dw) if src.ndim == 2 else (dh, dw, cn), np.uint8)
    lib().vr_remap_gold_u8(_p(src), sw, sh, cn, _p(xmap), _p(ymap), _p(dst), dw, dh)
    return dst


def gain_compensator_feed(imgs, masks, corners_xy, sizes_wh):
    """The reference's own GainCompensator::feed + gains() (exposure_compensate.cpp:71-142) on warped CV_8UC3 images / CV_8U masks."""
    n = len(imgs)
    imgs = [np.ascontiguousarray(a, np.uint8) for a in imgs]
    masks = [np.ascontiguousarray(a, np.uint8) for a in masks]
    ip = (C.c_void_p * n)(*[a.ctypes.data for a in imgs])
    mp = (C.c_void_p * n)(*[a.ctypes.data for a in masks])
    sz = (C.c_int * (2 * n))(*[int(v) for p in sizes_wh for v in p])
    co = (C.c_int * (2 * n))(*[int(v) for p in corners_xy for v in p])
    g = np.zeros(n, np.float64)
    lib().vr_gain_compensator_feed(n, ip, mp, sz, co, g.ctypes.data_as(C.c_void_p))
    return g


def copy_make_border(src, t, top, bottom, left, right, reflect=True):
    src = _typed(src, t)
    h, w = src.shape[:2]
    dst = np.empty((h + top + bottom, w + left + right) + src.shape[2:], src.dtype)
    lib().vr_copy_make_border(_p(src), w, h, t, top, bottom, left, right, int(reflect), _p(dst))
    return dst


def gain_u8(img, gain):
    out = np.ascontiguousarray(img, np.uint8).copy()
    lib().vr_gain_u8(_p(out), C.c_size_t(out.size), C.c_float(gain))
    return out


def cvt_nv12_bgr(nv12, w, h):
    nv12 = np.ascontiguousarray(nv12, np.uint8)
    dst = np.empty((h, w, 3), np.uint8)
    lib().vr_cvt_nv12_bgr(_p(nv12), w, h, _p(dst))
    return dst


def convert_s16_u8(a):
    a = np.ascontiguousarray(a, np.int16)
    h, w, cn = a.shape
    dst = np.empty(a.shape, np.uint8)
    lib().vr_convert_s16_u8(_p(a), w, h, cn, _p(dst))
    return dst


def consume(pano_u8, out_w, out_h, fmt, keep_aspect=True):
    pano_u8 = np.ascontiguousarray(pano_u8, np.uint8)
    h, w, _ = pano_u8.shape
    out = np.zeros(out_w * out_h * 3, np.uint8)
    ih = lib().vr_consume(_p(pano_u8), w, h, out_w, out_h, int(keep_aspect), fmt, _p(out))
    if fmt == 0:
        return out[:ih * out_w * 3].reshape(ih, out_w, 3).copy()
    return out[:out_w * out_h * 3 // 2].copy()


def resize_linear_u8c1(src, dw, dh):
    src = np.ascontiguousarray(src, np.uint8)
    sh, sw = src.shape
    dst = np.empty((dh, dw), np.uint8)
    lib().vr_resize_linear_u8c1(_p(src), sw, sh, _p(dst), dw, dh)
    return dst


def dilate3x3_u8c1(src):
    src = np.ascontiguousarray(src, np.uint8)
    h, w = src.shape
    dst = np.empty_like(src)
    lib().vr_dilate3x3_u8c1(_p(src), w, h, _p(dst))
    return dst


def distance_l1(src):
    src = np.ascontiguousarray(src, np.uint8)
    h, w = src.shape
    dst = np.empty((h, w), np.float32)
    lib().vr_distance_l1(_p(src), w, h, _p(dst))
    return dst


def _f9(a):
    return (C.c_float * 9)(*[float(v) for v in np.asarray(a, np.float32).reshape(9)])


def warp_roi(proj, scale, K, R, src_w, src_h):
    roi = (C.c_int * 4)()
    lib().vr_warp_roi(proj, C.c_float(scale), _f9(K), _f9(R), src_w, src_h, roi)
    return tuple(roi)


def build_maps(proj, scale, K, R, src_w, src_h):
    roi = warp_roi(proj, scale, K, R, src_w, src_h)
    # buildMaps allocates (br - tl + 1); warpRoi's Rect is tl..br+1 the same way
    xm = np.empty((roi[3], roi[2]), np.float32)
    ym = np.empty((roi[3], roi[2]), np.float32)
    r2 = (C.c_int * 4)()
    lib().vr_build_maps(proj, C.c_float(scale), _f9(K), _f9(R), src_w, src_h, None, None, r2)
    assert (r2[2], r2[3]) == (roi[2], roi[3]), (tuple(r2), roi)
    lib().vr_build_maps(proj, C.c_float(scale), _f9(K), _f9(R), src_w, src_h, _p(xm), _p(ym), r2)
    return xm, ym, tuple(r2)


def voronoi_find(sizes_wh, corners_xy, masks):
    n = len(masks)
    sizes = np.ascontiguousarray(np.array(sizes_wh, np.int32).reshape(-1))
    corners = np.ascontiguousarray(np.array(corners_xy, np.int32).reshape(-1))
    for m in masks:
        assert m.dtype == np.uint8 and m.flags["C_CONTIGUOUS"]
    ptrs = (C.c_void_p * n)(*[m.ctypes.data for m in masks])
    lib().vr_voronoi_find(n, _p(sizes), _p(corners), ptrs)
    return masks


def result_roi(corners_xy, sizes_wh):
    c = np.ascontiguousarray(np.array(corners_xy, np.int32).reshape(-1))
    s = np.ascontiguousarray(np.array(sizes_wh, np.int32).reshape(-1))
    roi = (C.c_int * 4)()
    lib().vr_result_roi(len(c) // 2, _p(c), _p(s), roi)
    return tuple(roi)


class BlenderC:
    """oracle-C: the CPU branch of MultiBandBlender on the vendored primitives."""

    def __init__(self, num_bands=5):
        self._h = C.c_void_p(lib().vr_blender_create(num_bands))
        self.geom = []

    def __del__(self):
        if getattr(self, "_h", None):
            lib().vr_blender_destroy(self._h)
            self._h = None

    def prepare(self, corners_xy, sizes_wh):
        c = np.ascontiguousarray(np.array(corners_xy, np.int32).reshape(-1))
        s = np.ascontiguousarray(np.array(sizes_wh, np.int32).reshape(-1))
        lib().vr_blender_prepare(self._h, len(c) // 2, _p(c), _p(s))
        self.geom = []

    @property
    def num_bands(self):
        return lib().vr_blender_num_bands(self._h)

    def dst_roi(self):
        a, b = (C.c_int * 4)(), (C.c_int * 4)()
        lib().vr_blender_roi(self._h, a, b)
        return tuple(a), tuple(b)

    def add_view(self, mask, tl):
        mask = np.ascontiguousarray(mask, np.uint8)
        h, w = mask.shape
        g = (C.c_int * 8)()
        lib().vr_blender_add_view(self._h, _p(mask), w, h, int(tl[0]), int(tl[1]), g)
        self.geom.append(dict(zip(["top", "bottom", "left", "right", "x_tl", "y_tl", "x_br", "y_br"], g)))

    def view_weight(self, i, level):
        w, h = C.c_int(), C.c_int()
        lib().vr_blender_view_weight(self._h, i, level, None, C.byref(w), C.byref(h))
        out = np.empty((h.value, w.value), np.float32)
        lib().vr_blender_view_weight(self._h, i, level, _p(out), None, None)
        return out

    def feed(self, i, img):
        img = np.ascontiguousarray(img, np.uint8)
        h, w = img.shape[:2]
        lib().vr_blender_feed(self._h, i, _p(img), w, h)

    def blend(self):
        (x, y, W, H), _ = self.dst_roi()
        out = np.empty((H, W, 3), np.int16)
        mask = np.empty((H, W), np.uint8)
        lib().vr_blender_blend(self._h, _p(out), _p(mask))
        return out, mask


class RigC:
    """Whole-frame CPU compose on the reference's OpenCV (the timed CPU baseline).  Static inputs (maps, masks,
    mesh maps, gains) come from an oracle.pipeline.OracleRig so all implementations share them."""

    def __init__(self, orig):
        self.n = orig.n
        self.blender = BlenderC(orig.num_bands)
        self.blender.prepare(orig.corners, orig.sizes)
        for i in range(orig.n):
            self.blender.add_view(orig.masks[i], orig.corners[i])
        self._h = C.c_void_p(lib().vr_rig_create(self.blender._h, orig.n, orig.src_w, orig.src_h, int(orig.enable_local)))
        for i in range(orig.n):
            xm = np.ascontiguousarray(orig.xmaps[i], np.float32)
            ym = np.ascontiguousarray(orig.ymaps[i], np.float32)
            lib().vr_rig_set_view(self._h, i, _p(xm), _p(ym), xm.shape[1], xm.shape[0], C.c_float(orig.gains[i]))
            if orig.enable_local:
                mx, my = orig.mesh_maps[i]
                mx = np.ascontiguousarray(mx, np.float32)
                my = np.ascontiguousarray(my, np.float32)
                lib().vr_rig_set_mesh_maps(self._h, i, _p(mx), _p(my), mx.shape[1], mx.shape[0])
        (_, _, self.W, self.H), _ = self.blender.dst_roi()
        self.src_w = orig.src_w

    def __del__(self):
        if getattr(self, "_h", None):
            lib().vr_rig_destroy(self._h)
            self._h = None

    def compose(self, frames, parallel_views=False, out=None):
        frames = [np.ascontiguousarray(f, np.uint8) for f in frames]
        ptrs = (C.c_void_p * self.n)(*[f.ctypes.data for f in frames])
        if out is None:
            out = np.empty((self.H, self.W, 3), np.int16)
        mask = np.empty((self.H, self.W), np.uint8)
        lib().vr_rig_compose(self._h, ptrs, C.c_size_t(self.src_w * 3), _p(out), _p(mask), int(parallel_views))
        return out, mask
