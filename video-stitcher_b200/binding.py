"""ctypes binding of libvsb200.so (the C ABI declared in include/vsb200.h).

Harness glue only: device memory, streams and process groups come from PyTorch in tests/ and bench.py;
all compute happens inside the shared library.  There is no CPU fallback -- if the library or a CUDA
device is missing the calls raise.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VSB200_LIB") or os.path.join(_HERE, "libvsb200.so")  # VSB200_LIB: an alternative build of the same ABI (kernel A/B runs)

OK = 0
PROJ_SPHERICAL = 0
PROJ_CYLINDRICAL = 1
MAX_VIEWS = 16
MAX_BANDS = 7

# every symbol include/vsb200.h declares (tests check the library exports all of them)
SYMBOLS = [
    "vsb_last_error", "vsb_version", "vsb_device_count", "vsb_create", "vsb_destroy", "vsb_warp_roi",
    "vsb_build_maps", "vsb_warp", "vsb_prepare", "vsb_get_roi", "vsb_init_view", "vsb_get_view_geometry", "vsb_set_maps",
    "vsb_set_gain", "vsb_set_mesh", "vsb_custom_resize", "vsb_feed", "vsb_feed_warped", "vsb_blend", "vsb_compose",
    "vsb_compose_host", "vsb_submit_host", "vsb_wait_host", "vsb_last_launch_count", "vsb_remap_linear_u8c3", "vsb_gain_u8",
    "vsb_border_reflect_u8c3_to_s16c3", "vsb_pyr_down_s16c3", "vsb_pyr_up_s16c3", "vsb_pyr_down_f32",
    "vsb_add_src_weight_32f", "vsb_normalize_32f", "vsb_debug_read",
    "vsb_shard_set", "vsb_shard_info", "vsb_shard_rect", "vsb_get_plane",
    "vsb_rig_camera", "vsb_voronoi_seams", "vsb_host_build_maps", "vsb_calibrate_rig", "vsb_rig_info_get", "vsb_get_config", "vsb_set_profiling", "vsb_get_profile",
    "vsb_set_formats", "vsb_nv12_to_bgr", "vsb_consumer_image_height", "vsb_consume",
    "vsb_calibrate_rig_device", "vsb_estimate_gains", "vsb_voronoi_seams_device", "vsb_dilate3x3_u8", "vsb_resize_linear_u8",
    "vsb_gain_compensator_feed", "vsb_shard_unique_id", "vsb_shard_init", "vsb_shard_compose", "vsb_shard_exchange_bytes",
    "vsb_shard_plan", "vsb_shard_peer_bytes", "vsb_shard_pack", "vsb_shard_unpack", "vsb_feed_batch", "vsb_blend_batch",
    "vsb_compose_size", "vsb_rig_camera_scaled", "vsb_set_compose_scale", "vsb_calibrate_rig_scaled",
    "vsb_split_plan", "vsb_calibrate_rig_split", "vsb_view_window",
    "vsb_ref_scales", "vsb_rig_camera_work", "vsb_calibrate_rig_megapix",
]
CONSUME_RGB, CONSUME_I420 = 0, 1
IN_BGR8, IN_NV12 = 0, 1
OUT_S16C3, OUT_U8C3 = 0, 1


class VsbError(RuntimeError):
    pass


class Config(C.Structure):
    _fields_ = [("num_views", C.c_int), ("num_bands", C.c_int), ("enable_local", C.c_int),
                ("max_batch", C.c_int), ("device", C.c_int)]


class RigInfo(C.Structure):
    _fields_ = [("projection", C.c_int), ("scale", C.c_float), ("src_w", C.c_int), ("src_h", C.c_int),
                ("num_views", C.c_int), ("num_bands", C.c_int),
                ("roi_final", C.c_int * 4), ("roi_padded", C.c_int * 4),
                ("view_roi", (C.c_int * 4) * MAX_VIEWS)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise VsbError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no fallback path)")
        _lib = C.CDLL(LIB_PATH)
        _lib.vsb_last_error.restype = C.c_char_p
        _lib.vsb_version.restype = C.c_char_p
    return _lib


def check(rc):
    if rc != OK:
        raise VsbError(f"vsb error {rc}: {lib().vsb_last_error().decode()}")
    return rc


def _vp(x):
    return C.c_void_p(int(x) if x is not None else 0)


def _fp9(a):
    return (C.c_float * 9)(*[float(v) for v in a])


def warp_roi(projection, scale, K, R, src_w, src_h):
    roi = (C.c_int * 4)()
    check(lib().vsb_warp_roi(projection, C.c_float(scale), _fp9(K), _fp9(R), src_w, src_h, roi))
    return tuple(roi)


def shard_unique_id():
    """ncclUniqueId (128 bytes) for vsb_shard_init; call on rank 0 and hand the bytes to every rank."""
    buf = (C.c_char * 128)()
    check(lib().vsb_shard_unique_id(buf))
    return bytes(buf.raw)


def split_plan(projection, pano_width, n_cameras, src_w, src_h, num_bands, hfov_deg=90.0):
    """-> [(camera, x0, width)] per view of the split calibration (cameras that wrap around +-pi become two views)"""
    n = C.c_int()
    cam, x0, w = (C.c_int * MAX_VIEWS)(), (C.c_int * MAX_VIEWS)(), (C.c_int * MAX_VIEWS)()
    check(lib().vsb_split_plan(projection, pano_width, n_cameras, src_w, src_h, C.c_double(hfov_deg), num_bands, C.byref(n), cam, x0, w))
    return [(cam[k], x0[k], w[k]) for k in range(n.value)]


def ref_scales(src_w, src_h, work_megapix=0.6, compose_megapix=1.4):
    ws, cs = C.c_double(), C.c_double()
    check(lib().vsb_ref_scales(src_w, src_h, C.c_double(work_megapix), C.c_double(compose_megapix), C.byref(ws), C.byref(cs)))
    return ws.value, cs.value


def rig_camera_work(n_views, i, src_w, src_h, hfov_deg, work_scale, aspect):
    K = (C.c_float * 9)()
    R = (C.c_float * 9)()
    check(lib().vsb_rig_camera_work(n_views, i, src_w, src_h, C.c_double(hfov_deg), C.c_double(work_scale), C.c_double(aspect), K, R))
    return list(K), list(R)


def compose_size(full_w, full_h, compose_scale):
    """-> (frame size remap #1 reads, img_size of the maps, whether the frames are resized)"""
    frame, map_src, resized = (C.c_int * 2)(), (C.c_int * 2)(), C.c_int()
    check(lib().vsb_compose_size(full_w, full_h, C.c_double(compose_scale), frame, map_src, C.byref(resized)))
    return tuple(frame), tuple(map_src), bool(resized.value)


def rig_camera_scaled(n_views, i, src_w, src_h, hfov_deg, compose_work_aspect):
    K = (C.c_float * 9)()
    R = (C.c_float * 9)()
    check(lib().vsb_rig_camera_scaled(n_views, i, src_w, src_h, C.c_double(hfov_deg), C.c_double(compose_work_aspect), K, R))
    return list(K), list(R)


def rig_camera(n_views, i, src_w, src_h, hfov_deg=90.0):
    K = (C.c_float * 9)()
    R = (C.c_float * 9)()
    check(lib().vsb_rig_camera(n_views, i, src_w, src_h, C.c_double(hfov_deg), K, R))
    return list(K), list(R)


class Stitcher:
    """Thin RAII wrapper over a vsb_stitcher handle; pointer arguments are raw device/host addresses."""

    def __init__(self, num_views, num_bands=5, enable_local=True, max_batch=1, device=-1):
        self.cfg = Config(num_views, num_bands, int(bool(enable_local)), max_batch, device)
        self._h = C.c_void_p()
        check(lib().vsb_create(C.byref(self.cfg), C.byref(self._h)))
        self.num_views = num_views

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().vsb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- static setup
    def prepare(self, corners_xy, sizes_wh):
        n = self.num_views
        c = (C.c_int * (2 * n))(*[int(v) for p in corners_xy for v in p])
        s = (C.c_int * (2 * n))(*[int(v) for p in sizes_wh for v in p])
        check(lib().vsb_prepare(self._h, c, s))

    def get_roi(self):
        a, b, nb = (C.c_int * 4)(), (C.c_int * 4)(), C.c_int()
        check(lib().vsb_get_roi(self._h, a, b, C.byref(nb)))
        return tuple(a), tuple(b), nb.value

    def init_view(self, i, mask_ptr, w, h, pitch, tl, on_device=False):
        check(lib().vsb_init_view(self._h, i, _vp(mask_ptr), w, h, C.c_size_t(pitch), int(tl[0]), int(tl[1]), int(on_device)))

    def view_geometry(self, i):
        g = (C.c_int * 8)()
        check(lib().vsb_get_view_geometry(self._h, i, g))
        return dict(zip(["top", "bottom", "left", "right", "x_tl", "y_tl", "x_br", "y_br"], g))

    def set_maps(self, i, xmap_ptr, ymap_ptr, w, h, pitch, src_w, src_h, on_device=False):
        check(lib().vsb_set_maps(self._h, i, _vp(xmap_ptr), _vp(ymap_ptr), w, h, C.c_size_t(pitch), int(on_device), src_w, src_h))

    def set_gain(self, i, gain):
        check(lib().vsb_set_gain(self._h, i, C.c_float(gain)))

    def set_mesh(self, i, mesh_x_ptr, mesh_y_ptr, rows, cols):
        check(lib().vsb_set_mesh(self._h, i, _vp(mesh_x_ptr), _vp(mesh_y_ptr), rows, cols))

    def calibrate_rig(self, projection, pano_width, src_w, src_h, hfov_deg=90.0, gains=None):
        g = None
        if gains is not None:
            g = (C.c_float * self.num_views)(*[float(v) for v in gains])
        check(lib().vsb_calibrate_rig(self._h, projection, pano_width, src_w, src_h, C.c_double(hfov_deg), g))

    def calibrate_rig_device(self, projection, pano_width, src_w, src_h, hfov_deg=90.0, gains=None):
        g = None
        if gains is not None:
            g = (C.c_float * self.num_views)(*[float(v) for v in gains])
        check(lib().vsb_calibrate_rig_device(self._h, projection, pano_width, src_w, src_h, C.c_double(hfov_deg), g))

    def calibrate_rig_scaled(self, projection, pano_width, src_w, src_h, compose_scale, hfov_deg=90.0, gains=None, on_device=False):
        g = None
        if gains is not None:
            g = (C.c_float * self.num_views)(*[float(v) for v in gains])
        check(lib().vsb_calibrate_rig_scaled(self._h, projection, pano_width, src_w, src_h, C.c_double(hfov_deg), g, C.c_double(compose_scale),
                                             int(bool(on_device))))

    def calibrate_rig_split(self, projection, pano_width, n_cameras, src_w, src_h, hfov_deg=90.0, gains=None, on_device=False):
        g = None
        if gains is not None:
            g = (C.c_float * n_cameras)(*[float(v) for v in gains])
        check(lib().vsb_calibrate_rig_split(self._h, projection, pano_width, n_cameras, src_w, src_h, C.c_double(hfov_deg), g, int(bool(on_device))))

    def view_window(self, view):
        """-> (camera, x0, full width of the camera's warped image)"""
        cam, x0, fw = C.c_int(), C.c_int(), C.c_int()
        check(lib().vsb_view_window(self._h, view, C.byref(cam), C.byref(x0), C.byref(fw)))
        return cam.value, x0.value, fw.value

    def calibrate_rig_megapix(self, projection, src_w, src_h, work_megapix=0.6, compose_megapix=1.4, hfov_deg=90.0, gains=None, on_device=False):
        """stitch_calib with its own constants: the reference's default panorama geometry"""
        g = None
        if gains is not None:
            g = (C.c_float * self.num_views)(*[float(v) for v in gains])
        check(lib().vsb_calibrate_rig_megapix(self._h, projection, src_w, src_h, C.c_double(hfov_deg), g, C.c_double(work_megapix),
                                              C.c_double(compose_megapix), int(bool(on_device))))

    def set_compose_scale(self, compose_scale, full_w, full_h):
        check(lib().vsb_set_compose_scale(self._h, C.c_double(compose_scale), full_w, full_h))

    def estimate_gains(self, frame_ptrs, pitch, apply=False, stream=0):
        fp = (C.c_void_p * len(frame_ptrs))(*[int(p) for p in frame_ptrs])
        out = (C.c_float * self.num_views)()
        check(lib().vsb_estimate_gains(self._h, fp, C.c_size_t(pitch), out, int(bool(apply)), _vp(stream)))
        return [float(v) for v in out]

    def rig_info(self):
        info = RigInfo()
        check(lib().vsb_rig_info_get(self._h, C.byref(info)))
        return info

    # ---- per frame
    def feed(self, i, src_ptr, pitch, stream=0):
        check(lib().vsb_feed(self._h, i, _vp(src_ptr), C.c_size_t(pitch), _vp(stream)))

    def feed_warped(self, i, warped_ptr, pitch, stream=0):
        check(lib().vsb_feed_warped(self._h, i, _vp(warped_ptr), C.c_size_t(pitch), _vp(stream)))

    def blend(self, out_ptr, out_pitch, stream=0):
        check(lib().vsb_blend(self._h, _vp(out_ptr), C.c_size_t(out_pitch), _vp(stream)))

    def compose(self, src_ptrs, src_pitch, out_ptrs, out_pitch, stream=0):
        n_frames = len(out_ptrs)
        assert len(src_ptrs) == n_frames * self.num_views
        sp = (C.c_void_p * len(src_ptrs))(*[int(p) for p in src_ptrs])
        op = (C.c_void_p * n_frames)(*[int(p) for p in out_ptrs])
        check(lib().vsb_compose(self._h, n_frames, sp, C.c_size_t(src_pitch), op, C.c_size_t(out_pitch), _vp(stream)))

    def make_compose_call(self, src_ptrs, src_pitch, out_ptrs, out_pitch, stream=0):
        """Pre-marshalled compose call (keeps ctypes overhead out of timed loops)."""
        n_frames = len(out_ptrs)
        sp = (C.c_void_p * len(src_ptrs))(*[int(p) for p in src_ptrs])
        op = (C.c_void_p * n_frames)(*[int(p) for p in out_ptrs])
        fn, h, a, b, st = lib().vsb_compose, self._h, C.c_size_t(src_pitch), C.c_size_t(out_pitch), _vp(stream)

        def call():
            rc = fn(h, n_frames, sp, a, op, b, st)
            if rc != OK:
                check(rc)
        call._keep = (sp, op)
        return call

    def compose_host(self, src_ptrs, src_pitch, out_ptrs, out_pitch):
        n_frames = len(out_ptrs)
        sp = (C.c_void_p * len(src_ptrs))(*[int(p) for p in src_ptrs])
        op = (C.c_void_p * n_frames)(*[int(p) for p in out_ptrs])
        check(lib().vsb_compose_host(self._h, n_frames, sp, C.c_size_t(src_pitch), op, C.c_size_t(out_pitch)))

    def make_submit_host_call(self, src_ptrs, src_pitch, out_ptrs, out_pitch):
        """Pre-marshalled vsb_submit_host (asynchronous; pair every call with wait_host())."""
        n_frames = len(out_ptrs)
        sp = (C.c_void_p * len(src_ptrs))(*[int(p) for p in src_ptrs])
        op = (C.c_void_p * n_frames)(*[int(p) for p in out_ptrs])
        fn, h, a, b = lib().vsb_submit_host, self._h, C.c_size_t(src_pitch), C.c_size_t(out_pitch)

        def call():
            rc = fn(h, n_frames, sp, a, op, b)
            if rc != OK:
                check(rc)
        call._keep = (sp, op)
        return call

    def wait_host(self):
        check(lib().vsb_wait_host(self._h))

    # ---- view-sharded mode
    def shard_set(self, rank, world):
        check(lib().vsb_shard_set(self._h, rank, world))

    def shard_info(self):
        a, b, m = C.c_int(), C.c_int(), C.c_uint()
        check(lib().vsb_shard_info(self._h, C.byref(a), C.byref(b), C.byref(m)))
        return a.value, b.value, [i for i in range(self.num_views) if m.value >> i & 1]

    def shard_rect(self, dst_rank, view, level):
        r = (C.c_int * 4)()
        check(lib().vsb_shard_rect(self._h, dst_rank, view, level, r))
        return tuple(r)

    def shard_plan(self, owners):
        arr = (C.c_int * len(owners))(*[int(o) for o in owners])
        check(lib().vsb_shard_plan(self._h, arr))

    def shard_peer_bytes(self, peer):
        a, b = C.c_size_t(), C.c_size_t()
        check(lib().vsb_shard_peer_bytes(self._h, peer, C.byref(a), C.byref(b)))
        return a.value, b.value

    def shard_pack(self, peer, n_frames, buf_ptr, stream=0):
        check(lib().vsb_shard_pack(self._h, peer, n_frames, _vp(buf_ptr), _vp(stream)))

    def shard_unpack(self, peer, n_frames, buf_ptr, stream=0):
        check(lib().vsb_shard_unpack(self._h, peer, n_frames, _vp(buf_ptr), _vp(stream)))

    def shard_init(self, rank, world, unique_id):
        """Native view-sharded mode: unique_id = the 128 bytes of shard_unique_id() from rank 0 (same on every rank)."""
        buf = (C.c_char * 128).from_buffer_copy(bytes(unique_id))
        check(lib().vsb_shard_init(self._h, rank, world, buf))

    def make_shard_compose_call(self, src_ptrs, src_pitch, out_ptrs, out_pitch, stream=0):
        """src_ptrs[f * num_views + v] (0 for views this rank does not own); out_ptrs[f]: full-size panoramas."""
        n_frames = len(out_ptrs)
        sp = (C.c_void_p * len(src_ptrs))(*[int(p) for p in src_ptrs])
        op = (C.c_void_p * n_frames)(*[int(p) for p in out_ptrs])
        fn, h, a, b, st = lib().vsb_shard_compose, self._h, C.c_size_t(src_pitch), C.c_size_t(out_pitch), _vp(stream)

        def call():
            rc = fn(h, n_frames, sp, a, op, b, st)
            if rc != OK:
                check(rc)
        call._keep = (sp, op)
        return call

    def shard_compose(self, src_ptrs, src_pitch, out_ptrs, out_pitch, stream=0):
        self.make_shard_compose_call(src_ptrs, src_pitch, out_ptrs, out_pitch, stream)()

    def shard_exchange_bytes(self):
        a, b = C.c_size_t(), C.c_size_t()
        check(lib().vsb_shard_exchange_bytes(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def feed_batch(self, v0, v1, n_frames, src_ptrs, pitch, stream=0):
        sp = (C.c_void_p * len(src_ptrs))(*[int(p) for p in src_ptrs])
        check(lib().vsb_feed_batch(self._h, v0, v1, n_frames, sp, C.c_size_t(pitch), _vp(stream)))

    def blend_batch(self, out_ptrs, out_pitch, stream=0):
        op = (C.c_void_p * len(out_ptrs))(*[int(p) for p in out_ptrs])
        check(lib().vsb_blend_batch(self._h, len(out_ptrs), op, C.c_size_t(out_pitch), _vp(stream)))

    def get_plane(self, view, level, frame=0):
        p, w, h = C.c_void_p(), C.c_int(), C.c_int()
        check(lib().vsb_get_plane(self._h, view, level, frame, C.byref(p), C.byref(w), C.byref(h)))
        return p.value, w.value, h.value

    def last_launch_count(self):
        return lib().vsb_last_launch_count(self._h)

    def consume(self, pano_u8_ptr, pitch, out_w, out_h, fmt, out_ptr, out_pitch, keep_aspect=True, stream=0):
        """Consumer epilogue on the device: resize + RGB, or letter-boxed I420 (A/timed.cpp:254-315)."""
        check(lib().vsb_consume(self._h, _vp(pano_u8_ptr), C.c_size_t(pitch), out_w, out_h, int(keep_aspect), int(fmt),
                                _vp(out_ptr), C.c_size_t(out_pitch), _vp(stream)))

    def set_formats(self, input_format=IN_BGR8, output_format=OUT_S16C3):
        """NV12 frames in (cvtColor on the device) and / or CV_8UC3 panoramas out (convertTo(CV_8U) fused into the blend)."""
        check(lib().vsb_set_formats(self._h, int(input_format), int(output_format)))

    def set_profiling(self, on):
        check(lib().vsb_set_profiling(self._h, int(bool(on))))

    def get_profile(self):
        n = C.c_int()
        names = (C.c_char_p * 16)()
        ms = (C.c_float * 16)()
        nbytes = (C.c_double * 16)()
        check(lib().vsb_get_profile(self._h, 16, C.byref(n), names, ms, nbytes))
        return [(names[i].decode(), float(ms[i]), float(nbytes[i])) for i in range(n.value)]

    def debug_read(self, what, view, level, frame, host_ptr, nbytes):
        check(lib().vsb_debug_read(self._h, what, view, level, frame, _vp(host_ptr), C.c_size_t(nbytes)))
