"""TEST INFRASTRUCTURE: the PRODUCT's primitive kernels (video-stitcher_b200/csrc/vsb_primitives.cu -- the device-launcher layer
of the C ABI) compiled to PTX with the product's own flags and executed on the CPU by oracle/ptx_interp.py,
on the inputs of oracle/ptx_cases.py.  tests/test_oracle_ptx.py compares what they write with what the reference's kernels
write (tests/golden/reference_ptx.npz): a device-code-vs-reference comparison that needs neither a GPU nor the oracle.
Nothing here is linked into, or called by, libvsb200.so."""
import os
import shutil
import struct
import subprocess

import numpy as np

from . import ptx_cases as PC
from . import ptx_interp as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
_parsed = {}


def available():
    return os.path.exists(_NVCC)


def kernels(unit="vsb_primitives"):
    """<unit>.cu -> PTX (the Makefile's flags: -O3 -std=c++17 --fmad=false, sm_100a) -> parsed kernels."""
    if unit not in _parsed:
        out = os.path.join(ROOT, "oracle", "_build", unit + ".ptx")
        os.makedirs(os.path.dirname(out), exist_ok=True)
        src = os.path.join(ROOT, "video-stitcher_b200", "csrc", unit + ".cu")
        subprocess.check_call([_NVCC, "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=compute_100a", "--fmad=false", "-ptx", src, "-o", out])
        _parsed[unit] = P.parse(open(out).read())
    return _parsed[unit]


def _q(v):
    return struct.pack("<Q", int(v))


def _i(v):
    return struct.pack("<i", int(v))


def _f(v):
    return struct.pack("<f", np.float32(v))


def _g(w, h):
    return ((w + 31) // 32, (h + 7) // 8)


def _run(frag, w, h, params, mem):
    P.launch(P.find(kernels(), frag), _g(w, h), (32, 8), params, mem)


def app_resize(inp):
    src, (tx, ty) = inp["in"], [int(v) for v in inp["size"]]
    rows, cols = src.shape
    out = np.zeros((ty, tx), np.float32)
    mem = P.Memory()
    a_in, a_out = mem.add(src), mem.add(out)
    _run("k_custom_resize", tx, ty, [_q(a_in), _i(cols), _i(rows), _q(cols * 4), _q(a_out), _i(tx), _i(ty), _q(tx * 4)], mem)
    return {"out": out}


def multiband_blend(inp):
    rows, cols = inp["w0"].shape
    dst, dw = np.zeros((rows, cols, 3), np.int16), np.zeros((rows, cols), np.float32)
    mem = P.Memory()
    a_d, a_dw = mem.add(dst), mem.add(dw)
    for v in range(3):
        a_s, a_w = mem.add(inp[f"src{v}"]), mem.add(inp[f"w{v}"])
        _run("k_add_src_weight_32f", cols, rows, [_q(a_s), _q(cols * 6), _q(a_w), _q(cols * 4), _q(a_d), _q(cols * 6), _q(a_dw), _q(cols * 4), _i(cols), _i(rows)], mem)
    acc, accw = dst.copy(), dw.copy()
    _run("k_normalize_32f", cols, rows, [_q(a_dw), _q(cols * 4), _q(a_d), _q(cols * 6), _i(cols), _i(rows)], mem)
    return {"acc": acc, "acc_w": accw, "normalized": dst}


def pyr_down(inp):
    out = {}
    for key, frag, elem in (("s16", "k_pyr_down_s16c3", 6), ("f32", "k_pyr_down_f32", 4)):
        src = inp[key]
        h, w = src.shape[:2]
        dh, dw = (h + 1) // 2, (w + 1) // 2
        dst = np.zeros((dh, dw) + src.shape[2:], src.dtype)
        mem = P.Memory()
        a_s, a_d = mem.add(src), mem.add(dst)
        _run(frag, dw, dh, [_q(a_s), _i(w), _i(h), _q(w * elem), _q(a_d), _i(dw), _i(dh), _q(dw * elem)], mem)
        out[key] = dst
    return out


def pyr_up(inp):
    src = inp["s16"]
    h, w = src.shape[:2]
    dst = np.zeros((2 * h, 2 * w, 3), np.int16)
    mem = P.Memory()
    a_s, a_d = mem.add(src), mem.add(dst)
    _run("k_pyr_up_s16c3", 2 * w, 2 * h, [_q(a_s), _i(w), _i(h), _q(w * 6), _q(a_d), _q(2 * w * 6)], mem)
    return {"s16": dst}


def _warp(frag, src, xm, ym, interp, border):
    sh, sw = src.shape[:2]
    cn = 1 if src.ndim == 2 else 3
    dh, dw = xm.shape
    dst = np.zeros((dh, dw) + src.shape[2:], np.uint8)
    mem = P.Memory()
    a_s, a_x, a_y, a_d = mem.add(src), mem.add(xm), mem.add(ym), mem.add(dst)
    base = [_q(a_s), _i(sw), _i(sh), _q(sw * cn), _q(a_x), _q(a_y), _q(dw * 4), _q(a_d), _i(dw), _i(dh), _q(dw * cn)]
    _run(frag, dw, dh, base + ([] if interp is None else [_i(interp), _i(border)]), mem)
    return dst


def remap(inp):
    xm, ym = inp["xmap"], inp["ymap"]
    return {"linear_constant_c3": _warp("k_remap_linear_u8c3", inp["c3"], xm, ym, None, None),          # vsb_remap_linear_u8c3 (the fused path's tap routine)
            "nearest_constant_c1": _warp("k_warp_remapILi1E", inp["c1"], xm, ym, 0, 0),                 # vsb_warp: VSB_INTER_NEAREST, VSB_BORDER_CONSTANT
            "linear_reflect_c3": _warp("k_warp_remapILi3E", inp["c3"], xm, ym, 1, 1)}                   # vsb_warp: VSB_INTER_LINEAR, VSB_BORDER_REFLECT


def remap_warp_linear_constant(inp):
    """the second LINEAR / CONSTANT implementation of the product (k_warp_remap<3>, behind vsb_warp) on the same case"""
    return _warp("k_warp_remapILi3E", inp["c3"], inp["xmap"], inp["ymap"], 1, 0)


def gain(inp):
    img = inp["img"]
    rows, wbytes = img.shape[0], img.shape[1] * 3
    out = {}
    for g in inp["gains"]:
        buf = img.copy()
        mem = P.Memory()
        a = mem.add(buf)
        _run("k_gain_u8", wbytes, rows, [_q(a), _i(wbytes), _i(rows), _q(wbytes), _f(g)], mem)
        out[f"{float(g):.2f}"] = buf
    return out


def copy_make_border(inp):
    out = {}
    for j, (h, w, t, b, l, r) in enumerate(PC.BORDER_SHAPES):
        src = inp[f"img{j}"]
        H, W = h + t + b, w + l + r
        dst = np.zeros((H, W, 3), np.int16)
        mem = P.Memory()
        a_s, a_d = mem.add(src), mem.add(dst)
        _run("k_border_reflect_u8c3_s16c3", W, H, [_q(a_s), _i(w), _i(h), _q(w * 3), _i(t), _i(l), _q(a_d), _i(W), _i(H), _q(W * 6)], mem)
        out[f"img{j}"] = dst.astype(np.uint8)   # the product's kernel also does the convertTo(CV_16S) that follows: exact, undone here
    return out


def cuda_resize(inp):
    """k_resize_linear_u8<CN> (vsb_calib.cu, behind vsb_resize_linear_u8 and the device calibration): the host wrapper passes
    kx = (float)(1.0 / fx), ky = (float)(1.0 / fy) like cuda::resize does"""
    K = kernels("vsb_calib")
    m, f = inp["mask"], inp["frame"]
    s = 0.3
    out = {}
    for key, frag, src, dw, dh, fx, fy in (("mask_up", "k_resize_linear_u8ILi1E", m, 61, 40, 61 / m.shape[1], 40 / m.shape[0]),
                                            ("frame_down", "k_resize_linear_u8ILi3E", f, int(np.rint(f.shape[1] * s)), int(np.rint(f.shape[0] * s)), s, s)):
        sh, sw = src.shape[:2]
        cn = 1 if src.ndim == 2 else 3
        dst = np.zeros((dh, dw) + src.shape[2:], np.uint8)
        mem = P.Memory()
        a_s, a_d = mem.add(src), mem.add(dst)
        P.launch(P.find(K, frag), _g(dw, dh), (32, 8), [_q(a_s), _i(sw), _i(sh), _q(sw * cn), _q(a_d), _i(dw), _i(dh), _q(dw * cn), _f(np.float32(1.0 / fx)), _f(np.float32(1.0 / fy))], mem)
        out[key] = dst
    return out


# ---- the fused hot kernels K1 / K2 (vsb_pipeline.cu) with their tap-table builders, driven the way launch_front / vsb_set_mesh
#      drive them: k_build_taps1 -> k_remap_stage1_tab<LANES = true> (remap #1 + gain -> zero-framed P), k_build_taps2 ->
#      k_remap_stage2_tab (remap #2 + REFLECT border -> planar u8 G0).  Parameter blocks mirrored with ctypes (natural alignment,
#      like nvcc lays the structs out); the sizes are checked against the PTX.
import ctypes as C

MAXV, MAX_BATCH = 16, 16


class _TapTable(C.Structure):
    _fields_ = [("off", C.c_uint64), ("w", C.c_uint64), ("plane", C.c_size_t), ("tab_pitch", C.c_int)]


class _Stage1TabView(C.Structure):
    _fields_ = [("tab", _TapTable), ("xmap", C.c_uint64), ("ymap", C.c_uint64), ("P", C.c_uint64), ("map_pitch", C.c_size_t), ("p_pitch", C.c_size_t),
                ("p_frame_stride", C.c_size_t), ("w", C.c_int), ("h", C.c_int), ("src_w", C.c_int), ("src_h", C.c_int), ("gain", C.c_float)]


class _Stage1TabParams(C.Structure):
    _fields_ = [("tiles", C.c_uint64), ("v", _Stage1TabView * MAXV), ("src", C.c_uint64 * (MAX_BATCH * MAXV)), ("src_pitch", C.c_uint),
                ("v0", C.c_int), ("n_views", C.c_int), ("n_frames", C.c_int), ("f0", C.c_int)]


class _Stage2TabView(C.Structure):
    _fields_ = [("tab", _TapTable), ("Pbase", C.c_uint64), ("G0", C.c_uint64), ("p_frame_stride", C.c_size_t), ("g0_frame_stride", C.c_size_t),
                ("p_pitch", C.c_uint), ("bw", C.c_int), ("bh", C.c_int)]


class _Stage2TabParams(C.Structure):
    _fields_ = [("tiles", C.c_uint64), ("v", _Stage2TabView * MAXV), ("n_frames", C.c_int), ("f0", C.c_int)]


def _align_up(v, a):
    return (v + a - 1) // a * a


def fused_remap(inp, lanes=True):
    K = kernels("vsb_pipeline")
    src, xmap, ymap, xmesh, ymesh = inp["c3"], inp["xmap"], inp["ymap"], inp["xmesh"], inp["ymesh"]
    sh, sw = src.shape[:2]
    h, w = xmap.shape
    t, b, l, r = PC.FUSED_BORDER
    bw, bh = w + l + r, h + t + b
    src_pitch = _align_up(sw * 3, 4)                            # launch_front takes the table-driven kernels for 4-byte aligned rows
    pitched = np.zeros((sh - 1) * src_pitch + sw * 3, np.uint8)   # exactly the bytes a caller owns: any read past them is an error here
    for y in range(sh):
        pitched[y * src_pitch:y * src_pitch + sw * 3] = src[y].reshape(-1)
    src = pitched
    # P with its zero frame (vsb_pipeline.cu: p_pitch / p_frame_stride / p_origin of vsb_init_view)
    p_pitch = _align_up((w + 5) * 3 + 8, 16)
    p_frame_stride = _align_up(p_pitch * (h + 2) + 16, 16)
    p_origin = p_pitch + 12
    P_alloc = np.zeros(p_frame_stride, np.uint8)
    t1_pitch, t2_pitch = _align_up(w, 4), _align_up(bw, 4)
    t1_off, t1_w = np.zeros(t1_pitch * h, np.int32), np.zeros(4 * t1_pitch * h, np.float32)
    t2_off, t2_w = np.zeros(t2_pitch * bh, np.int32), np.zeros(4 * t2_pitch * bh, np.float32)
    flag = np.zeros(1, np.int32)
    G0 = np.full(3 * bw * bh, 0xEE, np.uint8)
    tiles1 = np.array([0 | (tx << 8) | (ty << 20) for ty in range((h + 7) // 8) for tx in range((w + 127) // 128)], np.uint32)
    tiles2 = np.array([0 | (tx << 8) | (ty << 20) for ty in range((bh + 7) // 8) for tx in range((bw + 127) // 128)], np.uint32)
    mem = P.Memory()
    a = {k: mem.add(v) for k, v in dict(src=src, xmap=xmap, ymap=ymap, xmesh=xmesh, ymesh=ymesh, P=P_alloc, t1o=t1_off, t1w=t1_w, t2o=t2_off, t2w=t2_w,
                                        flag=flag, G0=G0, tiles1=tiles1, tiles2=tiles2).items()}
    # ---- table of remap #1 (build_taps1) and the frame kernel (launch_front)
    P.launch(P.find(K, "k_build_taps1"), _g(t1_pitch, h), (32, 8),
             [_q(a["xmap"]), _q(a["ymap"]), _q(w * 4), _i(w), _i(h), _i(sw), _i(sh), struct.pack("<I", src_pitch), _q(a["t1o"]), _q(a["t1w"]), _q(t1_pitch * h), _i(t1_pitch), _i(0), _q(a["flag"])], mem)
    p1 = _Stage1TabParams()
    p1.tiles = a["tiles1"]
    v = p1.v[0]
    v.tab.off, v.tab.w, v.tab.plane, v.tab.tab_pitch = a["t1o"], a["t1w"], t1_pitch * h, t1_pitch
    v.xmap, v.ymap, v.P, v.map_pitch, v.p_pitch, v.p_frame_stride = a["xmap"], a["ymap"], a["P"] + p_origin, w * 4, p_pitch, p_frame_stride
    v.w, v.h, v.src_w, v.src_h, v.gain = w, h, sw, sh, PC.FUSED_GAIN
    p1.src[0] = a["src"]
    p1.src_pitch, p1.v0, p1.n_views, p1.n_frames, p1.f0 = src_pitch, 0, 1, 1, 0
    k1 = P.find(K, "k_remap_stage1_tabILb1E" if lanes else "k_remap_stage1_tabILb0E")
    assert C.sizeof(p1) == k1.params[0][1], (C.sizeof(p1), k1.params[0][1])
    P.launch(k1, (len(tiles1),), (32, 8), [bytes(p1)], mem)
    slow_entries = int(np.count_nonzero(t1_off < 0))
    p_roi = np.stack([P_alloc[p_origin + y * p_pitch: p_origin + y * p_pitch + w * 3].reshape(w, 3) for y in range(h)])
    frame_untouched = int(P_alloc.sum()) == int(p_roi.sum())      # K1 wrote the ROI only: the zero frame is intact
    # ---- table of remap #2 (vsb_set_mesh) and the frame kernel
    P.launch(P.find(K, "k_build_taps2"), _g(t2_pitch, bh), (32, 8),
             [_q(a["xmesh"]), _q(a["ymesh"]), _q(w * 4), _i(w), _i(h), _i(bw), _i(bh), _i(t), _i(l), struct.pack("<I", p_pitch), struct.pack("<I", p_origin),
              _q(a["t2o"]), _q(a["t2w"]), _q(t2_pitch * bh), _i(t2_pitch), _q(a["flag"])], mem)
    p2 = _Stage2TabParams()
    p2.tiles = a["tiles2"]
    v2 = p2.v[0]
    v2.tab.off, v2.tab.w, v2.tab.plane, v2.tab.tab_pitch = a["t2o"], a["t2w"], t2_pitch * bh, t2_pitch
    v2.Pbase, v2.G0, v2.p_frame_stride, v2.g0_frame_stride, v2.p_pitch, v2.bw, v2.bh = a["P"], a["G0"], p_frame_stride, 3 * bw * bh, p_pitch, bw, bh
    p2.n_frames, p2.f0 = 1, 0
    k2 = P.find(K, "k_remap_stage2_tab")
    assert C.sizeof(p2) == k2.params[0][1], (C.sizeof(p2), k2.params[0][1])
    P.launch(k2, (len(tiles2),), (32, 8), [bytes(p2)], mem)
    g0 = np.ascontiguousarray(G0.reshape(3, bh, bw).transpose(1, 2, 0))   # planar -> interleaved for the comparison
    return {"p": p_roi, "g0": g0}, {"slow_entries": slow_entries, "zero_frame_intact": frame_untouched, "unsafe_flag": int(flag[0])}


def build_warp_maps(og):
    out = {}
    k = P.find(kernels(), "k_build_maps")
    for name, proj, view, dx, dy in PC.MAP_PATCHES:
        K, R, scale, tl_x, tl_y, w, h = PC.map_patch_args(og, proj, view, dx, dy)
        k_rinv, _, _ = og.projector(K, R)
        xm, ym = np.zeros((h, w), np.float32), np.zeros((h, w), np.float32)
        mem = P.Memory()
        ax, ay = mem.add(xm), mem.add(ym)
        P.launch(k, _g(w, h), (32, 8), [_i(proj), _f(scale), np.asarray(k_rinv, np.float32).tobytes(), _i(tl_x), _i(tl_y), _i(w), _i(h), _q(ax), _q(ay), _q(w * 4)], mem)
        out[f"{name}_x"], out[f"{name}_y"] = xm, ym
    return out


RUNNERS = {"app_resize": app_resize, "multiband_blend": multiband_blend, "pyr_down": pyr_down, "pyr_up": pyr_up, "remap": remap,
           "gain": gain, "copy_make_border": copy_make_border, "cuda_resize": cuda_resize}
