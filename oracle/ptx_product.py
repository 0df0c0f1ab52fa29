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
