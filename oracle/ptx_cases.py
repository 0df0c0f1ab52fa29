"""TEST INFRASTRUCTURE: the seeded cases on which the reference's own CUDA kernels (their nvcc-generated PTX, executed on the CPU
by oracle/ptx_interp.py) are compared with oracle-G.  Used by tests/test_oracle_ptx.py (live, where oracle/_ref/ptx exists) and by
tests/golden/make_ptx_golden.py, which stores the kernels' outputs in tests/golden/reference_ptx.npz so that the comparison also
runs where neither /root/reference nor nvcc exist.

Every case: inputs(rng) -> dict of arrays; run_ptx(kernels, inp) -> dict of output arrays (what the reference kernel wrote);
run_oracle(og, inp) -> the same outputs from oracle-G.  Launch geometry and parameter layouts are the reference launchers'
(file:line per case)."""
import os
import struct

import numpy as np

from . import ptx_interp as P

PTX_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "ptx")


def available():
    return all(os.path.exists(os.path.join(PTX_DIR, f)) for f in ("app_resize.ptx", "multiband_blend.ptx", "pyr_down.ptx", "pyr_up.ptx", "remap.ptx"))


_parsed = {}


def kernels(name):
    if name not in _parsed:
        _parsed[name] = P.parse(open(os.path.join(PTX_DIR, name + ".ptx")).read())
    return _parsed[name]


def _ptr_step_sz(addr, step, cols, rows):
    """cv::cuda::PtrStepSz<T>: {T *data; size_t step; int cols; int rows} (core/include/opencv2/core/cuda_types.hpp:91-120)"""
    return struct.pack("<QQii", addr, step, cols, rows)


def _grid(n, b):
    return (n + b - 1) // b


# ---- 360_stitcher/resize.cu:9-45 (custom_resize: 16 x 16 threads, one per output sample) ------------------------------------------
def resize_inputs(rng):
    return {"in": (rng.random((10, 10)) * 900).astype(np.float32), "size": np.array([61, 45], np.int32)}


def resize_ptx(inp):
    k = P.find(kernels("app_resize"), "resize")
    src, (tx, ty) = inp["in"], [int(v) for v in inp["size"]]
    rows, cols = src.shape
    out = np.zeros((ty, tx), np.float32)
    mem = P.Memory()
    a_in, a_out = mem.add(src), mem.add(out)
    P.launch(k, (_grid(tx, 16), _grid(ty, 16)), (16, 16),
             [P.i32(tx), P.i32(ty), P.i32(cols), P.i32(rows), P.ptr_step(a_in, cols * 4), P.ptr_step(a_out, tx * 4)], mem)
    return {"out": out}


def resize_oracle(og, inp):
    return {"out": og.custom_resize(inp["in"], int(inp["size"][0]), int(inp["size"][1]))}


# ---- sources/modules/stitching/src/cuda/multiband_blend.cu:36-59, 85-108 (16 x 16 threads) ---------------------------------------
def blend_inputs(rng):
    """Three views accumulated into a zeroed level, then normalised -- the way feed_online / blend drive the two kernels
    (S/src/blenders.cpp:742-746, 767-783), so the accumulator stays in the range the path can reach (|acc| <= 255 * sum of weights):
    there the quotient fits CV_16S and static_cast<short>(float) is unambiguous (nvcc 12.9 emits cvt.rzi.s32.f32 + a 16-bit store,
    i.e. out of range it would wrap, not saturate)."""
    rows, cols = 21, 37
    out = {}
    for v in range(3):
        w = rng.random((rows, cols)).astype(np.float32)
        w[rng.random((rows, cols)) < 0.4] = 0          # outside the view's mask
        w[rng.random((rows, cols)) < 0.2] = 1          # interior of the mask
        out[f"src{v}"] = rng.integers(-255, 256, (rows, cols, 3)).astype(np.int16)
        out[f"w{v}"] = w
    return out


def blend_ptx(inp):
    K = kernels("multiband_blend")
    ka, kn = P.find(K, "addSrcWeightKernel32F"), P.find(K, "normalizeUsingWeightKernel32F")
    rows, cols = inp["w0"].shape
    dst, dw = np.zeros((rows, cols, 3), np.int16), np.zeros((rows, cols), np.float32)
    mem = P.Memory()
    a_d, a_dw = mem.add(dst), mem.add(dw)
    g = (_grid(cols, 16), _grid(rows, 16))
    for v in range(3):
        a_s, a_w = mem.add(inp[f"src{v}"]), mem.add(inp[f"w{v}"])
        P.launch(ka, g, (16, 16), [P.ptr_step(a_s, cols * 6), P.ptr_step(a_w, cols * 4), P.ptr_step(a_d, cols * 6), P.ptr_step(a_dw, cols * 4),
                                   P.i32(rows), P.i32(cols)], mem)
    acc, accw = dst.copy(), dw.copy()
    P.launch(kn, g, (16, 16), [P.ptr_step(a_dw, cols * 4), P.ptr_step(a_d, cols * 6), P.i32(cols), P.i32(rows)], mem)
    return {"acc": acc, "acc_w": accw, "normalized": dst}


def blend_oracle(og, inp):
    rows, cols = inp["w0"].shape
    acc, accw = np.zeros((rows, cols, 3), np.int16), np.zeros((rows, cols), np.float32)
    for v in range(3):
        acc, accw = og.add_src_weight_32f(inp[f"src{v}"], inp[f"w{v}"], acc, accw)
    return {"acc": acc, "acc_w": accw, "normalized": og.normalize_32f(accw, acc)}


# ---- sources/modules/cudawarping/src/cuda/pyr_down.cu:55-188 (256 threads, grid (ceil(src.cols / 256), dst.rows)) ----------------------
def pyr_down_inputs(rng):
    return {"s16": rng.integers(-300, 300, (22, 38, 3)).astype(np.int16), "f32": rng.random((23, 37)).astype(np.float32)}


def _pyr_down(frag, src, elem):
    k = P.find(kernels("pyr_down"), frag)
    h, w = src.shape[:2]
    dh, dw = (h + 1) // 2, (w + 1) // 2
    out = np.zeros((dh, dw) + src.shape[2:], src.dtype)
    mem = P.Memory()
    a_s, a_d = mem.add(src), mem.add(out)
    P.launch(k, (_grid(w, 256), dh), (256,), [_ptr_step_sz(a_s, w * elem, w, h), P.ptr_step(a_d, dw * elem),
                                              struct.pack("<ii", h - 1, w - 1), P.i32(dw)], mem)   # BrdReflect101 {last_row, last_col}
    return out


def pyr_down_ptx(inp):
    return {"s16": _pyr_down("pyrDownI6short3", inp["s16"], 6), "f32": _pyr_down("pyrDownIfNS", inp["f32"], 4)}


def pyr_down_oracle(og, inp):
    return {"s16": og.pyr_down_s16(inp["s16"]), "f32": og.pyr_down_f32(inp["f32"])}


# ---- sources/modules/cudawarping/src/cuda/pyr_up.cu:55-157 (16 x 16 threads over the destination) ---------------------------------------
def pyr_up_inputs(rng):
    return {"s16": rng.integers(-300, 300, (11, 19, 3)).astype(np.int16)}


def pyr_up_ptx(inp):
    k = P.find(kernels("pyr_up"), "pyrUpI6short3")
    src = inp["s16"]
    h, w = src.shape[:2]
    out = np.zeros((2 * h, 2 * w, 3), np.int16)
    mem = P.Memory()
    a_s, a_d = mem.add(src), mem.add(out)
    P.launch(k, (_grid(2 * w, 16), _grid(2 * h, 16)), (16, 16), [_ptr_step_sz(a_s, w * 6, w, h), _ptr_step_sz(a_d, 2 * w * 6, 2 * w, 2 * h)], mem)
    return {"s16": out}


def pyr_up_oracle(og, inp):
    return {"s16": og.pyr_up_s16(inp["s16"])}


# ---- sources/modules/cudawarping/src/cuda/remap.cu:56-107 (32 x 8 threads; the generic BorderReader<PtrStep<T>, B> form) -------------
def remap_inputs(rng):
    sh, sw, dh, dw = 40, 52, 30, 44
    yy, xx = np.mgrid[0:dh, 0:dw].astype(np.float64)
    a = 0.35
    xm = (np.cos(a) * xx * 1.3 - np.sin(a) * yy + 3.3 + rng.random((dh, dw)) - 4).astype(np.float32)
    ym = (np.sin(a) * xx + np.cos(a) * yy * 1.4 - 6.1 + rng.random((dh, dw))).astype(np.float32)
    xm[0, 0] = ym[0, 0] = -1.0                       # the "behind the camera" value of the projection maps
    xm[1, 1], ym[2, 2] = sw - 0.5, sh - 1.0          # straddling the right / bottom edge
    return {"c3": rng.integers(0, 256, (sh, sw, 3), dtype=np.uint8), "c1": rng.integers(0, 256, (sh, sw), dtype=np.uint8), "xmap": xm, "ymap": ym}


def _remap(frag, src, xm, ym, brd):
    k = P.find(kernels("remap"), frag)
    sh, sw = src.shape[:2]
    cn = 1 if src.ndim == 2 else src.shape[2]
    dh, dw = xm.shape
    out = np.zeros((dh, dw) + src.shape[2:], np.uint8)
    mem = P.Memory()
    a_s, a_x, a_y, a_d = mem.add(src), mem.add(xm), mem.add(ym), mem.add(out)
    filt = P.ptr_step(a_s, sw * cn) + brd                  # Filter{BorderReader{PtrStep<T> ptr; B b}}
    filt += b"\0" * (-len(filt) % 8)
    P.launch(k, (_grid(dw, 32), _grid(dh, 8)), (32, 8), [filt, P.ptr_step(a_x, dw * 4), P.ptr_step(a_y, dw * 4), _ptr_step_sz(a_d, dw * cn, dw, dh)], mem)
    return out


def remap_ptx(inp):
    sh, sw = inp["c1"].shape
    return {
        # BrdConstant<float3> {int height, width; float3 val}; BrdConstant<float>; BrdReflect<float3> {int last_row, last_col}
        "linear_constant_c3": _remap("LinearFilterINS1_12BorderReaderINS0_7PtrStepI6uchar3EENS1_11BrdConstantI6float3", inp["c3"], inp["xmap"], inp["ymap"],
                                     struct.pack("<iifff", sh, sw, 0.0, 0.0, 0.0)),
        "nearest_constant_c1": _remap("PointFilterINS1_12BorderReaderINS0_7PtrStepIhEENS1_11BrdConstantIfEE", inp["c1"], inp["xmap"], inp["ymap"],
                                      struct.pack("<iif", sh, sw, 0.0)),
        "linear_reflect_c3": _remap("LinearFilterINS1_12BorderReaderINS0_7PtrStepI6uchar3EENS1_10BrdReflectI6float3", inp["c3"], inp["xmap"], inp["ymap"],
                                    struct.pack("<ii", sh - 1, sw - 1)),
    }


def remap_oracle(og, inp):
    return {"linear_constant_c3": og.remap_linear_u8(inp["c3"], inp["xmap"], inp["ymap"]),
            "nearest_constant_c1": og.remap_nearest_u8c1(inp["c1"], inp["xmap"], inp["ymap"]),
            "linear_reflect_c3": og.remap_u8(inp["c3"], inp["xmap"], inp["ymap"], og.INTER_LINEAR, og.BORDER_REFLECT)}


CASES = {
    "app_resize": (resize_inputs, resize_ptx, resize_oracle, 3),
    "multiband_blend": (blend_inputs, blend_ptx, blend_oracle, 11),
    "pyr_down": (pyr_down_inputs, pyr_down_ptx, pyr_down_oracle, 2),
    "pyr_up": (pyr_up_inputs, pyr_up_ptx, pyr_up_oracle, 4),
    "remap": (remap_inputs, remap_ptx, remap_oracle, 9),
}


def inputs_of(name):
    make, _, _, seed = CASES[name]
    return make(np.random.default_rng(seed))
