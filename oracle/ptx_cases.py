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
    return all(os.path.exists(os.path.join(PTX_DIR, f + ".ptx")) for f in ("app_resize", "multiband_blend", "pyr_down", "pyr_up", "remap", "resize", "gpu_mat", "copy_make_border", "build_warp_maps"))


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


# ---- sources/modules/core/src/cuda/gpu_mat.cu:488-512 (GpuMat::convertTo(type, alpha): the gain of 360_stitcher/timed.cpp:94) --------------
# cudev gridTransformUnary_ -> grid_transform_detail::transformSimple<GlobPtr<uchar>, uchar, Convertor<uchar, uchar, float>, WithOutMask>
# over the image reshaped to one channel, 32 x 8 threads (the 4-samples-per-thread "smart" form computes the same values)
def gain_inputs(rng):
    return {"img": rng.integers(0, 256, (13, 41, 3), dtype=np.uint8), "gains": np.array([0.97, 1.0, 1.03, 1.7], np.float32)}


def gain_ptx(inp):
    K = kernels("gpu_mat")
    k = [v for n, v in K.items() if "transformSimple" in n and "ConvertorIhhfE" in n]
    assert len(k) == 1
    img = inp["img"]
    rows, cols = img.shape[0], img.shape[1] * 3
    out = {}
    for g in inp["gains"]:
        dst = np.zeros_like(img)
        mem = P.Memory()
        a_s, a_d = mem.add(img), mem.add(dst)
        P.launch(k[0], (_grid(cols, 32), _grid(rows, 8)), (32, 8),
                 [P.ptr_step(a_s, cols), P.ptr_step(a_d, cols), struct.pack("<ff", np.float32(g), 0.0), b"\0", P.i32(rows), P.i32(cols)], mem)
        out[f"{float(g):.2f}"] = dst
    return out


def gain_oracle(og, inp):
    return {f"{float(g):.2f}": og.gain_u8(inp["img"], np.float32(g)) for g in inp["gains"]}


# ---- sources/modules/cudaarithm/src/cuda/copy_make_border.cu:56-115 (BORDER_REFLECT, CV_8UC3: S/src/blenders.cpp:711) -----------------------
# cudev gridCopy -> grid_copy_detail::copy<RemapPtr1<BrdBase<BrdReflect, GlobPtr<uchar3>>, ShiftMap>, uchar3, WithOutMask>, 32 x 8 threads
BORDER_SHAPES = ((20, 31, 5, 7, 9, 4), (6, 8, 3, 2, 7, 5), (5, 4, 7, 9, 9, 6))   # (h, w, top, bottom, left, right); the last: borders wider than the image


def border_inputs(rng):
    return {f"img{j}": rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for j, (h, w, *_) in enumerate(BORDER_SHAPES)}


def border_ptx(inp):
    k = P.find(kernels("copy_make_border"), "BrdBaseINS0_10BrdReflectENS0_7GlobPtrI6uchar3")
    out = {}
    for j, (h, w, t, b, l, r) in enumerate(BORDER_SHAPES):
        src = inp[f"img{j}"]
        H, W = h + t + b, w + l + r
        dst = np.zeros((H, W, 3), np.uint8)
        mem = P.Memory()
        a_s, a_d = mem.add(src), mem.add(dst)
        # RemapPtr1{BrdBase{GlobPtr{data, step}; int rows, cols}; ShiftMap{int top, left}}
        P.launch(k, (_grid(W, 32), _grid(H, 8)), (32, 8), [struct.pack("<QQiiii", a_s, w * 3, h, w, t, l), P.ptr_step(a_d, W * 3), b"\0", P.i32(H), P.i32(W)], mem)
        out[f"img{j}"] = dst
    return out


def border_oracle(og, inp):
    # the oracle's border function also does the convertTo(CV_16S) that follows (S/src/blenders.cpp:713): exact, undone here
    return {f"img{j}": og.border_reflect_u8c3_to_s16(inp[f"img{j}"], t, b, l, r).astype(np.uint8) for j, (h, w, t, b, l, r) in enumerate(BORDER_SHAPES)}


# ---- sources/modules/cudawarping/src/cuda/resize.cu:71-106, 211-219 (resize_linear, 32 x 8 threads; host side src/resize.cpp:76-105) ------
def cuda_resize_inputs(rng):
    return {"mask": (rng.random((17, 23)) > 0.5).astype(np.uint8) * 255, "frame": rng.integers(0, 256, (54, 96, 3), dtype=np.uint8)}


def _cuda_resize(frag, src, dw, dh, fx, fy):
    k = P.find(kernels("resize"), frag)
    sh, sw = src.shape[:2]
    cn = 1 if src.ndim == 2 else 3
    out = np.zeros((dh, dw) + src.shape[2:], np.uint8)
    mem = P.Memory()
    a_s, a_d = mem.add(src), mem.add(out)
    # cuda::resize passes static_cast<float>(1.0 / fy), static_cast<float>(1.0 / fx) (src/resize.cpp:103)
    P.launch(k, (_grid(dw, 32), _grid(dh, 8)), (32, 8), [_ptr_step_sz(a_s, sw * cn, sw, sh), _ptr_step_sz(a_d, dw * cn, dw, dh),
                                                        struct.pack("<f", np.float32(1.0 / fy)), struct.pack("<f", np.float32(1.0 / fx))], mem)
    return out


def cuda_resize_ptx(inp):
    m, f = inp["mask"], inp["frame"]
    s = 0.3
    return {"mask_up": _cuda_resize("resize_linearIhEE", m, 61, 40, 61 / m.shape[1], 40 / m.shape[0]),                    # A/calibration.cpp:236 (dsize given)
            "frame_down": _cuda_resize("resize_linearI6uchar3EE", f, int(np.rint(f.shape[1] * s)), int(np.rint(f.shape[0] * s)), s, s)}  # A/calibration.cpp:95


def cuda_resize_oracle(og, inp):
    m, f = inp["mask"], inp["frame"]
    s = 0.3
    return {"mask_up": og.resize_linear_u8c1(m, 61, 40),
            "frame_down": og.cuda_resize_linear_u8(f, int(np.rint(f.shape[1] * s)), int(np.rint(f.shape[0] * s)), s, s)}


# ---- the remap half of stitch_online + the head of feed_online as the reference runs it (360_stitcher/timed.cpp:84-100,
#      S/src/blenders.cpp:711): cuda::remap (projection maps) -> convertTo(gain) -> cuda::remap (mesh maps) -> copyMakeBorder(REFLECT),
#      four kernels in a row.  The product does this in two fused, table-driven kernels (K1 / K2, oracle/ptx_product.py).
FUSED_GAIN = 1.03
FUSED_BORDER = (5, 5, 9, 11)    # top, bottom, left, right


def fused_inputs(rng):
    inp = remap_inputs(rng)
    dh, dw = inp["xmap"].shape
    yy, xx = np.mgrid[0:dh, 0:dw].astype(np.float64)
    # mesh maps: identity plus a smooth displacement of a few pixels (what the CPW mesh produces), reaching past every edge
    inp["xmesh"] = (xx + 3.0 * np.sin(yy / 7.0) + 2.0 * np.cos(xx / 5.0) - 1.2 + 0.37 * rng.random((dh, dw))).astype(np.float32)
    inp["ymesh"] = (yy + 2.5 * np.cos(xx / 6.0) - 1.5 * np.sin(yy / 4.0) + 0.8 + 0.41 * rng.random((dh, dw))).astype(np.float32)
    # a camera frame whose rows are not a multiple of 4 bytes (51 px = 153 B, rows 156 B apart in the product's run): the word loads of
    # windows at the very end of the buffer would leave it -- the product's table marks those entries and takes its coordinate-driven
    # edge routine for them
    inp["c3"] = np.ascontiguousarray(inp["c3"][:, :51])
    sh, sw = inp["c3"].shape[:2]
    inp["xmap"][1, 1] = sw - 0.5
    inp["xmap"][3, 3], inp["ymap"][3, 3] = sw - 1.5, sh - 1.5
    inp["xmap"][3, 4], inp["ymap"][3, 4] = sw - 1.25, sh - 1.0
    inp["xmap"][3, 5], inp["ymap"][3, 5] = sw - 2.75, sh - 1.75
    return {k: v for k, v in inp.items() if k != "c1"}


def fused_ptx(inp):
    sh, sw = inp["c3"].shape[:2]
    dh, dw = inp["xmap"].shape
    p1 = _remap("LinearFilterINS1_12BorderReaderINS0_7PtrStepI6uchar3EENS1_11BrdConstantI6float3", inp["c3"], inp["xmap"], inp["ymap"],
                struct.pack("<iifff", sh, sw, 0.0, 0.0, 0.0))
    k = [v for n, v in kernels("gpu_mat").items() if "transformSimple" in n and "ConvertorIhhfE" in n][0]
    p2 = np.zeros_like(p1)
    mem = P.Memory()
    a_s, a_d = mem.add(p1), mem.add(p2)
    P.launch(k, (_grid(dw * 3, 32), _grid(dh, 8)), (32, 8), [P.ptr_step(a_s, dw * 3), P.ptr_step(a_d, dw * 3), struct.pack("<ff", np.float32(FUSED_GAIN), 0.0), b"\0", P.i32(dh), P.i32(dw * 3)], mem)
    q = _remap("LinearFilterINS1_12BorderReaderINS0_7PtrStepI6uchar3EENS1_11BrdConstantI6float3", p2, inp["xmesh"], inp["ymesh"],
               struct.pack("<iifff", dh, dw, 0.0, 0.0, 0.0))
    t, b, l, r = FUSED_BORDER
    H, W = dh + t + b, dw + l + r
    g0 = np.zeros((H, W, 3), np.uint8)
    kb = P.find(kernels("copy_make_border"), "BrdBaseINS0_10BrdReflectENS0_7GlobPtrI6uchar3")
    mem = P.Memory()
    a_s, a_d = mem.add(q), mem.add(g0)
    P.launch(kb, (_grid(W, 32), _grid(H, 8)), (32, 8), [struct.pack("<QQiiii", a_s, dw * 3, dh, dw, t, l), P.ptr_step(a_d, W * 3), b"\0", P.i32(H), P.i32(W)], mem)
    return {"p": p2, "g0": g0}


def fused_oracle(og, inp):
    p = og.gain_u8(og.remap_linear_u8(inp["c3"], inp["xmap"], inp["ymap"]), np.float32(FUSED_GAIN))
    q = og.remap_linear_u8(p, inp["xmesh"], inp["ymesh"])
    t, b, l, r = FUSED_BORDER
    return {"p": p, "g0": og.border_reflect_u8c3_to_s16(q, t, b, l, r).astype(np.uint8)}


# ---- sources/modules/stitching/src/cuda/build_warp_maps.cu:88-152, 176-215 (32 x 8 threads; k_rinv / scale in __constant__ memory) ---------
# Not in CASES: the kernel evaluates CUDA's sinf / cosf (inlined in its PTX, executed here as compiled), oracle-G the host libm's, so
# oracle and kernel agree to a few 1e-5 px, not bit for bit.  What IS bit-exact against this kernel is the product's k_build_maps
# executed the same way (tests/test_oracle_ptx.py).
MAP_PATCHES = (("spherical", 0, 1, 100, 60), ("spherical_wrapped", 0, 3, 0, 60), ("cylindrical", 1, 1, 100, 60))   # name, projection, view of a 6-view rig, offset into its ROI
MAP_RIG = dict(n_views=6, src_w=640, src_h=360, pano_width=1280, patch_w=40, patch_h=24)


def map_patch_args(og, proj, view, dx, dy):
    K, R = og.rig_camera(MAP_RIG["n_views"], view, MAP_RIG["src_w"], MAP_RIG["src_h"], 90.0)
    scale = np.float32(MAP_RIG["pano_width"] / (2 * 3.1415926535897932384626))
    roi = og.warp_roi(proj, scale, K, R, MAP_RIG["src_w"], MAP_RIG["src_h"])
    return K, R, scale, roi[0] + dx, roi[1] + dy, MAP_RIG["patch_w"], MAP_RIG["patch_h"]


def build_warp_maps_ptx(og):
    out = {}
    K_ = kernels("build_warp_maps")
    for name, proj, view, dx, dy in MAP_PATCHES:
        K, R, scale, tl_x, tl_y, w, h = map_patch_args(og, proj, view, dx, dy)
        k = [v for n, v in K_.items() if ("SphericalMapper" if proj == 0 else "CylindricalMapper") in n][0]
        k_rinv, r_kinv, _ = og.projector(K, R)      # ProjectorBase::setCameraParams: host-side 3 x 3 products (S/src/warpers.cpp:49-79)
        k.module.set_const("ck_rinv", np.asarray(k_rinv, np.float32).tobytes())
        k.module.set_const("cr_kinv", np.asarray(r_kinv, np.float32).tobytes())
        k.module.set_const("cscale", struct.pack("<f", scale))
        xm, ym = np.zeros((h, w), np.float32), np.zeros((h, w), np.float32)
        mem = P.Memory()
        ax, ay = mem.add(xm), mem.add(ym)
        P.launch(k, (_grid(w, 32), _grid(h, 8)), (32, 8), [P.i32(tl_x), P.i32(tl_y), P.i32(w), P.i32(h), P.ptr_step(ax, w * 4), P.ptr_step(ay, w * 4)], mem)
        out[f"{name}_x"], out[f"{name}_y"] = xm, ym
    return out


CASES = {
    "app_resize": (resize_inputs, resize_ptx, resize_oracle, 3),
    "multiband_blend": (blend_inputs, blend_ptx, blend_oracle, 11),
    "pyr_down": (pyr_down_inputs, pyr_down_ptx, pyr_down_oracle, 2),
    "pyr_up": (pyr_up_inputs, pyr_up_ptx, pyr_up_oracle, 4),
    "remap": (remap_inputs, remap_ptx, remap_oracle, 9),
    "gain": (gain_inputs, gain_ptx, gain_oracle, 1),
    "copy_make_border": (border_inputs, border_ptx, border_oracle, 6),
    "cuda_resize": (cuda_resize_inputs, cuda_resize_ptx, cuda_resize_oracle, 8),
    "fused_remap": (fused_inputs, fused_ptx, fused_oracle, 9),
}


def inputs_of(name):
    make, _, _, seed = CASES[name]
    return make(np.random.default_rng(seed))
