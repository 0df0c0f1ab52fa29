"""CPU: oracle-G against the reference's OWN CUDA kernels.

The reference's .cu files of the path compile to PTX with this image's nvcc (oracle/ref_ptx.mk: unmodified sources, from where they
lie under /root/reference; no GPU involved).  oracle/ptx_interp.py executes that PTX on the CPU with exact binary32 arithmetic, so
the kernels' results -- including what the compiler did to their expressions: which products are fused into an fma -- are known
bit for bit: cuda::remap (LINEAR / NEAREST, BORDER_CONSTANT / BORDER_REFLECT), cuda::pyrDown<short3> / <float>, cuda::pyrUp<short3>,
addSrcWeightKernel32F, normalizeUsingWeightKernel32F and the application's `resize` kernel.  Their outputs on the seeded cases of
oracle/ptx_cases.py are committed as tests/golden/reference_ptx.npz (generator: tests/golden/make_ptx_golden.py).

* test_oracle_equals_reference_kernels: oracle-G == those outputs, everywhere (no reference, no nvcc needed).
* test_live_*: where oracle/_ref/ptx exists the kernels are executed again (the fixture is current); where nvcc exists, the
  PRODUCT's primitive kernels (vsb_primitives.cu) are compiled with the product's flags and executed the same way: product
  device code == reference device code, bit for bit, with no GPU involved."""
import os
import shutil
import struct
import subprocess

import numpy as np
import pytest

from oracle import ptx_cases as PC
from oracle import ptx_interp as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_ptx.npz"))


def _same(a, b):
    return a.shape == b.shape and a.dtype == b.dtype and np.array_equal(a.view(np.uint8), b.view(np.uint8))


@pytest.mark.parametrize("case", sorted(PC.CASES))
def test_oracle_equals_reference_kernels(og, gold, case):
    _, _, run_oracle, _ = PC.CASES[case]
    got = run_oracle(og, PC.inputs_of(case))
    for key, arr in got.items():
        want = gold[f"{case}__{key}"]
        assert _same(arr, want), f"{case}/{key}: {int(np.count_nonzero(arr != want))} of {arr.size} samples differ from the reference kernel"
        assert np.count_nonzero(want) > want.size // 4


@pytest.mark.parametrize("case", sorted(PC.CASES))
def test_live_reference_ptx_matches_fixture(gold, case):
    if not PC.available():
        pytest.skip("oracle/_ref/ptx not built (needs /root/reference and nvcc: make -C oracle -f ref_ptx.mk)")
    _, run_ptx, _, _ = PC.CASES[case]
    for key, arr in run_ptx(PC.inputs_of(case)).items():
        assert _same(arr, gold[f"{case}__{key}"]), f"{case}/{key}"


@pytest.mark.parametrize("case", sorted(PC.CASES))
def test_live_product_primitives_match_reference_kernels(gold, case):
    """The PRODUCT's device code against the REFERENCE's device code, without a GPU and without the oracle in between: the kernels
    behind the device-launcher layer of the C ABI (video-stitcher_b200/csrc/vsb_primitives.cu: vsb_remap_linear_u8c3 -- the fused
    path's own tap routine --, vsb_warp, vsb_gain_u8, vsb_border_reflect_u8c3_to_s16c3, vsb_pyr_down_s16c3, vsb_pyr_up_s16c3,
    vsb_pyr_down_f32, vsb_add_src_weight_32f, vsb_normalize_32f, vsb_custom_resize; vsb_calib.cu: vsb_resize_linear_u8), compiled to PTX with the product's flags
    (--fmad=false) and executed on the CPU, write exactly what the reference's kernels write."""
    from oracle import ptx_product as PP
    if not PP.available():
        pytest.skip("nvcc not found")
    if case == "fused_remap":
        pytest.skip("covered by test_live_product_fused_remap_kernels_match_reference_chain")
    inp = PC.inputs_of(case)
    for key, arr in PP.RUNNERS[case](inp).items():
        assert _same(arr, gold[f"{case}__{key}"]), f"{case}/{key}"
    if case == "remap":   # vsb_warp's LINEAR / CONSTANT form (k_warp_remap<3>) is a second implementation of the same kernel
        assert _same(PP.remap_warp_linear_constant(inp), gold["remap__linear_constant_c3"])


@pytest.mark.parametrize("lanes", [True, False])
def test_live_product_fused_remap_kernels_match_reference_chain(gold, lanes):
    """The product's two hot remap kernels as they ship -- k_build_taps1 + k_remap_stage1_tab<LANES> (K1: remap #1 + gain through the
    tap table, denormal-scaled fma chain, lane-interleaved pixels transposed through shared memory, TAP_SLOW entries through the
    edge routine) and k_build_taps2 + k_remap_stage2_tab (K2: remap #2 with the REFLECT border resolved in the table, zero-framed P,
    planar u8 out) -- compiled from vsb_pipeline.cu and executed on the CPU, against the REFERENCE's four kernels for the same
    stretch of the path (cuda::remap -> convertTo(gain) -> cuda::remap -> copyMakeBorder(REFLECT), 360_stitcher/timed.cpp:84-100,
    S/src/blenders.cpp:711), executed the same way: identical bytes.  The camera frame is exactly as large as a caller's buffer,
    so a window load past its end would be caught (the interpreter bounds-checks every access)."""
    from oracle import ptx_product as PP
    if not PP.available():
        pytest.skip("nvcc not found")
    got, info = PP.fused_remap(PC.inputs_of("fused_remap"), lanes)
    assert _same(got["p"], gold["fused_remap__p"]), "K1: gain(remap #1)"
    assert _same(got["g0"], gold["fused_remap__g0"]), "K2: border(remap #2)"
    assert info["slow_entries"] >= 2 and info["zero_frame_intact"] and info["unsafe_flag"] == 0, info


def test_map_builder_product_equals_reference_kernel_and_oracle_is_close(og, gold):
    """buildWarpMapsKernel<SphericalMapper / CylindricalMapper> (S/src/cuda/build_warp_maps.cu:88-152) evaluates CUDA's sinf / cosf;
    its PTX carries them inline, so the interpreter reproduces the DEVICE's values.  The product's k_build_maps, compiled and executed
    the same way, equals the reference kernel bit for bit (same library code, same contraction of k0*x + k1*y + k2*z); oracle-G and
    the product's host twin use the host libm and stay within 1e-4 px of the kernel (the GPU tests allow 2e-3 px)."""
    from oracle import ptx_product as PP
    want = {k[len("build_warp_maps__"):]: gold[k] for k in gold.files if k.startswith("build_warp_maps__")}
    assert len(want) == 2 * len(PC.MAP_PATCHES)
    if PC.available():
        for key, arr in PC.build_warp_maps_ptx(og).items():
            assert _same(arr, want[key]), key
    if PP.available():
        for key, arr in PP.build_warp_maps(og).items():
            assert _same(arr, want[key]), f"k_build_maps {key}"
    for name, proj, view, dx, dy in PC.MAP_PATCHES:
        K, R, scale, tl_x, tl_y, w, h = PC.map_patch_args(og, proj, view, dx, dy)
        ox, oy = og.build_maps(proj, scale, K, R, tl_x, tl_y, w, h)
        assert np.abs(ox - want[f"{name}_x"]).max() <= 1e-4 and np.abs(oy - want[f"{name}_y"]).max() <= 1e-4
        assert np.count_nonzero(ox.view(np.uint32) == want[f"{name}_x"].view(np.uint32)) > ox.size // 2   # and most samples are identical


def test_host_map_builder_equals_oracle(og, vsb):
    """vsb_host_build_maps = the maps vsb_calibrate_rig builds on the host (pure host code, no device): bit-identical to oracle-G's
    (same arithmetic, same libm), which is what makes everything downstream of the maps comparable bit for bit on the GPU."""
    import ctypes as C
    L = vsb.lib()
    fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))
    for proj in (0, 1):
        for view in (0, 1, 3):
            K, R = og.rig_camera(6, view, 640, 360, 90.0)
            scale = np.float32(1280 / (2 * 3.1415926535897932384626))
            roi = og.warp_roi(proj, scale, K, R, 640, 360)
            w, h = roi[2], roi[3]
            xm, ym = np.zeros((h, w), np.float32), np.zeros((h, w), np.float32)
            Kf, Rf = np.ascontiguousarray(K, np.float32).reshape(9), np.ascontiguousarray(R, np.float32).reshape(9)
            assert L.vsb_host_build_maps(proj, C.c_float(scale), fp(Kf), fp(Rf), roi[0], roi[1], w, h, fp(xm), fp(ym)) == 0
            ox, oy = og.build_maps(proj, scale, K, R, *roi)
            assert _same(xm, ox) and _same(ym, oy), (proj, view)


def test_interpreter_fma_is_correctly_rounded():
    """The one operation the interpreter does not delegate to IEEE hardware: fma.rn.f32 through exact rational arithmetic."""
    rng = np.random.default_rng(0)
    a = rng.standard_normal(2000).astype(np.float32) * np.float32(37.0)
    b = rng.standard_normal(2000).astype(np.float32)
    for x, y in zip(a[:500], b[:500]):  # c = 0: fma(a, b, 0) is the correctly rounded product, which the hardware multiply also gives
        assert P.fma_f32(x, y, np.float32(0)) == np.float32(x * y)
    # a case where rounding the product first and the sum second differs from the fused result
    x, y, c = np.float32(1 + 2.0 ** -12), np.float32(1 + 2.0 ** -12), np.float32(-1.0)
    assert float(P.fma_f32(x, y, c)) == 2.0 ** -11 + 2.0 ** -24
    assert float(np.float32(np.float32(x * y) + c)) == 2.0 ** -11
    # ties to even, subnormal results, and exact cancellation
    assert float(P.round_fraction_to_f32(__import__("fractions").Fraction(2 ** 24 + 1))) == 2.0 ** 24
    assert float(P.round_fraction_to_f32(__import__("fractions").Fraction(2 ** 24 + 3))) == 2.0 ** 24 + 4
    assert float(P.fma_f32(np.float32(2.0 ** -100), np.float32(2.0 ** -40), np.float32(0))) == 2.0 ** -140
    assert P.fma_f32(np.float32(3.0), np.float32(2.0), np.float32(-6.0)) == 0 and not np.signbit(P.fma_f32(np.float32(3.0), np.float32(2.0), np.float32(-6.0)))


def test_live_reference_pyramid_kernels_on_degenerate_shapes(og):
    """The reference's pyrDown<short3> / pyrUp<short3> kernels (interpreted PTX) on planes where every sample is a border sample --
    one row, one column, 2 x 2 ... (the top levels of a deep pyramid) -- against oracle-G's border index rules: bit for bit."""
    if not PC.available():
        pytest.skip("oracle/_ref/ptx not built (needs /root/reference and nvcc: make -C oracle -f ref_ptx.mk)")
    rng = np.random.default_rng(41)
    for shape in ((1, 1), (1, 7), (7, 1), (2, 2), (3, 5), (5, 3), (2, 9), (9, 2), (3, 3), (1, 2), (2, 1), (4, 33)):
        a = rng.integers(-300, 600, shape + (3,)).astype(np.int16)
        assert _same(PC._pyr_down("pyrDownI6short3", a, 6), og.pyr_down_s16(a)), f"pyrDown {shape}"
        assert _same(PC.pyr_up_ptx({"s16": a})["s16"], og.pyr_up_s16(a)), f"pyrUp {shape}"
    for shape in ((1, 1), (1, 5), (6, 1), (2, 2), (3, 4)):
        w = rng.random(shape).astype(np.float32)
        assert _same(PC._pyr_down("pyrDownIfNS", w, 4), og.pyr_down_f32(w)), f"pyrDown<float> {shape}"


def test_live_product_pyramid_primitives_on_degenerate_shapes():
    """The PRODUCT's pyramid primitives (vsb_pyr_down_s16c3 / vsb_pyr_up_s16c3 / vsb_pyr_down_f32, compiled to PTX with the product's
    flags) against the REFERENCE's kernels on the same degenerate planes, both interpreted: identical bytes."""
    from oracle import ptx_product as PP
    if not PC.available() or not PP.available():
        pytest.skip("needs oracle/_ref/ptx (reference kernels) and nvcc (product PTX)")
    rng = np.random.default_rng(43)
    for shape in ((1, 1), (1, 7), (7, 1), (2, 2), (3, 5), (5, 3), (2, 9), (1, 2), (2, 1)):
        a = rng.integers(-300, 600, shape + (3,)).astype(np.int16)
        w = rng.random(shape).astype(np.float32)
        mine = PP.pyr_down({"s16": a, "f32": w})
        assert _same(mine["s16"], PC._pyr_down("pyrDownI6short3", a, 6)), f"pyrDown {shape}"
        assert _same(mine["f32"], PC._pyr_down("pyrDownIfNS", w, 4)), f"pyrDown<float> {shape}"
        assert _same(PP.pyr_up({"s16": a})["s16"], PC.pyr_up_ptx({"s16": a})["s16"]), f"pyrUp {shape}"
