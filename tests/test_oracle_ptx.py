"""CPU: oracle-G against the reference's OWN CUDA kernels.

The reference's .cu files of the path compile to PTX with this image's nvcc (oracle/ref_ptx.mk: unmodified sources, from where they
lie under /root/reference; no GPU involved).  oracle/ptx_interp.py executes that PTX on the CPU with exact binary32 arithmetic, so
the kernels' results -- including what the compiler did to their expressions: which products are fused into an fma -- are known
bit for bit: cuda::remap (LINEAR / NEAREST, BORDER_CONSTANT / BORDER_REFLECT), cuda::pyrDown<short3> / <float>, cuda::pyrUp<short3>,
addSrcWeightKernel32F, normalizeUsingWeightKernel32F and the application's `resize` kernel.  Their outputs on the seeded cases of
oracle/ptx_cases.py are committed as tests/golden/reference_ptx.npz (generator: tests/golden/make_ptx_golden.py).

* test_oracle_equals_reference_kernels: oracle-G == those outputs, everywhere (no reference, no nvcc needed).
* test_live_*: where oracle/_ref/ptx exists the kernels are executed again (the fixture is current); where nvcc exists, the
  product's own device function behind custom_resize is compiled with the product's flags and executed the same way."""
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


def test_live_product_custom_resize_matches_reference_kernel(gold, tmp_path):
    """vsb::custom_resize_at (video-stitcher_b200/csrc/vsb_internal.h: what vsb_custom_resize and the mesh -> map kernels evaluate),
    compiled to PTX with the product's flags (--fmad=false) and executed on the CPU, equals the reference's `resize` kernel
    (360_stitcher/resize.cu:9-27) bit for bit -- a device-vs-reference comparison that needs no GPU."""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not found")
    ptx = tmp_path / "wrap_device.ptx"
    subprocess.check_call([nvcc, "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=compute_100a", "--fmad=false", "-ptx",
                           os.path.join(ROOT, "oracle", "ptx_wrap_device.cu"), "-o", str(ptx)])
    k = P.find(P.parse(ptx.read_text()), "w_custom_resize")
    inp = PC.inputs_of("app_resize")
    src, (tx, ty) = inp["in"], [int(v) for v in inp["size"]]
    rows, cols = src.shape
    out = np.zeros((ty, tx), np.float32)
    mem = P.Memory()
    a_in, a_out = mem.add(src), mem.add(out)
    q = lambda v: struct.pack("<Q", v)
    P.launch(k, ((tx + 15) // 16, (ty + 15) // 16), (16, 16), [P.i32(tx), P.i32(ty), P.i32(cols), P.i32(rows), q(a_in), q(cols * 4), q(a_out), q(tx * 4)], mem)
    assert _same(out, gold["app_resize__out"])


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
