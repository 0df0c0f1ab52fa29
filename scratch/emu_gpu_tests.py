"""Dry run of -m gpu TEST CODE on the emulated runtime (no GPU): tests/gpu_util's device helpers are replaced by host buffers of the
stand-in CUDA runtime (oracle/emu), the rigs by tiny ones, and the test functions are called as they are.  Finds mistakes in the
tests themselves (shapes, argument order, view / camera indexing) before they meet hardware.  Minutes per case; run one case per
process in parallel, e.g.  for c in small4 cyl5 nolocal6 wire; do python scratch/emu_gpu_tests.py split $c & done
(round 2: every test function of both files passed this way).

    python scratch/emu_gpu_tests.py scale [case ...]     tests/test_gpu_vsb_compose_scale.py
    python scratch/emu_gpu_tests.py split [case ...]     tests/test_gpu_vsb_wrap_split.py
    python scratch/emu_gpu_tests.py parity [case ...]    the calibration tests of tests/test_gpu_parity.py
"""
import os
import sys
import time
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.emu import runtime as E  # noqa: E402

E.start()
os.environ["VSB200_LIB"] = os.path.join(E.BUILD, "libvsb200_emu.so")


class FakeTensor:
    def __init__(self, a):
        self.buf = E.Buffer(np.ascontiguousarray(a))

    def data_ptr(self):
        return self.buf.ptr

    def __del__(self):   # a stale registration would shadow a later allocation at the same address in the interpreter's bounds check
        try:
            self.buf.release()
        except Exception:
            pass


class FakeCuda:
    uint8, int16, float32 = np.uint8, np.int16, np.float32

    @staticmethod
    def full(shape, value, dtype=None, device=None):
        return FakeTensor(np.full(shape, value, dtype))

    @staticmethod
    def zeros(shape, dtype=None, device=None):
        return FakeTensor(np.zeros(shape, dtype))


fake_torch = types.ModuleType("torch")
fake_torch.full, fake_torch.zeros = FakeCuda.full, FakeCuda.zeros
fake_torch.uint8, fake_torch.int16, fake_torch.float32 = np.uint8, np.int16, np.float32
fake_torch.from_numpy = lambda a: types.SimpleNamespace(cuda=lambda: FakeTensor(a))
fake_torch.cuda = types.SimpleNamespace(current_stream=lambda: types.SimpleNamespace(cuda_stream=0), synchronize=lambda: None, is_available=lambda: True,
                                        set_device=lambda d: None)
sys.modules["torch"] = fake_torch

import tests.gpu_util as U  # noqa: E402
from oracle import oracle as og  # noqa: E402

U.host = lambda t: t.buf.a.copy()
og.build()

which, picks = sys.argv[1], sys.argv[2:]
t0 = time.time()
import atexit  # noqa: E402
atexit.register(lambda: print("emulated runtime error:", E.stats().get("error")))
if which == "scale":
    import tests.test_gpu_vsb_compose_scale as T
    T.CASES.update({
        "small4": dict(n_views=4, src_w=64, src_h=40, pano_width=192, num_bands=3, enable_local=True, compose_scale=0.75),
        "mismatch6": dict(n_views=4, src_w=61, src_h=41, pano_width=192, num_bands=3, enable_local=True, compose_scale=0.8),
        "cyl5_nolocal": dict(n_views=4, src_w=64, src_h=48, pano_width=200, num_bands=3, enable_local=False, projection=1, compose_scale=0.5),
        "near_one": dict(n_views=4, src_w=48, src_h=32, pano_width=192, num_bands=3, enable_local=True, compose_scale=0.95),
    })
    for case in (picks or ["small4", "mismatch6", "cyl5_nolocal", "near_one"]):
        if case in T.CASES:
            T.test_scaled_calibration_and_compose_match_oracle(fake_torch, og, case)
            print("ok: test_scaled_calibration_and_compose_match_oracle", case, round(time.time() - t0), "s", flush=True)
    if not picks or "device" in picks:
        T.test_scaled_device_calibration_and_low_level_route(fake_torch, og)
        print("ok: test_scaled_device_calibration_and_low_level_route", round(time.time() - t0), "s", flush=True)
    T.REF_RIG, T.REF_MEGAPIX = (4, 64, 40, 3), (0.0015, 0.0016)
    for dev_cal in (False, True):
        if not picks or ("megapix_dev" if dev_cal else "megapix") in picks:
            T.test_reference_default_scales_at_1080p(fake_torch, og, dev_cal)
            print("ok: test_reference_default_scales_at_1080p on_device =", dev_cal, round(time.time() - t0), "s", flush=True)
    if not picks or "wire" in picks:
        T.test_scaled_wire_formats_and_host_path(fake_torch, og)
        print("ok: test_scaled_wire_formats_and_host_path", round(time.time() - t0), "s", flush=True)
elif which == "parity":
    # the calibration tests of tests/test_gpu_parity.py (host and device calibration were restructured for compose_scale / the split)
    import tests.test_gpu_parity as T
    T.CASES.update({
        "small4": dict(n_views=4, src_w=48, src_h=32, pano_width=192, num_bands=3, enable_local=True),
        "small6_nolocal": dict(n_views=4, src_w=48, src_h=36, pano_width=192, num_bands=3, enable_local=False),
        "cyl5": dict(n_views=4, src_w=48, src_h=36, pano_width=208, num_bands=3, enable_local=True, projection=1),
    })
    for case in (picks or ["small4", "small6_nolocal", "cyl5"]):
        if case in ("small4", "small6_nolocal", "cyl5"):
            T.test_calibration_products_match_oracle(fake_torch, og, case)
            print("ok: test_calibration_products_match_oracle", case, round(time.time() - t0), "s", flush=True)
    if not picks or "device" in picks:
        T.test_device_calibration_products_and_compose(fake_torch, og, "small6_nolocal")
        print("ok: test_device_calibration_products_and_compose small6_nolocal", round(time.time() - t0), "s", flush=True)
else:
    import tests.test_gpu_vsb_wrap_split as T
    T.CASES.update({
        "small4": dict(n_views=4, src_w=48, src_h=32, pano_width=192, num_bands=3, enable_local=True),
        "cyl5": dict(n_views=4, src_w=48, src_h=36, pano_width=208, num_bands=3, enable_local=True, projection=1),
        "nolocal6": dict(n_views=4, src_w=48, src_h=32, pano_width=192, num_bands=3, enable_local=False),
    })
    for case in (picks or ["small4", "cyl5", "nolocal6"]):
        if case in T.CASES:
            T.test_split_calibration_composes_the_unsplit_panorama(fake_torch, og, case)
            print("ok: test_split_calibration_composes_the_unsplit_panorama", case, round(time.time() - t0), "s", flush=True)
    if not picks or "wire" in picks:
        T.test_split_wire_formats_and_wrong_view_count(fake_torch, og)
        print("ok: test_split_wire_formats_and_wrong_view_count", round(time.time() - t0), "s", flush=True)
    if not picks or "device" in picks:
        T.test_split_on_the_device_calibration_and_gain_refresh(fake_torch, og)
        print("ok: test_split_on_the_device_calibration_and_gain_refresh", round(time.time() - t0), "s", flush=True)
print("done", round(time.time() - t0), "s")
