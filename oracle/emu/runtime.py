"""TEST INFRASTRUCTURE: runs the product library on a machine without a GPU.

libvsb200_emu.so = the product's own object files linked against oracle/emu/fake_cudart.c; every kernel launch of the product's
host code arrives here and is executed by oracle/ptx_interp.py on the kernel's PTX (the same .cu files compiled with the product's
flags).  What this exercises is everything except the hardware: calibration, plans, tile lists, launch sequences, every kernel as
compiled.  (k_down2 runs in its plain-load form: there is no tensor-map encoder here, exactly as on a driver without one.)"""
import ctypes as C
import os
import shutil
import subprocess
import sys
import time

import numpy as np

from .. import ptx_interp as P

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
BUILD = os.path.join(ROOT, "oracle", "_build")
_NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
UNITS = ("vsb_common", "vsb_primitives", "vsb_pipeline", "vsb_calib")
_state = {}


def available():
    return os.path.exists(_NVCC)


def _build():
    subprocess.check_call(["make", "-s", "-j4", "-C", os.path.join(ROOT, "video-stitcher_b200", "csrc")])
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle", "emu")])
    kernels = {}
    for unit in UNITS:
        src = os.path.join(ROOT, "video-stitcher_b200", "csrc", unit + ".cu")
        out = os.path.join(BUILD, unit + ".ptx")
        if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(os.path.join(os.path.dirname(src), f)) for f in os.listdir(os.path.dirname(src)) if f.endswith((".cu", ".cuh", ".h"))):
            subprocess.check_call([_NVCC, "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=compute_100a", "--fmad=false", "-ptx", src, "-o", out])
        kernels.update(P.parse(open(out).read()))
    return kernels


_CB = C.CFUNCTYPE(C.c_int, C.c_char_p, C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.POINTER(C.c_void_p), C.c_size_t)


def start():
    """Builds what is needed, loads the fake runtime and the emulation build of the library, installs the launch callback.
    Returns the ctypes library (the C ABI of include/vsb200.h)."""
    if "lib" in _state:
        return _state["lib"]
    kernels = _build()
    fake = C.CDLL(os.path.join(BUILD, "libfakecudart.so"), mode=C.RTLD_GLOBAL)
    fake.fake_alloc_get.argtypes = [C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    fake.fake_register.argtypes = [C.c_void_p, C.c_size_t]
    fake.fake_unregister.argtypes = [C.c_void_p]
    fake.fake_launch_count.restype = C.c_long

    def regions():
        out = []
        for i in range(fake.fake_alloc_count()):
            b, s = C.c_void_p(), C.c_size_t()
            fake.fake_alloc_get(i, C.byref(b), C.byref(s))
            out.append((b.value, s.value))
        return out

    mem = P.HostMemory(regions)
    stats = {"launches": [], "seconds": 0.0}

    def on_launch(name, gx, gy, gz, bx, by, bz, args, smem):
        try:
            k = kernels[name.decode()]
            params = [C.string_at(args[i], size) for i, (_, size, _) in enumerate(k.params)]
            t0 = time.time()
            mem._refresh()
            P.launch(k, (gx, gy, gz), (bx, by, bz), params, mem, dyn_smem=smem)
            stats["launches"].append((name.decode(), (gx, gy, gz), time.time() - t0))
            stats["seconds"] += time.time() - t0
            return 0
        except BaseException as e:  # noqa: the error must not unwind through the C frames of the library
            import traceback
            stats["error"] = "".join(traceback.format_exception(type(e), e, e.__traceback__))[-3000:]
            return 1

    cb = _CB(on_launch)
    fake.fake_set_launch_callback(cb)
    lib = C.CDLL(os.path.join(BUILD, "libvsb200_emu.so"))
    _state.update(lib=lib, fake=fake, cb=cb, stats=stats, kernels=kernels, mem=mem)
    return lib


def stats():
    return _state["stats"]


class Buffer:
    """A numpy array that the emulated runtime accepts as device memory."""

    def __init__(self, arr):
        self.a = np.ascontiguousarray(arr)
        self.ptr = self.a.ctypes.data
        _state["fake"].fake_register(self.ptr, self.a.nbytes)

    def release(self):
        _state["fake"].fake_unregister(self.ptr)
