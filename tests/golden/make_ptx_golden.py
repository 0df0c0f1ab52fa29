"""Writes tests/golden/reference_ptx.npz: what the reference's own CUDA kernels compute on the seeded cases of oracle/ptx_cases.py.
The kernels are the reference's unmodified .cu files compiled to PTX by oracle/ref_ptx.mk (needs /root/reference and nvcc) and
executed on the CPU by oracle/ptx_interp.py.  Run from the repository root:  python tests/golden/make_ptx_golden.py"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    subprocess.check_call(["make", "-s", "-f", "ref_ptx.mk"], cwd=os.path.join(ROOT, "oracle"))
    from oracle import ptx_cases as PC
    out = {}
    for name, (_, run, _, _) in PC.CASES.items():
        for key, arr in run(PC.inputs_of(name)).items():
            out[f"{name}__{key}"] = arr
    from oracle import oracle as og
    og.build()
    for key, arr in PC.build_warp_maps_ptx(og).items():   # the map builder: device sinf / cosf as compiled into the kernel
        out[f"build_warp_maps__{key}"] = arr
    path = os.path.join(ROOT, "tests", "golden", "reference_ptx.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
