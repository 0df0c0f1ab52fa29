"""TEST INFRASTRUCTURE: the PRODUCT's device seam finder on the emulated runtime -- vsb_voronoi_seams_device (k_vor_columns +
k_vor_decide: VoronoiSeamFinder::find, S/src/seam_finders.cpp:72-162, as an exact separable L1 distance transform) interpreted from
the product's PTX, on the seam-scale masks of the first rig of tests/golden/make_golden.py SEAM_RIGS.  Prints one JSON line with the
SHA-256 of every resulting mask (packed bits), which tests/test_emulated_pipeline.py compares with the reference's own seams
(tests/golden/reference_cpu.npz, produced by the reference's VoronoiSeamFinder).

    python -m oracle.emu.run_voronoi_case
"""
import ctypes as C
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    from oracle.emu import runtime as E
    E.start()
    os.environ["VSB200_LIB"] = os.path.join(E.BUILD, "libvsb200_emu.so")
    import vsb200
    from oracle import oracle as og
    from tests.golden import make_golden as G
    og.build()
    L = vsb200.binding.lib()
    n, sw, sh, pano, proj = G.SEAM_RIGS[0]
    masks, corners, sizes = G.seam_inputs(og, n, sw, sh, pano, proj)
    bufs = [E.Buffer(m) for m in masks]
    ptrs = (C.c_void_p * n)(*[b.ptr for b in bufs])
    sz = (C.c_int * (2 * n))(*[int(v) for p in sizes for v in p])
    co = (C.c_int * (2 * n))(*[int(v) for p in corners for v in p])
    rc = L.vsb_voronoi_seams_device(n, sz, co, ptrs, None)
    print(json.dumps({"rc": rc, "error": E.stats().get("error"), "rig": [n, sw, sh, pano, proj],
                      "values": sorted(set(int(v) for b in bufs for v in np.unique(b.a))),
                      "packed_sha256": [hashlib.sha256(np.packbits(b.a > 0, axis=1).tobytes()).hexdigest() for b in bufs],
                      "launched": sorted(set(name for name, _, _ in E.stats()["launches"]))}), flush=True)


if __name__ == "__main__":
    main()
