"""TEST INFRASTRUCTURE: the PRODUCT's device gain estimation on the emulated runtime -- vsb_gain_compensator_feed, i.e. k_gain_pairs
(the pairwise overlap counts and intensity sums of GainCompensator::feed, S/src/exposure_compensate.cpp:89-121, accumulated in binary64
in the reference's order, interpreted from the product's PTX) + the host LU solve -- on the committed input of the reference's own
GainCompensator (tests/golden/make_golden.py gain_input).  Prints one JSON line with the gains as hex doubles.

    python -m oracle.emu.run_gain_case
"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    from oracle.emu import runtime as E
    E.start()
    os.environ["VSB200_LIB"] = os.path.join(E.BUILD, "libvsb200_emu.so")
    import vsb200
    from tests.golden import make_golden as G
    L = vsb200.binding.lib()
    imgs, masks, corners, sizes = G.gain_input()
    n = len(imgs)
    bi, bm = [E.Buffer(a) for a in imgs], [E.Buffer(a) for a in masks]
    ip, mp = (C.c_void_p * n)(*[b.ptr for b in bi]), (C.c_void_p * n)(*[b.ptr for b in bm])
    sz = (C.c_int * (2 * n))(*[int(v) for p in sizes for v in p])
    co = (C.c_int * (2 * n))(*[int(v) for p in corners for v in p])
    g = (C.c_double * n)()
    rc = L.vsb_gain_compensator_feed(n, ip, mp, sz, co, g, None)
    print(json.dumps({"rc": rc, "error": E.stats().get("error"), "gains_hex": [float(v).hex() for v in g],
                      "launched": [name for name, _, _ in E.stats()["launches"]]}), flush=True)


if __name__ == "__main__":
    main()
