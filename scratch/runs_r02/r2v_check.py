"""GPU side of r2v (no torch, no oracle: ctypes on libcudart + libvsb200): composes the four rigs of r2v_expected.json through
vsb_calibrate_rig_scaled / vsb_calibrate_rig_split and compares the SHA-256 of each panorama with oracle-G's."""
import ctypes as C
import hashlib
import json
import os
import sys
import time

import numpy as np

T0 = time.time()
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

EMU = bool(os.environ.get("R2V_EMU"))   # dry run of this script on the emulated runtime (tiny rigs; see README.md of this directory)
if EMU:
    from oracle.emu import runtime as E
    E.start()
    os.environ["VSB200_LIB"] = os.path.join(E.BUILD, "libvsb200_emu.so")
import vsb200  # noqa: E402

B, S = vsb200.binding, vsb200.synth
_keep = []
if EMU:
    def dev(a):
        b = E.Buffer(a)
        _keep.append(b)
        return b.ptr

    def fetch(out, d_out):
        out[...] = [b for b in _keep if b.ptr == d_out][0].a
else:
    rt = C.CDLL("libcudart.so.12")
    rt.cudaMalloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
    rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]

    def dev(a):
        p = C.c_void_p()
        assert rt.cudaMalloc(C.byref(p), a.nbytes) == 0
        assert rt.cudaMemcpy(p, a.ctypes.data, a.nbytes, 1) == 0
        return p.value

    def fetch(out, d_out):
        assert rt.cudaDeviceSynchronize() == 0
        assert rt.cudaMemcpy(out.ctypes.data, d_out, out.nbytes, 2) == 0


exp = json.load(open(sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "r2v_expected.json")))
res = {}
for name, e in exp.items():
    kw = e["kw"]
    n, sw, sh, nb = kw["n_views"], kw["src_w"], kw["src_h"], kw["num_bands"]
    proj = kw.get("projection", 0)
    gains = S.gains(n)
    frames = [S.frame(i, 0, sw, sh) for i in range(n)]
    if name.startswith("split"):
        plan = B.split_plan(proj, kw["pano_width"], n, sw, sh, nb)
        st = B.Stitcher(len(plan), nb, True, 1)
        st.calibrate_rig_split(proj, kw["pano_width"], n, sw, sh, 90.0, gains)
        cams = [st.view_window(k)[0] for k in range(len(plan))]
    else:
        st = B.Stitcher(n, nb, True, 1)
        st.calibrate_rig_scaled(proj, kw["pano_width"], sw, sh, kw["compose_scale"], 90.0, gains, on_device=False)
        cams = list(range(n))
    for k, c in enumerate(cams):
        mx, my = S.mesh(*e["sizes"][c])
        st.set_mesh(k, mx.ctypes.data, my.ctypes.data, mx.shape[0], mx.shape[1])
    roi, _, _ = st.get_roi()
    ok_roi = list(roi) == e["roi_final"]
    W, H = roi[2], roi[3]
    d_src = [dev(f) for f in frames]
    out = np.full((H, W, 3), -12345, np.int16)
    d_out = dev(out)
    st.compose([d_src[c] for c in cams], sw * 3, [d_out], W * 6, 0)
    fetch(out, d_out)
    got = hashlib.sha256(out.tobytes()).hexdigest()
    res[name] = dict(roi=ok_roi, pano=(got == e["sha256"]), views=len(cams), launches=st.last_launch_count())
    print(name, res[name], round(time.time() - T0, 2), flush=True)
print("R2V", "ALL OK" if all(v["roi"] and v["pano"] for v in res.values()) and len(res) == len(exp) else "MISMATCH", json.dumps(res), round(time.time() - T0, 2), "s")
