"""TEST INFRASTRUCTURE: one end-to-end run of the PRODUCT library on the emulated runtime (oracle/emu/runtime.py), compared with
oracle-G.  Run as a module in its own process (the emulation build of the library is selected through VSB200_LIB before
video-stitcher_b200/binding.py is imported):

    python -m oracle.emu.run_case '{"n_views": 4, "src_w": 48, "src_h": 32, "pano_width": 192, "num_bands": 3}'

Prints one JSON line: mismatch counts of the mesh maps, the warped views, the Gaussian levels the path materialises and the
panorama, the kernels that were launched and the time spent interpreting them."""
import collections
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    case = json.loads(sys.argv[1])
    from oracle.emu import runtime as E
    E.start()
    os.environ["VSB200_LIB"] = os.path.join(E.BUILD, "libvsb200_emu.so")
    import vsb200
    from oracle import oracle as og
    from oracle import pipeline as op
    og.build()
    B, S = vsb200.binding, vsb200.synth
    assert B.LIB_PATH.endswith("libvsb200_emu.so")
    n, sw, sh, pano, nb = case["n_views"], case["src_w"], case["src_h"], case["pano_width"], case["num_bands"]
    F = int(case.get("frames", 1))
    proj = int(case.get("projection", 0))                  # 0 spherical, 1 cylindrical
    gains = S.gains(n)
    t0 = time.time()
    split = bool(case.get("split"))                        # cameras that wrap around +-pi become two views (vsb_calibrate_rig_split)
    plan = B.split_plan(proj, pano, n, sw, sh, nb) if split else [(i, 0, 0) for i in range(n)]
    nv = len(plan)
    st = B.Stitcher(nv, nb, True, F)
    cs = float(case.get("compose_scale", 1.0))             # != 1: sw x sh are the full frames, resized on the device in front of remap #1
    ws = 1.0
    if case.get("megapix"):                                # stitch_calib's own scales: [WORK_MEGAPIX, COMPOSE_MEGAPIX]; pano_width is ignored (0)
        wm, cm = case["megapix"]
        ws, cs = op.ref_scales(sw, sh, wm, cm)
        pano = 0
        st.calibrate_rig_megapix(proj, sw, sh, wm, cm, 90.0, gains, on_device=bool(case.get("device_calibration")))
    elif split:
        st.calibrate_rig_split(proj, pano, n, sw, sh, 90.0, gains, on_device=bool(case.get("device_calibration")))
    elif cs != 1.0:
        st.calibrate_rig_scaled(proj, pano, sw, sh, cs, 90.0, gains, on_device=bool(case.get("device_calibration")))
    elif case.get("device_calibration"):                   # every per-pixel loop of the calibration as kernels (maps differ from libm's by ulps)
        st.calibrate_rig_device(proj, pano, sw, sh, 90.0, gains)
    else:
        st.calibrate_rig(proj, pano, sw, sh, 90.0, gains)   # the product's host calibration + its weight / plan kernels
    info = st.rig_info()
    win = [st.view_window(k) for k in range(nv)]            # (camera, x0, width of the camera's warped image) per view
    for k in range(nv):
        mx, my = S.mesh(win[k][2], info.view_roi[k][3])     # the mesh of the view's CAMERA
        st.set_mesh(k, mx.ctypes.data, my.ctypes.data, mx.shape[0], mx.shape[1])
    t_cal = time.time() - t0
    roi, _, _ = st.get_roi()
    W, H = roi[2], roi[3]
    wire = bool(case.get("wire"))       # NV12 frames in (converted inside remap #1's tap fetch), CV_8UC3 panoramas out
    if wire:
        st.set_formats(B.IN_NV12, B.OUT_U8C3)
        nvf = [[S.frame_nv12(i, f, sw, sh) for i in range(n)] for f in range(F)]
        frames = [[og.nv12_to_bgr(a, sw, sh) for a in fr] for fr in nvf]
        bufs = [E.Buffer(fr[win[k][0]]) for fr in nvf for k in range(nv)]
        outs = [E.Buffer(np.full((H, W, 3), 0xAB, np.uint8)) for _ in range(F)]
    else:
        frames = [[S.frame(i, f, sw, sh) for i in range(n)] for f in range(F)]
        bufs = [E.Buffer(fr[win[k][0]]) for fr in frames for k in range(nv)]
        outs = [E.Buffer(np.full((H, W, 3), -12345, np.int16)) for _ in range(F)]
    n0 = len(E.stats()["launches"])
    t0 = time.time()
    st.compose([b.ptr for b in bufs], sw if wire else sw * 3, [o.ptr for o in outs], W * (3 if wire else 6), 0)
    t_compose = time.time() - t0
    launched = [name for name, _, _ in E.stats()["launches"][n0:]]

    orig = op.OracleRig(n, sw, sh, pano, projection=proj, num_bands=nb, enable_local=True, gains=gains, compose_scale=cs, work_scale=ws)
    for i in range(n):
        orig.set_mesh(i, *S.mesh(*orig.sizes[i]))

    def read(what, view, level, shape, dtype):
        a = np.empty(shape, dtype)
        st.debug_read(what, view, level, 0, a.ctypes.data, a.nbytes)
        return a

    res = collections.OrderedDict(error=E.stats().get("error"), roi_equal=bool(tuple(roi) == tuple(orig.roi_final)), mesh_maps=0, warped=0, gauss0=0, gauss2=0)
    for i in range(nv):                                        # view i = columns [x0, x0 + w) of camera c's warped image
        w, h = info.view_roi[i][2], info.view_roi[i][3]
        c, x0, _ = win[i]
        g = st.view_geometry(i)
        bw, bh = g["x_br"] - g["x_tl"], g["y_br"] - g["y_tl"]
        omx = np.ascontiguousarray(orig.mesh_maps[c][0][:, x0:x0 + w] - np.float32(x0))   # (x - x0 in binary32, as the window kernel re-bases it)
        omy = np.ascontiguousarray(orig.mesh_maps[c][1][:, x0:x0 + w])
        gmx, gmy = read(4, i, 0, (h, w), np.float32), read(5, i, 0, (h, w), np.float32)
        res["mesh_maps"] += int(np.count_nonzero((gmx.view(np.uint32) != omx.view(np.uint32)) & ~(np.isnan(gmx) & np.isnan(omx))))
        res["mesh_maps"] += int(np.count_nonzero((gmy.view(np.uint32) != omy.view(np.uint32)) & ~(np.isnan(gmy) & np.isnan(omy))))
        fast = nb >= 3                                         # below 3 bands the generic per-level kernels run and every sample is computed
        done0 = read(7, i, 0, (bh, bw), np.uint8).astype(bool) if fast else np.ones((bh, bw), bool)
        ov = np.ascontiguousarray(orig.warp_view(c, frames[0][c])[:, x0:x0 + w])
        crop = done0[g["top"]:g["top"] + h, g["left"]:g["left"] + w]
        res["warped"] += int(np.count_nonzero(read(0, i, 0, (h, w, 3), np.uint8)[crop] != ov[crop]))
        gk = og.border_reflect_u8c3_to_s16(ov, g["top"], g["bottom"], g["left"], g["right"])
        res["gauss0"] += int(np.count_nonzero(read(1, i, 0, (bh, bw, 3), np.int16)[done0] != gk[done0]))
        g2 = og.pyr_down_s16(og.pyr_down_s16(gk))
        done2 = read(6, i, 0, (bh >> 2, bw >> 2), np.uint8).astype(bool) if fast else np.ones((bh >> 2, bw >> 2), bool)
        if nb >= 2:
            res["gauss2"] += int(np.count_nonzero(read(1, i, 2, (bh >> 2, bw >> 2, 3), np.int16)[done2] != g2[done2]))
    res["pano"] = 0
    for f in range(F):
        want, _ = orig.compose(frames[f])
        if wire:
            want = og.s16_to_u8(want)
        res["pano"] += int(np.count_nonzero(outs[f].a != want))
        if os.environ.get("VSB_EMU_DUMP"):               # debugging aid: the two panoramas side by side
            np.save(os.environ["VSB_EMU_DUMP"] + f"_got{f}.npy", outs[f].a)
            np.save(os.environ["VSB_EMU_DUMP"] + f"_want{f}.npy", want)
    if wire:   # the consumer epilogue on the device (vsb_consume: fixed-point cv::resize + BGR2RGB / letter-boxed BGR2YUV_I420, A/timed.cpp:254-315)
        want_u8 = og.s16_to_u8(orig.compose(frames[0])[0])
        ow, oh = 96, 64
        ih = og.consumer_image_height(W, H, ow, oh, True)
        rgb = E.Buffer(np.full((ih, ow, 3), 0xCD, np.uint8))
        st.consume(outs[0].ptr, W * 3, ow, oh, B.CONSUME_RGB, rgb.ptr, ow * 3)
        res["consume_rgb"] = int(np.count_nonzero(rgb.a != og.consume(want_u8, ow, oh, 0)))
        yuv = E.Buffer(np.full(ow * oh * 3 // 2, 0xCD, np.uint8))
        st.consume(outs[0].ptr, W * 3, ow, oh, B.CONSUME_I420, yuv.ptr, ow)
        res["consume_i420"] = int(np.count_nonzero(yuv.a != og.consume(want_u8, ow, oh, 1).reshape(-1)))
    import hashlib
    res["pano_sha256"] = hashlib.sha256(b"".join(np.ascontiguousarray(o.a).tobytes() for o in outs)).hexdigest()[:16]
    res["pano_samples"] = int(F * H * W * 3)
    res["pano_nonzero"] = int(np.count_nonzero(outs[0].a))
    res["launched"] = [k.split("vsb")[-1][:28] for k in launched]
    res["views"] = [list(p) for p in plan] if split else nv
    res["launch_count"] = st.last_launch_count()
    res["calibration_launches"] = n0
    res["seconds"] = {"calibrate_and_mesh": round(t_cal, 1), "compose": round(t_compose, 1)}
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    try:
        main()
    except Exception:
        from oracle.emu import runtime as E
        print("emulated runtime error:", E.stats().get("error"), file=sys.stderr)
        raise
