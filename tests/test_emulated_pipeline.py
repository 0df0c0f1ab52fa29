"""CPU: the whole PRODUCT library, end to end, without a GPU.

oracle/emu builds libvsb200_emu.so from the product's own object files and a stand-in CUDA runtime (device memory = host memory,
streams inert); every kernel the product's host code launches -- calibration (weight pyramids, plans), vsb_set_mesh (splat /
divide / upsample / tap table), and the frame path K1 K2 k_down2 k_down_tail k_coarse k_blend -- is executed by the PTX
interpreter (oracle/ptx_interp.py) on the PTX of the same .cu files, compiled with the product's flags.  So this runs the shipped
host logic (ROIs, seams, tile lists, launch sequences) and the shipped device code, and compares the panorama with oracle-G bit
for bit -- the `-m gpu` parity test of a small rig, minus the hardware.  (k_down2 takes its plain-load form: no tensor-map encoder
here, as on a driver without one; atomics are sequential; timing-dependent behaviour is not modelled.)"""
import json
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


DEFAULT_CASES = {   # interpreted in parallel (one process each, ~70 s): started together by the first test that needs one
    "base": dict(n_views=4, src_w=48, src_h=32, pano_width=192, num_bands=3),
    "megapix": dict(n_views=4, src_w=64, src_h=40, pano_width=0, num_bands=3, megapix=[0.0015, 0.0016]),
    "split": dict(n_views=4, src_w=48, src_h=32, pano_width=192, num_bands=3, split=True),
    "wire": dict(n_views=4, src_w=48, src_h=32, pano_width=192, num_bands=3, wire=True),
    "bands2": dict(n_views=4, src_w=48, src_h=24, pano_width=192, num_bands=2),
}
_procs = {}


def _spawn(case, module="oracle.emu.run_case"):
    env = {k: v for k, v in os.environ.items() if k != "VSB200_LIB"}
    cmd = [sys.executable, "-m", module] + ([json.dumps(case)] if case is not None else [])
    return subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=ROOT, env=env)


def _collect(p):
    try:
        out, err = p.communicate(timeout=1500)
    except subprocess.TimeoutExpired:
        p.kill()
        raise
    assert p.returncode == 0, err[-3000:]
    return json.loads([l for l in out.splitlines() if l.startswith("{")][-1])


def _need_nvcc():
    if not (shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc")):
        pytest.skip("nvcc not found (the emulation needs the product's PTX)")


def _run(case):
    _need_nvcc()
    return _collect(_spawn(case))


def _default(name):
    _need_nvcc()
    if not _procs:
        # build the emulation library once, before the processes race for it
        subprocess.run([sys.executable, "-c", "from oracle.emu import runtime as E; E.start()"], cwd=ROOT, check=True, capture_output=True, timeout=900)
        for k, case in DEFAULT_CASES.items():
            _procs[k] = _spawn(case)
        _procs["voronoi"] = _spawn(None, "oracle.emu.run_voronoi_case")
    return _collect(_procs[name])


def test_product_library_end_to_end_on_the_emulated_runtime():
    res = _default("base")
    assert res["error"] is None, res["error"]
    assert res["roi_equal"]
    assert res["mesh_maps"] == 0, "vsb_set_mesh kernels vs oracle mesh -> map"
    assert res["warped"] == 0 and res["gauss0"] == 0 and res["gauss2"] == 0, res
    assert res["pano"] == 0 and res["pano_nonzero"] > res["pano_samples"] // 2, res
    names = " ".join(res["launched"])
    for k in ("k_remap_stage1_tab", "k_remap_stage2_tab", "k_down2", "k_down_tail", "k_coarse", "k_blend"):
        assert k in names, (k, res["launched"])
    assert res["launch_count"] in (6, 7)      # K1 K2 down2 down_tail coarse blend_seam (+ blend_int when a tile is interior)


@pytest.mark.skipif(not os.environ.get("VSB_EMU_FULL"), reason="several minutes of interpretation: set VSB_EMU_FULL=1 (two frames, interior blend tiles)")
def test_product_library_end_to_end_larger_rig():
    res = _run(dict(n_views=4, src_w=96, src_h=64, pano_width=384, num_bands=3, frames=2))
    assert res["error"] is None and res["pano"] == 0 and res["warped"] == 0 and res["gauss2"] == 0, res
    assert "k_blend_int" in " ".join(res["launched"]), res["launched"]


def test_product_library_wire_formats_on_the_emulated_runtime():
    """VSB_IN_NV12 / VSB_OUT_U8C3: the NV12 conversion fused into remap #1's tap fetch (k_remap_stage1_nv12) and the CV_8UC3 store of the
    blend kernels, against og.nv12_to_bgr -> compose -> og.s16_to_u8."""
    res = _default("wire")
    assert res["error"] is None and res["pano"] == 0 and res["warped"] == 0, res
    assert "k_remap_stage1_nv12" in " ".join(res["launched"]), res["launched"]
    # ... and the consumer epilogue on that CV_8UC3 panorama (vsb_consume: the reference's fixed-point cv::resize, then BGR2RGB or the
    # letter-boxed BGR2YUV_I420 frame the encoder is fed, A/timed.cpp:254-315) against og.consume
    assert res["consume_rgb"] == 0 and res["consume_i420"] == 0, res


def test_device_seam_finder_on_the_emulated_runtime():
    """vsb_voronoi_seams_device (k_vor_columns + k_vor_decide, interpreted from the product's PTX) on the seam-scale masks of a 6-view
    rig: the masks it leaves are the ones the REFERENCE's own VoronoiSeamFinder leaves (tests/golden/reference_cpu.npz), bit for bit."""
    import hashlib
    import numpy as np
    res = _default("voronoi")
    assert res["rc"] == 0 and res["error"] is None and set(res["values"]) <= {0, 255}, res
    assert any("k_vor_columns" in k for k in res["launched"]) and any("k_vor_decide" in k for k in res["launched"]), res["launched"]
    gold = np.load(os.path.join(ROOT, "tests", "golden", "reference_cpu.npz"))
    n, _, _, pano, proj = res["rig"]
    want = [hashlib.sha256(np.ascontiguousarray(gold[f"voronoi_{n}_{pano}_{proj}_{i}"]).tobytes()).hexdigest() for i in range(n)]
    assert res["packed_sha256"] == want


def test_product_library_reference_scales_on_the_emulated_runtime():
    """stitch_calib's own scales through the shipped library (vsb_calibrate_rig_megapix; A/calibration.cpp:256-305,137-205, A/timed.cpp:74-81):
    work_scale 0.765 and compose_scale 0.791 from (scaled-down) WORK / COMPOSE_MEGAPIX constants, cameras at work scale, warped_image_scale
    = (float)cameras[0].focal, seam_work_aspect and compose_work_aspect as ratios; blender sized from cvRound(full * scale) = (51, 32), maps
    and masks built for (int)(full * scale) = (50, 31) -- the reference's own mismatch, reproduced -- and the per-frame cuda::resize in
    front of remap #1 (k_prescale): bit-identical to oracle-G."""
    res = _default("megapix")
    assert res["error"] is None and res["roi_equal"], res
    assert res["mesh_maps"] == 0 and res["warped"] == 0 and res["gauss0"] == 0 and res["gauss2"] == 0, res
    assert res["pano"] == 0 and res["pano_nonzero"] > res["pano_samples"] // 2, res
    assert "k_prescale" in " ".join(res["launched"]), res["launched"]


def test_product_library_split_calibration_on_the_emulated_runtime():
    """Modular wrap-around ROI (wrapAround, A/defs.h:25): vsb_calibrate_rig_split installs the camera that looks across +-pi as two
    views -- column windows of its warped image, with the margin and origin rules of tests/test_oracle_wrap_split.py -- so no buffer
    is panorama-wide; vsb_set_mesh on a window view takes the camera's mesh (k_mesh_upsample_win).  The panorama the shipped library
    composes from the five views equals the UNSPLIT oracle's four-view panorama bit for bit; the per-view intermediates equal the
    corresponding columns of the full-width view's."""
    res = _default("split")
    assert res["error"] is None and res["roi_equal"], res
    assert len(res["views"]) == 5 and [v[0] for v in res["views"]] == [0, 1, 2, 2, 3], res["views"]
    assert max(v[2] for v in res["views"]) < 192 // 2, res["views"]
    assert res["mesh_maps"] == 0 and res["warped"] == 0 and res["gauss0"] == 0 and res["gauss2"] == 0, res
    assert res["pano"] == 0 and res["pano_nonzero"] > res["pano_samples"] // 2, res


@pytest.mark.skipif(not os.environ.get("VSB_EMU_FULL"), reason="a quarter of an hour of interpretation: set VSB_EMU_FULL=1 (view-sharded mode, 2 and 3 ranks)")
@pytest.mark.parametrize("world,split", [(2, False), (3, True)])
def test_view_sharded_mode_on_the_emulated_runtime(world, split):
    """The north-star multi-GPU split without GPUs: `world` handles in one process (one per rank) on the emulated runtime -- shard
    ownership and plan, front halves of the owned views, k_shard_copy pack, the per-peer messages handed over in-process (what
    vsb_shard_compose's grouped ncclSend / ncclRecv moves), k_shard_copy unpack, strip blends -- and the strips summed are oracle-G's
    panorama bit for bit.  With split=True the rig is calibrated with vsb_calibrate_rig_split (5 views from 4 cameras)."""
    _need_nvcc()
    case = dict(n_views=4, src_w=96, src_h=64, pano_width=512, num_bands=3, world=world, split=split)
    env = {k: v for k, v in os.environ.items() if k != "VSB200_LIB"}
    r = subprocess.run([sys.executable, "-m", "oracle.emu.run_shard_case", json.dumps(case)], capture_output=True, text=True, timeout=3000, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-3000:]
    res = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert res["error"] is None and res["roi_equal"] and res["pano"] == 0 and res["pano_nonzero"] > res["pano_samples"] // 2, res
    assert len(res["ranks"]) == world and sum(x["send_bytes_per_frame"] for x in res["ranks"]) > 0 and res["shard_copy_launches"] >= 2, res
    owned = sorted(v for x in res["ranks"] for v in x["owned"])
    assert owned == list(range(5 if split else 4)), res["ranks"]


def test_device_gain_estimation_on_the_emulated_runtime():
    """vsb_gain_compensator_feed on the emulated runtime: k_gain_pairs -- binary64 sums in the reference's summation order, interpreted
    from the product's PTX -- and the host LU solve give the gains of the REFERENCE's own GainCompensator class (committed in
    tests/golden/reference_cpu.npz from oracle/_ref), bit for bit, float64."""
    import numpy as np
    _need_nvcc()
    env = {k: v for k, v in os.environ.items() if k != "VSB200_LIB"}
    r = subprocess.run([sys.executable, "-m", "oracle.emu.run_gain_case"], capture_output=True, text=True, timeout=900, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-3000:]
    res = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert res["rc"] == 0 and res["error"] is None and any("k_gain_pairs" in k for k in res["launched"]), res
    want = np.load(os.path.join(ROOT, "tests", "golden", "reference_cpu.npz"))["gain_compensator"]
    assert [float.fromhex(h) for h in res["gains_hex"]] == [float(v) for v in want]


def test_generic_path_below_three_bands_on_the_emulated_runtime():
    """num_bands = 2: the generic per-level kernels (k_pyr_down<u8> / <s16>, k_legacy_blend_collapse: Laplacian, weighted add,
    normalise, collapse, mask and crop of every level in one kernel) instead of the fused fast path -- bit-identical to oracle-G."""
    res = _default("bands2")
    assert res["error"] is None and res["roi_equal"] and res["warped"] == 0 and res["gauss0"] == 0 and res["gauss2"] == 0, res
    assert res["pano"] == 0 and res["pano_nonzero"] > res["pano_samples"] // 2, res
    assert "k_legacy_blend_collapse" in " ".join(res["launched"]) and res["launch_count"] == 5, res["launched"]


@pytest.mark.skipif(not os.environ.get("VSB_EMU_FULL"), reason="another minute of interpretation: set VSB_EMU_FULL=1 (compose_scale with a pano_width)")
def test_product_library_compose_scale_on_the_emulated_runtime():
    """vsb_calibrate_rig_scaled: compose_scale 0.8 on 61 x 41 frames with this repository's pano_width parametrisation (work_scale 1)."""
    res = _run(dict(n_views=4, src_w=61, src_h=41, pano_width=192, num_bands=3, compose_scale=0.8))
    assert res["error"] is None and res["roi_equal"] and res["pano"] == 0 and res["warped"] == 0 and res["mesh_maps"] == 0, res
    assert "k_prescale" in " ".join(res["launched"]), res["launched"]
