"""TEST INFRASTRUCTURE: the view-sharded mode of the PRODUCT library walked through on the emulated runtime (oracle/emu/runtime.py) --
`world` handles in one process, one per rank, as tests/test_gpu_shard.py::_emulated_ranks does on one GPU: every rank calibrates the
same rig, takes its shard (vsb_shard_set / vsb_shard_plan), runs the front half of its views, packs one message per peer
(k_shard_copy); the messages are handed over in-process (what the grouped ncclSend / ncclRecv of vsb_shard_compose moves), unpacked,
and every rank blends its canvas strip.  The strips summed must be oracle-G's panorama, bit for bit.

    python -m oracle.emu.run_shard_case '{"n_views": 4, "src_w": 96, "src_h": 64, "pano_width": 512, "num_bands": 3, "world": 2}'

"split": true takes the split calibration (cameras that wrap around +-pi are two views, which different ranks may own).
Prints one JSON line."""
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
    B, S, D = vsb200.binding, vsb200.synth, vsb200.dist
    n, sw, sh, pano, nb, world = case["n_views"], case["src_w"], case["src_h"], case["pano_width"], case["num_bands"], case["world"]
    F = int(case.get("frames", 1))
    split = bool(case.get("split"))
    gains = S.gains(n)
    plan = B.split_plan(0, pano, n, sw, sh, nb) if split else [(i, 0, 0) for i in range(n)]
    nv = len(plan)
    t0 = time.time()
    hs = []
    for r in range(world):
        st = B.Stitcher(nv, nb, True, F)
        if split:
            st.calibrate_rig_split(0, pano, n, sw, sh, 90.0, gains)
        else:
            st.calibrate_rig(0, pano, sw, sh, 90.0, gains)
        info = st.rig_info()
        for k in range(nv):
            mx, my = S.mesh(st.view_window(k)[2], info.view_roi[k][3])
            st.set_mesh(k, mx.ctypes.data, my.ctypes.data, mx.shape[0], mx.shape[1])
        st.shard_set(r, world)
        hs.append(st)
    cam = [hs[0].view_window(k)[0] for k in range(nv)]
    roi, _, _ = hs[0].get_roi()
    W, H = roi[2], roi[3]
    infos = [st.shard_info() for st in hs]
    owners = [None] * nv
    for r, (_, _, owned) in enumerate(infos):
        for v in owned:
            assert owners[v] is None, f"view {v} has two owners"
            owners[v] = r
    assert None not in owners, f"unowned view: {owners}"
    for st in hs:
        st.shard_plan(owners)
    t_cal = time.time() - t0
    frames = [[S.frame(i, f, sw, sh) for i in range(n)] for f in range(F)]
    d_fr = [[E.Buffer(fr[cam[k]]) for k in range(nv)] for fr in frames]
    t0 = time.time()
    n0 = len(E.stats()["launches"])
    bufs, facts = {}, []
    for r, st in enumerate(hs):                                  # front halves + pack
        owned = infos[r][2]
        for v0, v1 in D.contiguous_runs(owned):
            st.feed_batch(v0, v1, F, [d_fr[f][v].ptr for f in range(F) for v in range(v0, v1)], sw * 3, 0)
        sent = 0
        for p in range(world):
            if p == r:
                continue
            sb, _ = st.shard_peer_bytes(p)
            _, rb = hs[p].shard_peer_bytes(r)
            assert sb == rb and sb % 16 == 0, (r, p, sb, rb)
            if sb:
                bufs[(r, p)] = E.Buffer(np.full(sb * F, 0xAB, np.uint8))
                st.shard_pack(p, F, bufs[(r, p)].ptr, 0)
                sent += sb
        facts.append({"rank": r, "owned": owned, "strip": list(infos[r][:2]), "send_bytes_per_frame": sent})
    total = [np.zeros((H, W, 3), np.int32) for _ in range(F)]
    for r, st in enumerate(hs):                                  # "exchange" + unpack + back halves
        for p in range(world):
            if (p, r) in bufs:
                st.shard_unpack(p, F, bufs[(p, r)].ptr, 0)
        outs = [E.Buffer(np.zeros((H, W, 3), np.int16)) for _ in range(F)]
        st.blend_batch([o.ptr for o in outs], W * 6, 0)
        for f in range(F):
            total[f] += outs[f].a
    t_run = time.time() - t0
    launched = collections.Counter(name for name, _, _ in E.stats()["launches"][n0:])

    orig = op.OracleRig(n, sw, sh, pano, num_bands=nb, enable_local=True, gains=gains)
    for i in range(n):
        orig.set_mesh(i, *S.mesh(*orig.sizes[i]))
    res = collections.OrderedDict(error=E.stats().get("error"), roi_equal=bool(tuple(roi) == tuple(orig.roi_final)), pano=0)
    for f in range(F):
        want, _ = orig.compose(frames[f])
        res["pano"] += int(np.count_nonzero(total[f].astype(np.int16) != want))
    res["pano_samples"] = int(F * H * W * 3)
    res["pano_nonzero"] = int(np.count_nonzero(total[0]))
    res["ranks"] = facts
    res["views"] = [list(p) for p in plan] if split else nv
    res["shard_copy_launches"] = int(sum(c for k, c in launched.items() if "k_shard_copy" in k))
    res["seconds"] = {"calibrate_and_plan": round(t_cal, 1), "front_exchange_back": round(t_run, 1)}
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    try:
        main()
    except Exception:
        from oracle.emu import runtime as E
        print("emulated runtime error:", E.stats().get("error"), file=sys.stderr)
        raise
