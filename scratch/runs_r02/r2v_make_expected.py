"""CPU side of r2v: oracle-G panoramas (SHA-256) for the two features written after the GPU budget was spent -- compose_scale != 1 and
the split calibration -- so that the GPU side (r2v_check.py) needs neither torch nor the oracle and fits the ~10 s of box time left."""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import vsb200  # noqa: E402
from oracle import oracle as og  # noqa: E402
from oracle import pipeline as op  # noqa: E402

CASES = {
    "scale_small4": dict(n_views=4, src_w=320, src_h=240, pano_width=1024, num_bands=3, enable_local=True, compose_scale=0.75),
    "scale_mismatch6": dict(n_views=6, src_w=322, src_h=182, pano_width=960, num_bands=4, enable_local=True, compose_scale=0.8),
    "split_small4": dict(n_views=4, src_w=320, src_h=240, pano_width=1024, num_bands=3, enable_local=True),
    "split_cyl5": dict(n_views=5, src_w=256, src_h=192, pano_width=800, num_bands=4, enable_local=True, projection=1),
}

TINY = {   # for a dry run of r2v_check.py on the emulated runtime (R2V_EMU=1)
    "scale_tiny": dict(n_views=4, src_w=64, src_h=40, pano_width=192, num_bands=3, enable_local=True, compose_scale=0.75),
    "split_tiny": dict(n_views=4, src_w=48, src_h=32, pano_width=192, num_bands=3, enable_local=True),
}

if __name__ == "__main__":
    og.build()
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "r2v_expected.json")
    if len(sys.argv) > 2 and sys.argv[1] == "tiny":
        CASES, dst = TINY, sys.argv[2]
    S = vsb200.synth
    out = {}
    for name, kw in CASES.items():
        n = kw["n_views"]
        rig = op.OracleRig(gains=S.gains(n), **kw)
        for i in range(n):
            rig.set_mesh(i, *S.mesh(*rig.sizes[i]))
        frames = [S.frame(i, 0, kw["src_w"], kw["src_h"]) for i in range(n)]
        pano, _ = rig.compose(frames)
        out[name] = dict(kw=kw, roi_final=list(rig.roi_final), sha256=hashlib.sha256(np.ascontiguousarray(pano).tobytes()).hexdigest(),
                         sizes=[list(s) for s in rig.sizes])
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps({k: v["sha256"][:12] for k, v in out.items()}))
