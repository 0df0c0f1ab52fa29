"""Writes tests/golden/oracle_compose_hashes.json: SHA-256 of oracle-G's composed panoramas for two small rigs (whole-path known-answer
vectors: they freeze the oracle this round's GPU parity runs were green against; tests/test_oracle_pin.py replays them)."""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import vsb200
from oracle import oracle as og, pipeline as op
S = vsb200.synth
og.set_num_threads(8)
out = {}
for name, kw in (("small4", dict(n_views=4, src_w=320, src_h=240, pano_width=1024, num_bands=3, enable_local=True)),
                 ("cyl5", dict(n_views=5, src_w=256, src_h=192, pano_width=800, num_bands=4, enable_local=True, projection=1)),
                 # the rig bench.py composes view-sharded on N GPUs to verify parity inside the scaling run ("parity_checked")
                 ("shard6", dict(n_views=6, src_w=480, src_h=270, pano_width=1536, num_bands=4, enable_local=True)),
                 # compose_scale != 1 (exact sizes; the reference's cvRound / (int) mismatch): the two panoramas a B200 reproduced through
                 # the C ABI in round 2 (profiles/r02_hw_check_compose_scale_and_split.log; same hashes in scratch/runs_r02/r2v_expected.json)
                 ("scale_small4", dict(n_views=4, src_w=320, src_h=240, pano_width=1024, num_bands=3, enable_local=True, compose_scale=0.75)),
                 ("scale_mismatch6", dict(n_views=6, src_w=322, src_h=182, pano_width=960, num_bands=4, enable_local=True, compose_scale=0.8))):
    rig = op.OracleRig(gains=S.gains(kw["n_views"]), **kw)
    for i in range(kw["n_views"]):
        rig.set_mesh(i, *S.mesh(*rig.sizes[i]))
    frames = [S.frame(i, 0, kw["src_w"], kw["src_h"]) for i in range(kw["n_views"])]
    pano, mask = rig.compose(frames)
    out[name] = {"shape": list(pano.shape), "sha256": hashlib.sha256(np.ascontiguousarray(pano).tobytes()).hexdigest(),
                 "mask_sha256": hashlib.sha256(np.ascontiguousarray(mask).tobytes()).hexdigest(), "sum": int(pano.astype(np.int64).sum())}
json.dump(out, open(os.path.join(ROOT, 'tests', 'golden', 'oracle_compose_hashes.json'), 'w'), indent=1)
print(out)
