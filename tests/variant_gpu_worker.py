"""Subprocess worker for test_remap_kernel_variants: composes one frame of a small rig with the remap kernel form chosen
by VSB_REMAP_VARIANT (read once per process by libvsb200) and saves the panorama + warped views for the parent to check."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch


def main():
    out_path, pad = sys.argv[1], int(sys.argv[2])
    import vsb200
    from tests.gpu_util import GpuRig, host, stream
    from tests.test_gpu_parity import CASES
    kw = dict(CASES["small4"])
    gains = vsb200.synth.gains(kw["n_views"])
    rig = GpuRig(gains=gains, max_batch=2, **kw)
    for i in range(kw["n_views"]):
        mx, my = vsb200.synth.mesh(*rig.sizes[i])
        rig.set_mesh(i, mx, my)
    W, H = rig.roi_final[2], rig.roi_final[3]
    sw, sh = kw["src_w"], kw["src_h"]
    pitch = sw * 3 + pad  # pad 0: tight rows (the last bytes of the frame end the allocation); pad 1: unaligned rows
    srcs = []
    for f in range(2):
        for i in range(kw["n_views"]):
            buf = torch.zeros(sh * pitch, dtype=torch.uint8, device="cuda")
            buf.view(sh, pitch)[:, :sw * 3] = torch.from_numpy(vsb200.synth.frame(i, f, sw, sh).reshape(sh, sw * 3)).cuda()
            srcs.append(buf)
    outs = [torch.full((H, W, 3), -12345, dtype=torch.int16, device="cuda") for _ in range(2)]
    rig.st.compose([t.data_ptr() for t in srcs], pitch, [o.data_ptr() for o in outs], W * 6, stream())
    np.savez(out_path, pano0=host(outs[0]), pano1=host(outs[1]), **{f"warped{i}": rig.warped(i, 1) for i in range(kw["n_views"])})


if __name__ == "__main__":
    main()
