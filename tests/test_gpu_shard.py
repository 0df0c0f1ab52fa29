"""-m gpu, needs >= 2 GPUs (skipped on a 1-GPU box): the view-sharded mode over NCCL gives the oracle's panorama bit-exactly."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 4])
def test_view_sharded_compose_matches_oracle(cuda, world):
    if cuda.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    case = dict(n_views=6, src_w=480, src_h=270, pano_width=1536, num_bands=4)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
                        "--master-port", "29621", os.path.join(ROOT, "tests", "shard_gpu_worker.py"), json.dumps(case)],
                       capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    out = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert out["bad"] == 0, out
    owned = sorted(v for rk in out["ranks"] for v in rk["owned"])
    assert owned == list(range(case["n_views"]))
    assert sum(rk["send_bytes"] for rk in out["ranks"]) == sum(rk["recv_bytes"] for rk in out["ranks"]) > 0


def test_view_sharded_batched_exchange_matches_oracle(cuda):
    """Batched form of the view-sharded mode (vsb_shard_plan / pack / unpack, vsb_feed_batch, vsb_blend_batch): 4 frames per
    exchange, one packed message per peer; every frame bit-exact."""
    world = 2
    if cuda.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    case = dict(n_views=6, src_w=480, src_h=270, pano_width=1536, num_bands=4, batch=4)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
                        "--master-port", "29623", os.path.join(ROOT, "tests", "shard_gpu_worker.py"), json.dumps(case)],
                       capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    out = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert out["bad"] == 0 and out["batch"] == 4, out


@pytest.mark.parametrize("world,batch", [(2, 1), (2, 3), (4, 2)])
def test_view_sharded_native_transport_matches_oracle(cuda, world, batch):
    """The exchange inside libvsb200 (vsb_shard_init / vsb_shard_compose: ownership rule evaluated on every rank, one grouped
    ncclSend / ncclRecv per peer, three internal streams, two alternating halves of the frame slots): three overlapped submissions
    of `batch` frames on two caller streams, every frame bit-exact against oracle-G."""
    if cuda.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    case = dict(n_views=6, src_w=480, src_h=270, pano_width=1536, num_bands=4, batch=batch, native=True)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
                        "--master-port", str(29630 + world + batch), os.path.join(ROOT, "tests", "shard_gpu_worker.py"), json.dumps(case)],
                       capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    out = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert out["bad"] == 0 and out["native"], out
    owned = sorted(v for rk in out["ranks"] for v in rk["owned"])
    assert owned == list(range(case["n_views"]))
    assert sum(rk["send_bytes"] for rk in out["ranks"]) == sum(rk["recv_bytes"] for rk in out["ranks"]) > 0
