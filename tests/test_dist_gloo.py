"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: frame sharding is a partition, the job rate uses the
slowest rank, and the reference arm only runs on rank 0."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import json, os, sys
sys.path.insert(0, sys.argv[1])
import torch, torch.distributed as dist
import vsb200
D = vsb200._load("dist")
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n_frames = 37
mine = D.frames_of_rank(rank, world, n_frames)
gathered = [None] * world
dist.all_gather_object(gathered, mine)
ms = D.reduce_step_time(10.0 + 5.0 * rank, dist)          # rank 1 is the slow one
rate = D.job_rate([len(g) for g in gathered], ms)
dist.barrier()
if rank == 0:
    print(json.dumps({"gathered": gathered, "ms": ms, "rate": rate, "ring": [D.ring_seed_offset(r, 8) for r in range(world)]}))
dist.destroy_process_group()
'''


def _torchrun(args, timeout=300):
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    return subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                           "--master-port", "29611"] + args, capture_output=True, text=True, timeout=timeout, env=env, cwd=ROOT)


def test_frame_sharding_is_a_partition_and_rate_uses_slowest_rank(tmp_path):
    w = tmp_path / "worker.py"
    w.write_text(WORKER)
    r = _torchrun([str(w), ROOT])
    assert r.returncode == 0, r.stderr[-2000:]
    out = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    a, b = out["gathered"]
    assert sorted(a + b) == list(range(37)) and not set(a) & set(b)
    assert out["ms"] == 15.0                                  # MAX over ranks
    assert abs(out["rate"] - 37 / 0.015) < 1e-6
    assert out["ring"] == [0, 8]


def test_reference_arm_runs_on_rank0_only():
    from oracle import ref as vr
    if not vr.available():
        pytest.skip("oracle/_ref not built")
    r = _torchrun(["bench.py", "--gpus", "2", "--impl", "reference", "--workload", "tiny", "--steps", "2", "--warmup", "1"], timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1 and lines[0]["impl"] == "reference" and lines[0]["n_gpus"] == 2
    assert lines[0]["cpu_baseline"]["kind"] == "reference" and lines[0]["value"] > 0
    assert lines[0]["e2e"]["h2d_bytes_per_step"] == 0
