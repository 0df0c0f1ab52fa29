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


SHARD_WORKER = r'''
import json, sys
sys.path.insert(0, sys.argv[1])
import torch, torch.distributed as dist
import vsb200
D = vsb200._load("dist")
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n_views, levels, owners = 4, [0, 1, 2], [0, 0, 1, 1]
def shape(v, k): return (3, 40 >> k, (96 + 16 * v) >> k)
def rect_fn(dst, v, k):       # what rank dst reads of plane (v, k): a deterministic pseudo-geometry, empty for some pairs
    _, h, w = shape(v, k)
    if (v + k + dst) % 3 == 0: return (0, 0, 0, 0)
    x0 = (7 * v + 3 * dst) % (w // 2); y0 = (5 * k + dst) % (h // 2)
    return (x0, y0, w // 2 - 1, h // 2 - 1)
def truth(v, k):              # what the owner computed
    c, h, w = shape(v, k)
    return ((torch.arange(c * h * w).reshape(c, h, w) * (v + 2) + k) % 251).to(torch.uint8)
planes = {(v, k): (truth(v, k) if owners[v] == rank else torch.full(shape(v, k), 255, dtype=torch.uint8)) for v in range(n_views) for k in levels}
sends, recvs = D.build_exchange_plan(rect_fn, owners, rank, world, levels)
D.run_exchange(dist, torch, lambda v, k: planes[(v, k)], sends, recvs)
ok = True
for v in range(n_views):
    for k in levels:
        if owners[v] != rank:
            x0, y0, w, h = rect_fn(rank, v, k)
            t = truth(v, k)
            ok &= bool((planes[(v, k)][:, y0:y0 + h, x0:x0 + w] == t[:, y0:y0 + h, x0:x0 + w]).all())
            outside = planes[(v, k)].clone(); outside[:, y0:y0 + h, x0:x0 + w] = 255
            ok &= bool((outside == 255).all())          # nothing outside the planned rectangle was touched
        else:
            ok &= bool((planes[(v, k)] == truth(v, k)).all())
all_plans = [None] * world
dist.all_gather_object(all_plans, (sends, recvs))
dist.barrier()
if rank == 0:
    s0, r0 = all_plans[0]; s1, r1 = all_plans[1]
    sym = [(v, k, r) for v, k, r, _ in s0] == [(v, k, r) for v, k, r, _ in r1] and [(v, k, r) for v, k, r, _ in s1] == [(v, k, r) for v, k, r, _ in r0]
    print(json.dumps({"ok0": ok, "symmetric": sym, "bytes": [D.exchange_bytes(s0), D.exchange_bytes(s1)]}))
else:
    assert ok
dist.destroy_process_group()
'''


def test_view_shard_exchange_plan_and_transport(tmp_path):
    """The view-sharded mode's host logic: the plan one rank sends is the plan the peer receives, and the grouped
    point-to-point transport lands every block in the right sub-rectangle of the right plane (gloo, world size 2)."""
    w = tmp_path / "shard_worker.py"
    w.write_text(SHARD_WORKER)
    r = _torchrun([str(w), ROOT])
    assert r.returncode == 0, r.stderr[-3000:]
    out = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert out["ok0"] and out["symmetric"] and min(out["bytes"]) > 0


def test_contiguous_runs_of_owned_views():
    import vsb200
    D = vsb200.dist
    assert D.contiguous_runs([0, 1, 5]) == [(0, 2), (5, 6)]
    assert D.contiguous_runs([4, 3, 5]) == [(3, 6)]
    assert D.contiguous_runs([]) == []
