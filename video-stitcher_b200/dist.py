"""Multi-GPU host logic of the compose path (one process per GPU, torch.distributed for the plumbing).

The unit of work is a frame and frames are independent, so N GPUs shard the FRAME STREAM: rank r of W composes frames
r, r + W, r + 2W, ... with its own handle; there is no data-path collective.  The only exchanges
are control-plane: a barrier around the timed region and a MAX reduction of the per-rank device time.
No CUDA in this module: it is exercised on CPU with the gloo backend (tests/test_dist_gloo.py).
"""


def frames_of_rank(rank, world, n_frames):
    """Indices of the frames rank `rank` composes out of a stream of `n_frames` (round robin keeps per-rank latency even)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return list(range(rank, n_frames, world))


def owner_of_frame(frame, world):
    return frame % world


def ring_seed_offset(rank, ring):
    """First synthetic frame index of a rank's resident ring (bench.py): rings of different ranks never overlap."""
    return rank * ring


def job_rate(frames_per_rank, step_ms_max):
    """Whole-job frames/s: what all ranks composed divided by the SLOWEST rank's device time."""
    return sum(frames_per_rank) / (step_ms_max / 1000.0)


def reduce_step_time(ms_local, dist=None, device=None):
    """MAX over ranks of the locally measured device time (identity without a process group)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return ms_local
    import torch
    t = torch.tensor([ms_local], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
