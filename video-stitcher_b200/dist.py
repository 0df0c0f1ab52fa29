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


# ---------------------------------------------------------------------------------------------------------------------
# View-sharded mode: rank r composes its views' front half and its canvas strip's back half; in between
# the ranks exchange the Gaussian sub-planes (u8) that foreign strips read.  The plan is a pure function of the static
# calibration tables (vsb_shard_rect), identical on every rank; the transport is grouped point-to-point over
# torch.distributed (NCCL send/recv on NVLink for CUDA tensors, gloo for the CPU tests).

def build_exchange_plan(rect_fn, owners, rank, world, levels):
    """rect_fn(dst_rank, view, level) -> (x0, y0, w, h) of what dst_rank reads of that plane (w == 0: nothing).
    Returns (sends, recvs): lists of (view, level, rect, peer), both sorted by (peer, view, level) so that the k-th send
    to a peer pairs with that peer's k-th receive."""
    sends, recvs = [], []
    for v, owner in enumerate(owners):
        for k in levels:
            if owner == rank:
                for dst in range(world):
                    if dst != rank:
                        r = tuple(rect_fn(dst, v, k))
                        if r[2] > 0 and r[3] > 0:
                            sends.append((v, k, r, dst))
            else:
                r = tuple(rect_fn(rank, v, k))
                if r[2] > 0 and r[3] > 0:
                    recvs.append((v, k, r, owner))
    key = lambda e: (e[3], e[0], e[1])
    return sorted(sends, key=key), sorted(recvs, key=key)


def exchange_bytes(plan_half):
    return sum(3 * r[2] * r[3] for _, _, r, _ in plan_half)


def run_exchange(dist, torch, plane_fn, sends, recvs):
    """plane_fn(view, level) -> uint8 tensor [3, h, w] aliasing that Gaussian plane.  One grouped batch of isend / irecv,
    then the received blocks are written into the local copies of the foreign planes."""
    ops, landing = [], []
    for v, k, (x0, y0, w, h), dst in sends:
        ops.append(dist.P2POp(dist.isend, plane_fn(v, k)[:, y0:y0 + h, x0:x0 + w].contiguous(), dst))
    for v, k, (x0, y0, w, h), src in recvs:
        p = plane_fn(v, k)
        buf = torch.empty((3, h, w), dtype=p.dtype, device=p.device)
        ops.append(dist.P2POp(dist.irecv, buf, src))
        landing.append((p, x0, y0, w, h, buf))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    for p, x0, y0, w, h, buf in landing:
        p[:, y0:y0 + h, x0:x0 + w].copy_(buf)


def contiguous_runs(views):
    """[0, 1, 5] -> [(0, 2), (5, 6)]: half-open runs of consecutive view indices (one vsb_feed_batch call each)."""
    runs = []
    for v in sorted(views):
        if runs and runs[-1][1] == v:
            runs[-1] = (runs[-1][0], v + 1)
        else:
            runs.append((v, v + 1))
    return runs


class _DeviceArray:
    """Zero-copy view of handle-owned device memory for torch.as_tensor (CUDA array interface, version 2)."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "|u1", "data": (int(ptr), False), "version": 2, "strides": None}


class ShardedStitcher:
    """One per rank, around a calibrated binding.Stitcher (every rank calibrates the same rig)."""

    def __init__(self, st, dist, torch):
        self.st, self.dist, self.torch = st, dist, torch
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        st.shard_set(self.rank, self.world)
        self.strip_x0, self.strip_x1, self.owned = st.shard_info()
        gathered = [None] * self.world
        dist.all_gather_object(gathered, self.owned)
        self.owners = [None] * st.num_views
        for r, views in enumerate(gathered):
            for v in views:
                self.owners[v] = r
        assert None not in self.owners, "every view needs exactly one owner"
        _, _, nb = st.get_roi()
        self.levels = list(range(nb + 1))
        self.sends, self.recvs = build_exchange_plan(st.shard_rect, self.owners, self.rank, self.world, self.levels)
        self._planes = {}

    def plane(self, view, level):
        key = (view, level)
        if key not in self._planes:
            ptr, w, h = self.st.get_plane(view, level, 0)
            self._planes[key] = self.torch.as_tensor(_DeviceArray(ptr, (3, h, w)), device="cuda")
        return self._planes[key]

    # ---- batched form: F frames per exchange, one packed message per peer (vsb_shard_plan / pack / unpack)
    def _batch_setup(self, n_frames):
        if getattr(self, "_batch_frames", 0) >= n_frames:
            return
        self.st.shard_plan(self.owners)
        self.runs = contiguous_runs(self.owned)
        self._send, self._recv = {}, {}
        for peer in range(self.world):
            if peer == self.rank:
                continue
            sb, rb = self.st.shard_peer_bytes(peer)
            if sb:
                self._send[peer] = (sb, self.torch.empty(sb * n_frames, dtype=self.torch.uint8, device="cuda"))
            if rb:
                self._recv[peer] = (rb, self.torch.empty(rb * n_frames, dtype=self.torch.uint8, device="cuda"))
        self._batch_frames = n_frames

    def compose_batch(self, src_ptrs_per_frame, src_pitch, out_ptrs, out_pitch, stream):
        """src_ptrs_per_frame[f][v]: device BGR frame of view v, frame f (only the owned views are read); out_ptrs[f]: full-size
        CV_16SC3 buffers, this rank writes its strip of each."""
        F = len(out_ptrs)
        self._batch_setup(F)
        for v0, v1 in self.runs:
            self.st.feed_batch(v0, v1, F, [src_ptrs_per_frame[f][v] for f in range(F) for v in range(v0, v1)], src_pitch, stream)
        ops = []
        for peer, (sb, buf) in self._send.items():
            self.st.shard_pack(peer, F, buf.data_ptr(), stream)
            ops.append(self.dist.P2POp(self.dist.isend, buf[:sb * F], peer))
        for peer, (rb, buf) in self._recv.items():
            ops.append(self.dist.P2POp(self.dist.irecv, buf[:rb * F], peer))
        if ops:
            for req in self.dist.batch_isend_irecv(ops):
                req.wait()
        for peer, (rb, buf) in self._recv.items():
            self.st.shard_unpack(peer, F, buf.data_ptr(), stream)
        self.st.blend_batch(out_ptrs, out_pitch, stream)

    def compose(self, src_ptrs, src_pitch, out_ptr, out_pitch, stream):
        """src_ptrs: device BGR frames of ALL views (only the owned ones are read).  Writes this rank's strip of the
        CV_16SC3 panorama into out_ptr (full-size buffer)."""
        for v in self.owned:
            self.st.feed(v, src_ptrs[v], src_pitch, stream)
        run_exchange(self.dist, self.torch, self.plane, self.sends, self.recvs)
        self.st.blend(out_ptr, out_pitch, stream)
