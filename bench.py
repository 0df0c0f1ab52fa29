#!/usr/bin/env python
"""bench.py -- stitched frames/s of the per-frame 360-degree compose path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch F] [--workload cfg2] [--impl ours|reference]

A step = one vsb_compose submission of F frames (default 16) of the workload (BASELINE.json configs[1]: 6 x 1080p ->
3840-wide spherical panorama, CPW mesh remap on, 5-band blend).  Source frames live in a ring of frame sets larger than
L2, so every step reads its inputs from HBM.
  value        frames/s with inputs resident in HBM; CUDA events on the launching stream; max over ranks
  e2e          frames/s through vsb_compose_host: pinned HOST frames in, host panoramas out, H2D + D2H inside the timed region
  roofline     the dominant kernel: algorithmic bytes / mean device time, from CUDA events around every kernel of the
               timed submissions (vsb_set_profiling); `traffic` = DRAM bytes of the committed ncu capture (profiles/)
  roofline_path  the whole path against B_io = sources once + CV_16SC3 panorama once
  cpu_baseline the reference's own vendored OpenCV 3.4.0 CPU code (oracle/_ref, kind "reference") -- or the oracle-G port
               when that library is absent -- timed on this box's host cores on a bounded sample of the same workload
N > 1 (torchrun): every rank composes its own frame stream (frame-level replicas, no data-path collective; "weak").
--impl reference : the CPU implementation alone, all host threads, same config / metric (rank 0 only).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1] (and [4] without the recalibration thread)
    "cfg2": dict(n_views=6, src_w=1920, src_h=1080, pano_width=3840, num_bands=5, enable_local=True, projection=0,
                 name="6x1080p->3840 spherical, CPW on, 5 bands"),
    "cfg3": dict(n_views=6, src_w=1920, src_h=1080, pano_width=7680, num_bands=5, enable_local=True, projection=0,
                 name="6x1080p->7680 spherical, CPW on, 5 bands"),
    # BASELINE.json configs[4]: config 2 with the recalibration thread publishing a new CPW mesh while frames are composed
    # (A/timed.cpp:414-463 re-solves every RECALIB_DEL = 1000 ms; here every `recalib_ms` to stress the overlap)
    "cfg5": dict(n_views=6, src_w=1920, src_h=1080, pano_width=3840, num_bands=5, enable_local=True, projection=0, recalib_ms=20,
                 name="6x1080p->3840 spherical, CPW on, 5 bands, mesh re-installed every 20 ms by a second host thread"),
    "cfg4": dict(n_views=12, src_w=3840, src_h=2160, pano_width=15360, num_bands=5, enable_local=True, projection=0,
                 name="12x2160p->15360 spherical, CPW on, 5 bands"),
    "cfg1": dict(n_views=2, src_w=1280, src_h=720, pano_width=4021, num_bands=5, enable_local=False, projection=0,
                 name="2x720p->4021 spherical, CPW off, 5 bands"),
    "tiny": dict(n_views=4, src_w=320, src_h=240, pano_width=1024, num_bands=3, enable_local=True, projection=0,
                 name="4x320x240->1024 (debug)"),
}
RING = 8  # distinct frame sets resident in HBM (8 x 37 MB = 299 MB > 126 MB L2 at cfg2)
METRIC = "stitched equirect frames/sec"
CTL_DEV = "cpu"  # device of the control-plane tensors (gloo: cpu; --ctl-backend nccl: cuda)
_OUT = None


def emit(line):
    """The ONE JSON line of the run, written to the process's original stdout.  main() points file descriptor 1 at stderr for
    everything else, so that what a library prints there (NCCL's version banner at communicator init) cannot land next to it."""
    out = _OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu, self.rows, self._stop_evt, self.proc = gpu_index, [], threading.Event(), None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append((time.time(), [c.strip() for c in line.split(",")]))
                if self._stop_evt.is_set():
                    break
        except Exception:
            pass

    def stop(self):
        self._stop_evt.set()
        if self.proc:
            try:
                self.proc.terminate()
            except Exception:
                pass

    def summary(self, t0, t1):
        rows = [r for t, r in self.rows if t0 <= t <= t1] or [r for _, r in self.rows[-3:]]
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def make_frames(cfg, n_sets, first=0):
    import vsb200
    S = vsb200.synth
    return [[S.frame(i, first + f, cfg["src_w"], cfg["src_h"]) for i in range(cfg["n_views"])] for f in range(n_sets)]


def cpu_compose_run(cfg, frames, budget_s, threads):
    """Times the CPU implementation of the path on `threads` host threads for about `budget_s` seconds.

    Preferred: oracle/_ref (the reference's vendored OpenCV 3.4.0 CPU primitives + the CPU branch of its MultiBandBlender),
    in its two honest configurations -- views in sequence with OpenCV's own parallel_for_ threads, and one thread per view
    (cv::pyrDown / pyrUp are single-threaded in this OpenCV) -- the faster one is reported.  Fallback: the oracle-G port.
    Static inputs (maps, seam masks, mesh maps, gains) are the oracle's, shared by every implementation."""
    import vsb200
    from oracle import oracle as og
    from oracle import pipeline as op
    from oracle import ref as vr
    og.set_num_threads(threads)
    rig = op.OracleRig(cfg["n_views"], cfg["src_w"], cfg["src_h"], cfg["pano_width"], cfg["projection"], cfg["num_bands"],
                       cfg["enable_local"], vsb200.synth.gains(cfg["n_views"]))
    if cfg["enable_local"]:
        for i in range(cfg["n_views"]):
            rig.set_mesh(i, *vsb200.synth.mesh(*rig.sizes[i]))
    runs = []
    if vr.available():
        cpu = vr.RigC(rig)
        modes = [("views in sequence, OpenCV parallel_for_ threads", False), ("one thread per view (OpenMP)", True)]
        kind = "reference"
        def step(k, par):
            cpu.compose(frames[k % len(frames)], parallel_views=par)
    else:
        modes = [("oracle-G C port, OpenMP over rows", False)]
        kind = "port"
        def step(k, par):
            rig.compose(frames[k % len(frames)])
    for label, par in modes:
        if kind == "reference":
            vr.set_num_threads(1 if par else threads)
        step(0, par)  # warm-up (allocations, thread pool)
        t0 = time.perf_counter()
        n = 0
        while True:
            step(n, par)
            n += 1
            dt = time.perf_counter() - t0
            if dt >= budget_s / len(modes) or n >= 400:
                break
        runs.append((n / dt, n, dt, label))
    best = max(runs)
    used = min(threads, cfg["n_views"]) if best[3].startswith("one thread") else threads
    sample = (f"{best[1]} frames of the full {cfg['name']} workload in {best[2]:.1f} s; {best[3]}; "
              + "; ".join(f"{lab}: {fps:.2f} fps" for fps, _, _, lab in runs))
    return best[0], kind, used, sample


def run_reference(args, cfg, rank, world):
    if rank != 0:
        return
    threads = host_cores()
    frames = make_frames(cfg, 4)
    budget = min(90.0, max(8.0, 2.0 * max(1, args.steps) / 10.0))
    fps, kind, used, sample = cpu_compose_run(cfg, frames, budget, threads)
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 / fps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8/s16 (fp32 taps)", "data": "synthetic",
        "config": {"workload": cfg["name"], "frames_per_step": 1, "inputs": "host memory", "host_threads": threads},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": used, "kind": kind, "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def timed_repeats(step, K, sync, repeats=5, min_total_s=2.0, probe=None):
    """Times `repeats` runs of (at least) K steps each with CUDA events on the launching stream; every run is bracketed by sync()
    (barrier + cudaDeviceSynchronize).  Runs are lengthened beyond K steps until the whole timed region covers >= min_total_s
    (>= 2 s, median of 5 repeats), independent of the --steps the caller passed.  Returns (median ms per step,
    all ms per step, steps per run).  `probe` = ms per step from a previous measurement (sizes the runs without a probe run)."""
    import torch
    if probe is None:
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(K):
            step(k)
        e1.record()
        sync()
        probe = e0.elapsed_time(e1) / K
    per_run = max(K, int(min_total_s * 1000.0 / repeats / max(probe, 1e-3)) + 1)
    out = []
    for r in range(repeats):
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(per_run):
            step(k)
        e1.record()
        sync()
        out.append(e0.elapsed_time(e1) / per_run)
    return sorted(out)[len(out) // 2], out, per_run


def hbm_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured)"
    except Exception:
        return 6650.0, "B200_PROFILING.md fallback 6650 GB/s"


def run_replicas_phase(dist, torch, rank, world, local_rank):
    """Frame-level replicas of the headline single-GPU workload: every rank composes its own frame stream, no data-path
    collective (weak scaling).  The secondary number of an N > 1 run -- and its fallback `value` if the sharded attempt fails."""
    import vsb200
    S, D = vsb200.synth, vsb200.dist
    rcfg = WORKLOADS["cfg2"]
    F = 16
    st2, _ = make_rig(rcfg, F)
    roi2, _, _ = st2.get_roi()
    op2 = (roi2[2] * 6 + 255) // 256 * 256
    n_sets = F
    sets2 = [[torch.from_numpy(S.frame(i, f + D.ring_seed_offset(rank, n_sets), rcfg["src_w"], rcfg["src_h"])).cuda() for i in range(rcfg["n_views"])] for f in range(n_sets)]
    outs2 = [torch.empty((roi2[3], op2 // 2), dtype=torch.int16, device="cuda") for _ in range(F)]
    stream = torch.cuda.current_stream().cuda_stream
    c2 = [st2.make_compose_call([sets2[(s0 + j) % n_sets][i].data_ptr() for j in range(F) for i in range(rcfg["n_views"])], rcfg["src_w"] * 3,
                                [o.data_ptr() for o in outs2], op2, stream) for s0 in range(n_sets)]
    for w in range(3):
        c2[w]()
    def sync2():
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    m2, runs, per_run = timed_repeats(lambda k: c2[k % n_sets](), 10, sync2, repeats=3, min_total_s=0.9)
    m2 = D.reduce_step_time(m2, dist, CTL_DEV)
    launches = st2.last_launch_count()
    st2.close()
    del sets2, outs2
    torch.cuda.empty_cache()
    return {"value": world * F / (m2 / 1000.0), "unit": "frames/s", "workload": rcfg["name"], "frames_per_step": F, "ms_per_step": m2,
            "launches_per_step": launches * world, "steps_per_run": per_run,
            "note": "frame-level replicas: every rank composes its own frame stream, no data-path collective (weak scaling)"}


class Watchdog(threading.Thread):
    """An N > 1 run must end within the driver's limit even if a rank dies inside a collective or a peer-to-peer exchange (the
    other ranks then spin in a kernel forever).  Every rank runs one: when the deadline passes, or when any rank has dropped a
    failure marker, rank 0 prints the fallback line (the replicas measurement, clearly labelled) and every rank leaves with
    os._exit(0)."""

    def __init__(self, rank, deadline_s, marker, fallback_line):
        super().__init__(daemon=True)
        self.rank, self.t_end, self.marker, self.fallback_line, self.done = rank, time.time() + deadline_s, marker, fallback_line, threading.Event()

    def fail(self, why):
        try:
            with open(self.marker, "a") as fh:
                fh.write(f"rank {self.rank}: {why}\n")
        except Exception:
            pass

    def run(self):
        while not self.done.wait(0.5):
            why = None
            if os.path.exists(self.marker):
                try:
                    why = open(self.marker).read().strip().replace("\n", " | ")[:300]
                except Exception:
                    why = "a rank failed"
            elif time.time() > self.t_end:
                why = "deadline passed (a rank hangs)"
            if why:
                if self.rank == 0:
                    time.sleep(1.0)  # let the failing rank finish writing its reason
                    line = self.fallback_line(why)
                    if line:
                        emit(line)
                os._exit(0)


PARITY_RIG = dict(n_views=6, src_w=480, src_h=270, pano_width=1536, num_bands=4, enable_local=True, projection=0)  # tests/golden: "shard6"


def make_rig(cfg, max_batch):
    import vsb200
    B, S = vsb200.binding, vsb200.synth
    n = cfg["n_views"]
    st = B.Stitcher(n, cfg["num_bands"], cfg["enable_local"], max_batch)
    st.calibrate_rig(cfg["projection"], cfg["pano_width"], cfg["src_w"], cfg["src_h"], 90.0, S.gains(n))
    info = st.rig_info()
    if cfg["enable_local"]:
        for i in range(n):
            mx, my = S.mesh(info.view_roi[i][2], info.view_roi[i][3])
            st.set_mesh(i, mx.ctypes.data, my.ctypes.data, mx.shape[0], mx.shape[1])
    return st, info


def shard_unique_id(dist, rank):
    import vsb200
    box = [vsb200.binding.shard_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    return box[0]


def sharded_parity_check(dist, torch, rank, world):
    """Composes the small 'shard6' rig view-sharded over all ranks (the same library path the timed run uses), sums the disjoint
    strips and compares the SHA-256 of the CV_16SC3 panorama with the committed oracle-G hash (tests/golden/oracle_compose_hashes.json,
    generator tests/golden/make_compose_hashes.py).  Nothing under oracle/ is touched here."""
    import hashlib
    import vsb200
    S = vsb200.synth
    cfg = PARITY_RIG
    st, _ = make_rig(cfg, 2)
    st.shard_init(rank, world, shard_unique_id(dist, rank))
    roi, _, _ = st.get_roi()
    W, H = roi[2], roi[3]
    _, _, owned = st.shard_info()
    srcs = [torch.from_numpy(S.frame(i, 0, cfg["src_w"], cfg["src_h"])).cuda() if i in owned else None for i in range(cfg["n_views"])]
    out = torch.zeros((H, W, 3), dtype=torch.int16, device="cuda")
    st.shard_compose([t.data_ptr() if t is not None else 0 for t in srcs], cfg["src_w"] * 3, [out.data_ptr()], W * 6, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    full = out.to(torch.int32).to(CTL_DEV)
    dist.all_reduce(full)  # strips are disjoint and the rest of every rank's buffer is zero
    got = full.to(torch.int16).cpu().numpy()
    want = json.load(open(os.path.join(ROOT, "tests", "golden", "oracle_compose_hashes.json")))["shard6"]
    sha = hashlib.sha256(got.tobytes()).hexdigest()
    st.close()
    return {"rig": "6x480x270->1536 spherical, CPW on, 4 bands (tests/golden shard6)", "sha256": sha, "expected": want["sha256"],
            "bit_exact": sha == want["sha256"] and list(got.shape) == want["shape"]}


def run_sharded(args, cfg, rank, world, local_rank):
    """View-sharded mode (the north-star split): all ranks work on ONE frame stream.  Rank r remaps / builds the
    pyramids of its views, the ranks exchange the Gaussian u8 sub-planes foreign strips read (one grouped ncclSend / ncclRecv per
    peer inside libvsb200), rank r blends its canvas strip.  --batch frames per exchange; submissions alternate between two caller
    streams so that exchange k overlaps front half k + 1."""
    import torch
    import torch.distributed as dist
    import vsb200
    B, S, D = vsb200.binding, vsb200.synth, vsb200.dist
    if not dist.is_initialized():
        dist.init_process_group("gloo", init_method="tcp://127.0.0.1:29655", rank=0, world_size=1)
    n, K, W_, F = cfg["n_views"], args.steps, max(args.warmup, 3), max(1, args.batch)
    replicas = None if args.no_replicas else run_replicas_phase(dist, torch, rank, world, local_rank)
    peak, peak_src = hbm_peak()

    def fallback_line(why):  # what rank 0 prints when the view-sharded attempt does not complete: the replicas measurement, labelled as such
        if replicas is None:
            return {"metric": METRIC, "value": None, "unit": "frames/s", "n_gpus": world, "error": "view-sharded run failed: " + why}
        rc = WORKLOADS["cfg2"]
        b2 = rc["n_views"] * rc["src_w"] * rc["src_h"] * 3 + 3839 * 627 * 6
        return {"metric": METRIC, "value": replicas["value"], "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": W_,
                "ms_per_step": replicas["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u8/s16 (fp32 taps)", "data": "synthetic",
                "config": {"workload": rc["name"], "frames_per_step": replicas["frames_per_step"],
                           "multi_gpu": "FALLBACK: frame-level replicas, no collective -- the view-sharded run of '" + cfg["name"] + "' did not complete: " + why},
                "e2e": None, "gpu_launches": replicas["launches_per_step"] * K, "parity_checked": False,
                "roofline_path": {"alg_bytes_per_frame": b2, "achieved": b2 * replicas["value"] / world / 1e9, "peak": peak, "unit": "GB/s",
                                  "frac": b2 * replicas["value"] / world / 1e9 / peak}, "replicas": replicas, "sharded_error": why}

    marker = os.path.join("/tmp", f"vsb_bench_fail_{os.environ.get('MASTER_PORT', '0')}_{os.getppid()}")
    if rank == 0 and os.path.exists(marker):
        os.remove(marker)
    dist.barrier(); torch.cuda.synchronize()
    dog = Watchdog(rank, args.shard_deadline, marker, fallback_line)
    dog.start()
    try:
        _run_sharded_body(args, cfg, rank, world, local_rank, replicas, dog)
    except BaseException as e:  # noqa: the process must not linger inside torchrun with peers spinning
        dog.fail(f"{type(e).__name__}: {str(e)[:200]}")
        time.sleep(30)  # the watchdog threads (this rank's included) pick the marker up and end the run
        os._exit(0)


def _run_sharded_body(args, cfg, rank, world, local_rank, replicas, dog):
    import torch
    import torch.distributed as dist
    import vsb200
    B, S, D = vsb200.binding, vsb200.synth, vsb200.dist
    n, K, W_, F = cfg["n_views"], args.steps, max(args.warmup, 3), max(1, args.batch)
    parity = sharded_parity_check(dist, torch, rank, world)
    if os.environ.get("VSB_BENCH_INJECT_FAIL") and rank == world - 1:
        raise RuntimeError("injected failure (VSB_BENCH_INJECT_FAIL): exercises the watchdog / fallback line")
    st, info = make_rig(cfg, 2 * F if 2 * F <= 16 else F)  # two submissions in flight when both fit the handle's frame slots (max 16)
    st.shard_init(rank, world, shard_unique_id(dist, rank))
    roi, _, nb = st.get_roi()
    OW, OH = roi[2], roi[3]
    x0, x1, owned = st.shard_info()
    n_sets = max(2, min(args.ring, RING))
    host_sets = [[torch.from_numpy(S.frame(i, f, cfg["src_w"], cfg["src_h"])).pin_memory() if i in owned else None for i in range(n)] for f in range(n_sets)]
    dev_sets = [[t.cuda() if t is not None else None for t in fs] for fs in host_sets]
    out_pitch = (OW * 6 + 255) // 256 * 256
    outs = [[torch.zeros((OH, out_pitch // 2), dtype=torch.int16, device="cuda") for _ in range(F)] for _ in range(2)]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    main = torch.cuda.current_stream()
    calls = {}
    for par in range(2):
        for s0 in range(n_sets):
            srcs = [(dev_sets[(s0 + j) % n_sets][i].data_ptr() if i in owned else 0) for j in range(F) for i in range(n)]
            calls[(par, s0)] = st.make_shard_compose_call(srcs, cfg["src_w"] * 3, [o.data_ptr() for o in outs[par]], out_pitch, streams[par].cuda_stream)

    def step(k):
        calls[(k & 1, (k * F) % n_sets)]()

    def sync():
        torch.cuda.synchronize()  # drain the library's own NCCL traffic before torch's communicator runs its barrier
        dist.barrier()
        torch.cuda.synchronize()

    class Timer:  # events on the main stream around work that runs on the two caller streams
        def __enter__(self):
            self.e0, self.e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.e0.record(main)
            for s_ in streams:
                s_.wait_stream(main)
            return self

        def __exit__(self, *a):
            for s_ in streams:
                main.wait_stream(s_)
            self.e1.record(main)

    for w in range(W_):
        step(w)
    launches = st.last_launch_count()
    sync()
    sampler = ClockSampler(local_rank); sampler.start(); time.sleep(0.3)
    t0 = time.time()
    with Timer() as tm:
        for k in range(K):
            step(k)
    sync()
    probe = D.reduce_step_time(tm.e0.elapsed_time(tm.e1), dist, CTL_DEV) / K
    per_run = max(K, int(2000.0 / 5 / max(probe, 1e-3)) + 1)
    runs = []
    for r in range(5):
        sync()
        with Timer() as tm:
            for k in range(per_run):
                step(k)
        sync()
        runs.append(D.reduce_step_time(tm.e0.elapsed_time(tm.e1), dist, CTL_DEV) / per_run)  # MAX over ranks
    t1 = time.time()
    sampler.stop()
    ms = sorted(runs)[2]
    fps = F / (ms / 1000.0)

    # ---- e2e: pinned host frames of the owned views in, this rank's strip of the host panorama out, every step
    e2e = None
    if not args.no_e2e:
        xe = max(x0, min(x1, OW))  # the strip is cut from the padded canvas; the panorama ends at OW
        h_out = [torch.zeros((OH, 3 * (xe - x0)), dtype=torch.int16).pin_memory() for _ in range(F)]
        strip_dev = [torch.zeros((OH, 3 * (xe - x0)), dtype=torch.int16, device="cuda") for _ in range(F)]
        stage = [[torch.empty_like(host_sets[0][i], device="cuda") if i in owned else None for i in range(n)] for _ in range(F)]
        e2e_call = st.make_shard_compose_call([(stage[j][i].data_ptr() if i in owned else 0) for j in range(F) for i in range(n)], cfg["src_w"] * 3,
                                              [o.data_ptr() for o in outs[0]], out_pitch, main.cuda_stream)
        def host_step(k):
            for j in range(F):
                for i in owned:
                    stage[j][i].copy_(host_sets[(k * F + j) % n_sets][i], non_blocking=True)
            e2e_call()
            for j in range(F):  # the strip leaves as ONE contiguous DMA (a strided 2-D copy runs at a fraction of the link rate)
                strip_dev[j].copy_(outs[0][j][:, 3 * x0:3 * xe])
                h_out[j].copy_(strip_dev[j], non_blocking=True)
        for w in range(2):
            host_step(w)
        Ke = max(3, min(K, 20))
        sync()
        tw = time.perf_counter()
        for k in range(Ke):
            host_step(k)
        torch.cuda.synchronize()
        dt = D.reduce_step_time(time.perf_counter() - tw, dist, CTL_DEV)
        bytes_in = sum(host_sets[0][i].numel() for i in owned) * F
        tot = torch.tensor([float(bytes_in), float(OH * max(0, xe - x0) * 6 * F)], dtype=torch.float64, device=CTL_DEV)
        dist.all_reduce(tot)
        e2e = {"value": F * Ke / dt, "unit": "frames/s", "h2d_bytes_per_step": int(tot[0].item()), "d2h_bytes_per_step": int(tot[1].item()), "steps": Ke,
               "api": "per rank: pinned host frames of the owned views -> device, vsb_shard_compose, the rank's strip of the panorama -> host"}

    sb, rb = st.shard_exchange_bytes()
    stats = [None] * world
    dist.all_gather_object(stats, {"rank": rank, "views": owned, "strip": [x0, x1], "send_bytes_per_frame": sb, "recv_bytes_per_frame": rb, "launches_per_step": launches})
    st.close()
    del dev_sets, outs
    torch.cuda.empty_cache()

    # ---- secondary numbers: (a) ONE GPU on the same workload through the same handle type (what the sharded rate is a speed-up of),
    #      (b) frame-level replicas of the headline single-GPU workload (every rank its own stream, no collective)
    single = None
    if rank == 0 and not args.no_single:
        try:
            Fs = min(F, 4) if cfg["n_views"] > 8 else F
            st1, _ = make_rig(cfg, Fs)
            sets1 = [[torch.from_numpy(S.frame(i, f, cfg["src_w"], cfg["src_h"])).cuda() for i in range(n)] for f in range(max(2, min(n_sets, 4)))]
            outs1 = [torch.zeros((OH, out_pitch // 2), dtype=torch.int16, device="cuda") for _ in range(Fs)]
            c1 = [st1.make_compose_call([sets1[(s0 + j) % len(sets1)][i].data_ptr() for j in range(Fs) for i in range(n)], cfg["src_w"] * 3,
                                        [o.data_ptr() for o in outs1], out_pitch, main.cuda_stream) for s0 in range(len(sets1))]
            for w in range(3):
                c1[w % len(c1)]()
            m1, _, _ = timed_repeats(lambda k: c1[k % len(c1)](), max(3, K // 4), torch.cuda.synchronize, repeats=3, min_total_s=0.6)
            single = {"value": Fs / (m1 / 1000.0), "unit": "frames/s", "frames_per_step": Fs, "note": "rank 0 alone, vsb_compose, same workload"}
            st1.close()
            del sets1, outs1
            torch.cuda.empty_cache()
        except Exception as e:  # e.g. the full rig does not fit next to the sharded buffers
            single = {"error": str(e)[:200]}
    dist.barrier()
    if rank == 0:
        b_io = n * cfg["src_w"] * cfg["src_h"] * 3 + OW * OH * 6
        peak, peak_src = hbm_peak()
        xbytes = sum(s_["send_bytes_per_frame"] for s_ in stats)
        emit({
            "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": W_, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8/s16 (fp32 taps)", "data": "synthetic",
            "config": {"workload": cfg["name"], "frames_per_step": F, "pano": f"{OW}x{OH} CV_16SC3", "bands": nb,
                       "multi_gpu": "view-sharded (north-star split): views + canvas strips per rank, ONE exchange of Gaussian u8 sub-planes per submission "
                                    "(grouped ncclSend/ncclRecv inside libvsb200, one packed message per peer), exchange k overlapped with front half k+1",
                       "control_plane": f"torch.distributed {args.ctl_backend} (barriers / timing reductions only); data path: the library's own NCCL communicator",
                       "l2_policy": f"ring of {n_sets} frame sets per owned view", "timing": f"median of 5 runs of {per_run} steps (>= 2 s in total), CUDA events, max over ranks"},
            "clocks": sampler.summary(t0, t1), "e2e": e2e, "gpu_launches": sum(s_["launches_per_step"] for s_ in stats) * K, "launches_per_step": sum(s_["launches_per_step"] for s_ in stats),
            "ms_per_frame": ms / F, "runs_ms_per_step": runs,
            "roofline_path": {"alg_bytes_per_frame": b_io, "achieved": b_io * fps / world / 1e9, "peak": peak, "unit": "GB/s", "frac": b_io * fps / world / 1e9 / peak,
                              "peak_source": peak_src, "note": "per GPU: B_io x frames/s / n_gpus"},
            "parity_checked": bool(parity["bit_exact"]), "parity": parity,
            "shards": stats, "exchange_bytes_per_frame": xbytes,
            "exchange": {"bytes_per_frame": xbytes, "nvlink_GBps_per_gpu": xbytes * fps / world / 1e9},
            "single_gpu_same_workload": single, "replicas": replicas,
            "scale_note": ("BASELINE.json names a different workload per GPU count (config 2 on 1 GPU, config 3 = 4x the panorama pixels on 2 and 4, "
                           "config 4 = 16x on 8): `value` is ONE frame stream of THIS line's workload sharded over all ranks (strong scaling). Its one-GPU "
                           "counterpart is `single_gpu_same_workload` (measured in this run on rank 0), not the N=1 line of config 2; the frame-level-"
                           "replicas rate of config 2 -- comparable with the N=1 line -- is under `replicas`.")})
    dog.done.set()
    torch.cuda.synchronize()
    os._exit(0)  # (a clean communicator teardown can itself wait on peers; everything is measured and printed)


def ncu_traffic(kernel, frames_per_launch):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture (profiles/ncu_traffic.json), scaled to
    this run's frames per launch; None when no capture of that kernel is on file."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        e = t["kernels"][kernel]
        return e["dram_bytes_per_launch"] * frames_per_launch / e["frames_per_launch"]
    except Exception:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--batch", type=int, default=16, help="frames per vsb_compose submission (F; default 16 = the handle's frame slots)")
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS),
                    help="default: cfg2 on one GPU; view-sharded runs take the configurations BASELINE.json names for N GPUs "
                         "(cfg3 = 7680-wide on 2 and 4, cfg4 = 12 x 4K -> 15360 on 8)")
    ap.add_argument("--ring", type=int, default=RING, help="distinct frame sets resident in HBM (inputs must exceed L2: 8 at cfg2, 2 suffice at cfg4)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=16.0, help="CPU baseline sample budget")
    ap.add_argument("--mode", default=None, choices=["replicas", "shard"],
                    help="N > 1: shard (default) = ONE frame stream, views and canvas strips split across ranks with one exchange of "
                         "Gaussian sub-planes per submission (the north-star split, SURVEY.md 8e); replicas = every rank composes its own frames")
    ap.add_argument("--no-replicas", action="store_true", help="shard mode: skip the secondary frame-level-replicas number")
    ap.add_argument("--no-single", action="store_true", help="shard mode: skip the one-GPU run of the same workload on rank 0")
    ap.add_argument("--ctl-backend", choices=["gloo", "nccl"], default="gloo", help="N > 1: torch.distributed backend of the control plane (barriers, timing reductions); the frame data never goes through it")
    ap.add_argument("--shard-deadline", type=float, default=300.0, help="shard mode: seconds after which a run that has not completed falls back to the replicas line")
    args = ap.parse_args()
    global _OUT
    sys.stdout.flush()
    _OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.mode is None:
        args.mode = "shard" if world > 1 else "replicas"
    batch_given = any(a == "--batch" or a.startswith("--batch=") for a in sys.argv[1:])
    if args.workload is None:
        args.workload = "cfg2" if (world == 1 or args.mode == "replicas") else ("cfg4" if world >= 8 else "cfg3")
    if args.mode == "shard" and not batch_given:
        args.batch = 2 if args.workload == "cfg4" else 8  # two submissions in flight: 2 x batch frame slots (16 at most)
    cfg = WORKLOADS[args.workload]

    if args.impl == "reference":
        run_reference(args, cfg, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    import vsb200
    B, S = vsb200.binding, vsb200.synth
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the compose path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        import datetime
        import resource
        try:  # NCCL's peer-to-peer transport passes file descriptors between ranks: lift the soft limit to the hard one
            soft, hard = resource.getrlimit(resource.RLIMIT_NOFILE)
            resource.setrlimit(resource.RLIMIT_NOFILE, (hard, hard))
        except Exception:
            pass
        # Control plane (barriers, max-over-ranks of the timings, the unique id) over gloo on the loopback interface; the data
        # path's NCCL communicator is the one libvsb200 owns (vsb_shard_init), so each process holds ONE set of NCCL peer
        # connections.  --ctl-backend nccl gives torch its own communicator instead.  The timeout: a rank that dies must not
        # leave its peers waiting in a collective for the rest of the driver's time limit.
        global CTL_DEV
        if args.ctl_backend == "gloo":
            os.environ.setdefault("GLOO_SOCKET_IFNAME", "lo")
            dist.init_process_group("gloo", timeout=datetime.timedelta(seconds=180))
            CTL_DEV = "cpu"
        else:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), timeout=datetime.timedelta(seconds=180))
            CTL_DEV = "cuda"

    if args.mode == "shard":
        run_sharded(args, cfg, rank, world, local_rank)
        return
    F, K, W = args.batch, args.steps, max(args.warmup, 3)
    n = cfg["n_views"]
    st = B.Stitcher(n, cfg["num_bands"], cfg["enable_local"], F)
    st.calibrate_rig(cfg["projection"], cfg["pano_width"], cfg["src_w"], cfg["src_h"], 90.0, S.gains(n))
    info = st.rig_info()
    if cfg["enable_local"]:
        for i in range(n):
            mx, my = S.mesh(info.view_roi[i][2], info.view_roi[i][3])
            st.set_mesh(i, mx.ctypes.data, my.ctypes.data, mx.shape[0], mx.shape[1])
    roi, _, nb = st.get_roi()
    OW, OH = roi[2], roi[3]
    src_pitch = cfg["src_w"] * 3
    out_pitch = (OW * 6 + 255) // 256 * 256  # pitched rows like cv::cuda::GpuMat (cudaMallocPitch): 16-byte vector stores

    # ring of RING frame sets (each rank gets different frames: frame-level data parallelism)
    n_sets = max(args.ring, F)
    host_sets = []
    for f in range(n_sets):
        host_sets.append([torch.from_numpy(S.frame(i, f + vsb200.dist.ring_seed_offset(rank, n_sets), cfg["src_w"], cfg["src_h"])).pin_memory() for i in range(n)])
    dev_sets = [[t.cuda(non_blocking=True) for t in fs] for fs in host_sets]
    outs = [torch.empty((OH, out_pitch // 2), dtype=torch.int16, device="cuda") for _ in range(F)]
    stream = torch.cuda.current_stream().cuda_stream
    torch.cuda.synchronize()

    calls = []
    for s0 in range(0, n_sets):
        srcs = [dev_sets[(s0 + j) % n_sets][i].data_ptr() for j in range(F) for i in range(n)]
        calls.append(st.make_compose_call(srcs, src_pitch, [o.data_ptr() for o in outs], out_pitch, stream))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # config 5: a second host thread keeps publishing alternating meshes (vsb_set_mesh is thread-safe and double buffered:
    # the kernels that build the maps and the tap table run on the handle's mesh stream while the compose stream keeps going)
    recal = {"stop": False, "installs": 0}
    recal_thread = None
    if cfg.get("recalib_ms"):
        meshes = [[S.mesh(info.view_roi[i][2], info.view_roi[i][3], phase=ph) for i in range(n)] for ph in (0.0, 0.7)]
        def recalibrate():
            k = 0
            while not recal["stop"]:
                for i in range(n):
                    mx, my = meshes[k & 1][i]
                    st.set_mesh(i, mx.ctypes.data, my.ctypes.data, mx.shape[0], mx.shape[1])
                recal["installs"] += 1
                k += 1
                time.sleep(cfg["recalib_ms"] / 1000.0)
        recal_thread = threading.Thread(target=recalibrate, daemon=True)
        recal_thread.start()

    for w in range(W):
        calls[w % n_sets]()
    launches_per_step = st.last_launch_count()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    t_wall0 = time.time()
    # exactly K steps first (the contract's timed region), then the protocol of SURVEY.md 8d on top of it: 5 runs of >= K steps
    # covering >= 2 s in total, the MEDIAN run is reported (`value`, `ms_per_step`); every run is barrier + synchronize bracketed
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for k in range(K):
        calls[k % n_sets]()
    e1.record()
    barrier()
    ms_k = vsb200.dist.reduce_step_time(e0.elapsed_time(e1), dist if world > 1 else None, CTL_DEV) / K  # MAX over ranks
    ms_med, runs, per_run = timed_repeats(lambda k: calls[k % n_sets](), K, barrier, repeats=5, min_total_s=2.0, probe=ms_k)
    runs = [vsb200.dist.reduce_step_time(r_, dist if world > 1 else None, CTL_DEV) for r_ in runs]
    ms_step = sorted(runs)[len(runs) // 2]
    t_wall1 = time.time()
    sampler.stop()
    if recal_thread is not None:
        recal["stop"] = True
        recal_thread.join()
        torch.cuda.synchronize()
    clocks = sampler.summary(t_wall0, t_wall1)
    ms = ms_step * K
    fps = world * F / (ms_step / 1000.0)

    # ---- F = 1: the reference's own cadence (one frame per stitch_one call, A/timed.cpp:574-615) through the same entry point:
    #      back-to-back single-frame submissions (throughput) and one frame at a time on an idle GPU (latency)
    f1 = None
    if not cfg.get("recalib_ms"):
        calls1 = [st.make_compose_call([dev_sets[s0][i].data_ptr() for i in range(n)], src_pitch, [outs[0].data_ptr()], out_pitch, stream) for s0 in range(n_sets)]
        for w in range(3):
            calls1[w % n_sets]()
        m1, _, _ = timed_repeats(lambda k: calls1[k % n_sets](), max(20, K), barrier, repeats=5, min_total_s=0.5)
        lat = []
        for k in range(21):
            torch.cuda.synchronize()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record(); calls1[k % n_sets](); a1.record()
            torch.cuda.synchronize()
            lat.append(a0.elapsed_time(a1))
        f1 = {"value_f1": world * 1000.0 / vsb200.dist.reduce_step_time(m1, dist if world > 1 else None, CTL_DEV), "unit": "frames/s",
              "latency_ms_f1": sorted(lat)[len(lat) // 2], "launches_per_frame": st.last_launch_count(),
              "note": "vsb_compose with n_frames = 1: the remap tap tables are read once per frame instead of once per submission"}

    # ---- per-kernel device times (events on the launching stream around every kernel), dominant kernel roofline
    st.set_profiling(True)
    acc = {}
    prof_steps = min(K, 50)
    for k in range(prof_steps):
        calls[k % n_sets]()
        for name, t_ms, nbytes in st.get_profile():
            a = acc.setdefault(name, [0.0, nbytes])
            a[0] += t_ms
    st.set_profiling(False)
    kernels = {name: {"ms": a[0] / prof_steps, "alg_bytes": a[1], "GBps": a[1] / (a[0] / prof_steps) / 1e6} for name, a in acc.items()}
    top = max(kernels, key=lambda k_: kernels[k_]["ms"])
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "B200_PROFILING.md fallback 6650 GB/s"
    b_io = n * cfg["src_w"] * cfg["src_h"] * 3 + OW * OH * 6  # sources once + CV_16SC3 pano once
    roofline = {"bound": "hbm", "kernel": top, "achieved": kernels[top]["GBps"], "peak": peak, "unit": "GB/s",
                "frac": kernels[top]["GBps"] / peak, "traffic": ncu_traffic(top, F), "peak_source": peak_src,
                "alg_bytes_per_launch": kernels[top]["alg_bytes"], "ms_per_launch": kernels[top]["ms"],
                "share_of_step": kernels[top]["ms"] / sum(v["ms"] for v in kernels.values())}
    roofline_path = {"alg_bytes_per_frame": b_io, "achieved": b_io * fps / world / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": b_io * fps / world / 1e9 / peak, "note": "whole path, B_io = sources once + pano once (SURVEY.md 8d)"}

    # ---- e2e through the host-buffer entry point (pinned host memory in, host memory out)
    e2e = None
    if not args.no_e2e:
        # two sets of pinned host panoramas: submission k + 1 is enqueued while submission k drains (vsb_submit_host / vsb_wait_host)
        h_outs = [[torch.empty((OH, OW, 3), dtype=torch.int16).pin_memory() for _ in range(F)] for _ in range(2)]
        hcalls = {(par, s0): st.make_submit_host_call([host_sets[(s0 + j) % n_sets][i].data_ptr() for j in range(F) for i in range(n)], src_pitch,
                                                      [o.data_ptr() for o in h_outs[par]], OW * 6) for par in range(2) for s0 in range(n_sets)}
        def host_run(n_steps):
            for k in range(n_steps):
                hcalls[(k & 1, k % n_sets)]()
                if k >= 1:
                    st.wait_host()
            st.wait_host()
        host_run(3)
        Ke = max(4, min(K, 40))
        barrier()
        t0 = time.perf_counter()
        host_run(Ke)
        dt = vsb200.dist.reduce_step_time(time.perf_counter() - t0, dist if world > 1 else None, CTL_DEV)
        e2e = {"value": world * F * Ke / dt, "unit": "frames/s", "h2d_bytes_per_step": F * n * cfg["src_w"] * cfg["src_h"] * 3,
               "d2h_bytes_per_step": F * OW * OH * 6, "steps": Ke,
               "api": "vsb_submit_host / vsb_wait_host: pinned host frames in, host panoramas out; upload / compose / download pipelined over "
                      "2-frame sub-batches and over two submissions in flight; wall clock over all steps incl. the last wait"}

    # ---- the same through the wire / consumer formats: NV12 frames in as the capture boards send them
    #      (the reference converts them on the CPU before its upload, A/networking.cpp:46), CV_8UC3 panoramas out (the reference
    #      converts on the GPU before its download, A/timed.cpp:250-251): half the bytes in each direction over PCIe
    e2e_wire = None
    if not args.no_e2e and cfg["src_w"] % 2 == 0 and cfg["src_h"] % 2 == 0 and nb >= 3:
        sw_, sh_ = cfg["src_w"], cfg["src_h"]
        nv_sets = []
        for fs in host_sets:
            one = []
            for t in fs:
                nvf = torch.empty((sh_ * 3 // 2, sw_), dtype=torch.uint8)
                nvf[:sh_] = t[..., 1]
                nvf[sh_:, 0::2] = t[::2, ::2, 0]
                nvf[sh_:, 1::2] = t[::2, ::2, 2]
                one.append(nvf.pin_memory())
            nv_sets.append(one)
        h_outs8 = [[torch.empty((OH, OW, 3), dtype=torch.uint8).pin_memory() for _ in range(F)] for _ in range(2)]
        st.set_formats(B.IN_NV12, B.OUT_U8C3)
        wcalls = {(par, s0): st.make_submit_host_call([nv_sets[(s0 + j) % n_sets][i].data_ptr() for j in range(F) for i in range(n)], sw_,
                                                      [o.data_ptr() for o in h_outs8[par]], OW * 3) for par in range(2) for s0 in range(n_sets)}
        def wire_run(n_steps):
            for k in range(n_steps):
                wcalls[(k & 1, k % n_sets)]()
                if k >= 1:
                    st.wait_host()
            st.wait_host()
        wire_run(3)
        Ke = max(4, min(K, 40))
        barrier()
        t0 = time.perf_counter()
        wire_run(Ke)
        dt = vsb200.dist.reduce_step_time(time.perf_counter() - t0, dist if world > 1 else None, CTL_DEV)
        st.set_formats(B.IN_BGR8, B.OUT_S16C3)
        e2e_wire = {"value": world * F * Ke / dt, "unit": "frames/s", "h2d_bytes_per_step": F * n * sw_ * sh_ * 3 // 2,
                    "d2h_bytes_per_step": F * OW * OH * 3, "steps": Ke,
                    "api": "vsb_set_formats(VSB_IN_NV12, VSB_OUT_U8C3) + vsb_submit_host / vsb_wait_host: NV12 host frames in, CV_8UC3 host panoramas out"}

    # ---- CPU baseline on the host cores, rank 0 at N=1 only, bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = host_cores()
        frames_np = [[t.numpy() for t in fs] for fs in host_sets[:4]]
        cfps, kind, used, sample = cpu_compose_run(cfg, frames_np, args.cpu_seconds, threads)
        cpu = {"value": cfps, "unit": "frames/s", "cores": used, "kind": kind, "sample": sample}

    if rank == 0:
        line = {
            "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8/s16 (fp32 taps)", "data": "synthetic",
            "config": {"workload": cfg["name"], "frames_per_step": F, "ring_frame_sets": n_sets,
                       "l2_policy": f"inputs larger than L2: ring of {n_sets} frame sets = {n_sets * n * cfg['src_w'] * cfg['src_h'] * 3 / 1e6:.0f} MB",
                       "pano": f"{OW}x{OH} CV_16SC3", "bands": nb, **({"mesh_installs_during_run": recal["installs"]} if cfg.get("recalib_ms") else {}), "multi_gpu": "frame-level replicas, no collective" if world > 1 else "single GPU"},
            "clocks": clocks, "e2e": e2e, "e2e_wire": e2e_wire, "gpu_launches": launches_per_step * K, "launches_per_step": launches_per_step,
            "timing": {"protocol": f"median of 5 runs of {per_run} steps (>= 2 s in total), CUDA events on the launching stream, barrier + synchronize around every run, max over ranks",
                       "runs_ms_per_step": runs, "first_K_steps_ms_per_step": ms_k},
            "f1": f1,
            "roofline": roofline, "roofline_path": roofline_path, "kernels": kernels, "cpu_baseline": cpu,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
