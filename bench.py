#!/usr/bin/env python
"""bench.py — stabilized frames/s of the LiveVisionKit stabilization path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA, one stream per GPU)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port) on host cores
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...   # N > 1: one rank per GPU

A "step" is one frame through the whole hot path (ingest -> FAST grid -> pyramidal LK -> homography RANSAC ->
path smoother -> EASU remap of the 10-frames-older frame).  Workload at N=1: BASELINE.json configs[1]
(1920x1080, 60 fps synthetic hand-shake clip, OBS "Homography" preset).  Per-GPU work is fixed as N grows
(one independent stream per GPU, seeds 42+rank) -> weak scaling; NCCL only gathers the counters.

Timing.  A WINDOW is exactly K consecutive steps bracketed by barrier + synchronize before and a stream
synchronize after; no barrier and no collective inside a window.  The window is repeated R times back to back
on fresh frames (R = --windows, default 300 // K clamped to 1..15, so a 20-step request is measured 15 times);
each rank keeps the MEDIAN of its R windows, the line reports the MAX of those medians over ranks (the slowest
rank), and `per_rank` carries every rank's min / median / max so that rank skew and host noise can be told apart.
value : whole-job fps with every input frame already resident in HBM and outputs left in HBM (CUDA events on the
        library's stream around each window; the window's K frames go through lvkb200_stream_submit_batch).
e2e   : the same metric through the public pipelined API with pinned HOST buffers (H2D of the frame and D2H of the
        result inside the window, every step), host clock between the rank's own stream synchronisations.
roofline / roofline_issue: the dominant kernel (EASU remap): algorithmic bytes (6 B/px) resp. executed
        warp-instructions per launch / its average duration measured with CUDA events on the library's own CUDA
        stream inside the timed windows, against MEASURED_PEAKS.json resp. SMs x 4 schedulers x SM clock.
cpu_baseline: the oracle port (cv2 + scalar C EASU, all host cores) on a bounded sample of the same workload.
configs: the other single-GPU BASELINE configurations measured in the same run at N=1 (4K60 with fewer steps,
        720p30 pan with library defaults on the CPU path).
Host hygiene at N > 1: every rank pins itself (and its clock sampler) to a disjoint core set by LOCAL_RANK, NVML is
polled every 10 ms, pinned staging memory is ONE cudaHostAlloc per rank.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

SIZES = {"720p": (1280, 720), "1080p": (1920, 1080), "4k": (3840, 2160)}
_TRACK = {
    "H": ("OBS Homography", "OBS Homography preset (480x270 detection, FAST grid -> pyramidal LK -> homography RANSAC"),
    "F": ("OBS Vector Field", "OBS Vector Field preset (480x270 detection, FAST grid -> pyramidal LK -> 16x16 local-motion "
                              "LSCG mesh"),
    "D": ("library defaults", "library-default settings (256x256 detection, FAST grid -> pyramidal LK -> local-motion LSCG "
                              "mesh"),
}


class Workload:
    """BASELINE.json configs[1] (1080p60, default: the configuration the metric is quoted on), configs[2] (4K60,
    `--resolution 4k`), configs[0] (720p30 pan, defaults; CPU line) and configs[4]'s per-stream chain (`--deblock`).
    Tracking presets: H (OBS "Homography", default), D (library defaults), F (OBS "Vector Field")."""

    def __init__(self, res="1080p", preset="H", deblock=False, kind="shake", fps=60):
        self.res, self.preset, self.deblock, self.kind, self.fps = res, preset, deblock, kind, fps
        self.width, self.height = SIZES[res]
        self.preset_name = _TRACK[preset][0]
        self.metric = f"stabilized_frames_per_second_{res}"
        label = {"720p": "720p", "1080p": "1080p", "4k": "4K"}[res]
        motion = "pan clip" if kind == "pan" else "hand-shake sequence"
        self.workload = (f"{label}{fps} synthetic {motion}, {_TRACK[preset][1]} -> path smoother -> FSR-EASU remap), "
                         f"1 stream per GPU")
        if deblock:
            self.workload = "DeblockingFilter (defaults) -> " + self.workload

    def gpu_settings(self, L):
        S = L.StabilizationFilterSettings
        return {"H": S.obs_homography_preset, "F": S.obs_field_preset, "D": S}[self.preset]()

    def oracle_settings(self, O):
        S = O.StabilizationSettings
        return {"H": S.obs_homography_preset, "F": S.obs_field_preset, "D": S}[self.preset]()

    def make_filter(self, L, device):
        """The timed filter: lvk::StabilizationFilter, or with --deblock CompositeFilter{DeblockingFilter,
        StabilizationFilter} (BASELINE configs[4]) fused on the device."""
        flt = L.StabilizationFilter(self.gpu_settings(L), device=device)
        if self.deblock:
            flt.stream.set_deblocking(L.DeblockingFilterSettings())
        return flt

    def frame_bytes(self):
        return self.width * self.height * 3


class _OracleChain:
    """The same chain on the CPU for the baseline legs."""

    def __init__(self, O, wl, remap_threads):
        self.flt = O.StabilizationFilter(wl.oracle_settings(O), remap_threads=remap_threads)
        self.deblock = None
        if wl.deblock:
            from oracle import deblock_oracle as D
            self.deblock = D.DeblockingFilter()

    def apply(self, frame, fmt, ts):
        if self.deblock is not None:
            frame = self.deblock.apply(frame, fmt)
        return self.flt.apply(frame, fmt, ts)


def _peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


def _dist_env():
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    return rank, local, world


def rank_cores(local, world, cores=None):
    """Disjoint core set of one rank: the process's allowed cores split evenly by LOCAL_RANK (>= 1 core each)."""
    cores = sorted(os.sched_getaffinity(0)) if cores is None else sorted(cores)
    if world <= 1 or len(cores) < world:
        return cores
    per = len(cores) // world
    return cores[local * per:(local + 1) * per]


def _pin_rank(local, world):
    try:
        mine = rank_cores(local, world)
        os.sched_setaffinity(0, mine)
        return mine
    except Exception:
        return None


def default_windows(steps):
    return max(1, min(15, 300 // max(steps, 1)))


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU while the timed windows run (NVML every 10 ms; nvidia-smi as the
    fallback).  The thread inherits the rank's core set."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")
    PERIOD_S = 0.010

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None
        self.halt = threading.Event()
        self.nvml = None
        try:  # NVML is initialised HERE (before the timed region starts)
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(index)
            self.nvml = (nv, h, nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None

    def _run_nvml(self):
        if self.nvml is None:
            raise RuntimeError("NVML unavailable")
        nv, h, smax = self.nvml
        bits = (("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap))
        while not self.halt.is_set():
            sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
            mask = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
            self.rows.append([str(sm), str(smax), "0"] + ["Active" if mask & b else "Not Active" for _, b in bits])
            self.halt.wait(self.PERIOD_S)

    def run(self):
        try:
            self._run_nvml()
            return
        except Exception:
            self.rows = []
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        self.halt.set()
        if self.proc:
            self.proc.terminate()
        self.join(timeout=1.0)
        sm, reasons, smax = [], set(), None
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                smax = float(r[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm), "period_ms": 1e3 * self.PERIOD_S}


# ---------------------------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU algorithm (oracle port) on the host cores


def _reference_stream(wl_args, seed, warmup, steps, threads, start, out):
    """One CPU stream (own process): builds its clip, warms up, waits at `start`, times `steps` frames."""
    import cv2
    from oracle import lvk_oracle as O
    from tools.synth import Clip
    wl = Workload(*wl_args)
    cv2.setNumThreads(threads)
    clip = Clip(wl.res, wl.kind, frames=warmup + steps, seed=seed, fps=wl.fps)
    frames = [clip[i] for i in range(len(clip))]
    flt = _OracleChain(O, wl, threads)
    for i in range(warmup):
        flt.apply(frames[i], O.BGR, i)
    start.wait()
    t0 = time.perf_counter()
    for i in range(warmup, warmup + steps):
        flt.apply(frames[i], O.BGR, i)
    out.put((seed, t0, time.perf_counter()))


def run_reference(args, wl):
    """--impl reference: rank 0 alone runs and prints.  At --gpus N it runs N independent CPU streams concurrently (the
    like-for-like of N GPU streams), the host cores split between them; value = N * K frames / the time until the last
    stream finishes."""
    rank, _, world = _dist_env()
    if rank != 0:
        return  # the other ranks exit 0 without work
    import cv2
    from oracle import lvk_oracle as O
    O.build_native()
    cores = len(os.sched_getaffinity(0)) or 1
    streams = max(1, args.gpus)
    wl_args = (wl.res, wl.preset, wl.deblock, wl.kind, wl.fps)
    if streams == 1:
        from tools.synth import Clip
        cv2.setNumThreads(cores)
        clip = Clip(wl.res, wl.kind, frames=args.warmup + args.steps, seed=42, fps=wl.fps)
        frames = [clip[i] for i in range(len(clip))]
        flt = _OracleChain(O, wl, cores)
        for i in range(args.warmup):
            flt.apply(frames[i], O.BGR, i)
        t0 = time.perf_counter()
        for i in range(args.warmup, args.warmup + args.steps):
            flt.apply(frames[i], O.BGR, i)
        dt = time.perf_counter() - t0
        threads = cores
    else:
        import multiprocessing as mp
        ctx = mp.get_context("spawn")
        threads = max(1, cores // streams)
        start, out = ctx.Barrier(streams), ctx.Queue()
        procs = [ctx.Process(target=_reference_stream, args=(wl_args, 42 + r, args.warmup, args.steps, threads, start, out))
                 for r in range(streams)]
        for p in procs:
            p.start()
        spans = [out.get() for _ in procs]
        for p in procs:
            p.join()
        dt = max(s[2] for s in spans) - min(s[1] for s in spans)
    fps = streams * args.steps / dt
    line = {
        "impl": "reference", "metric": wl.metric, "value": fps, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl.workload, "resolution": [wl.width, wl.height], "preset": wl.preset_name},
        "reference_streams": streams,
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} frames after {args.warmup} warm-up frames on each of {streams} concurrent "
                                   f"stream(s), cv2 {cv2.__version__} (reference pins 4.8.0) + scalar C EASU, "
                                   f"{threads} threads per stream on {cores} cores"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_sample(wl, frames, warm=12, count=60):
    """Bounded CPU sample of the same workload (oracle port), rank 0 at N=1 only."""
    import cv2
    from oracle import lvk_oracle as O
    O.build_native()
    cores = len(os.sched_getaffinity(0)) or 1
    cv2.setNumThreads(cores)
    flt = _OracleChain(O, wl, cores)
    count = min(count, len(frames) - warm)
    for i in range(warm):
        flt.apply(frames[i], O.BGR, i)
    t0 = time.perf_counter()
    for i in range(warm, warm + count):
        flt.apply(frames[i], O.BGR, i)
    dt = time.perf_counter() - t0
    return {"value": count / dt, "unit": "frames/s", "cores": cores, "kind": "port",
            "sample": f"{count} frames of the same {wl.res} clip after {warm} warm-up frames, oracle port "
                      f"(cv2 {cv2.__version__} + scalar C EASU, {cores} threads)"}


# ---------------------------------------------------------------------------------------------------------------------
# the GPU arm


def _stats(xs):
    xs = sorted(float(x) for x in xs)
    return {"min": xs[0], "median": float(np.median(xs)), "max": xs[-1]}


def measure_gpu(wl, steps, warmup, windows, local, world, lookahead=True, apply_pass=True, stage_pass=True, sampler=None,
                nv12_pass=False):
    """Runs `windows` timed windows of `steps` steps each for the device-resident pass and for the pipelined host pass
    on this rank's GPU.  Returns the per-window times (ms) and the side measurements; no cross-rank reduction here."""
    import torch
    import torch.distributed as dist
    import livevisionkit_b200 as L
    from tools.synth import Clip
    from tools.scaling import stream_seed

    rank = int(os.environ.get("RANK", "0"))
    dev = torch.device("cuda", local)
    K, R = steps, windows
    n_frames = warmup + R * K

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- synthetic clip, rendered straight into ONE pinned allocation (cudaHostAlloc once per rank)
    clip = Clip(wl.res, wl.kind, frames=n_frames, seed=stream_seed(rank), fps=wl.fps)
    pinned = torch.empty((n_frames + 4, wl.height, wl.width, 3), dtype=torch.uint8).pin_memory()
    pinned_in, pinned_out = [pinned[i] for i in range(n_frames)], [pinned[n_frames + i] for i in range(4)]
    for i in range(n_frames):
        pinned_in[i].copy_(torch.from_numpy(clip[i]))
    settings_lookahead = lookahead

    # ======== pass 1: device-resident (value + roofline) ========
    dev_all = torch.empty((n_frames, wl.height, wl.width, 3), dtype=torch.uint8, device=dev)
    dev_all.copy_(pinned[:n_frames], non_blocking=True)
    dev_frames = [dev_all[i] for i in range(n_frames)]
    out_ring = [torch.empty_like(dev_frames[0]) for _ in range(16)]
    torch.cuda.synchronize()
    flt = wl.make_filter(L, local)
    s = flt.stream
    # Each frame is announced one step ahead (lvkb200_stream_prefetch_frame, the input thread of VideoFilter::stream
    # running ahead of the filter thread): its copy into the stream's ring and its detection image + pyramid are queued
    # behind the current frame's tracking chain.  All of a frame's work still happens inside the timed windows.
    dev_refs = [L.FrameRef(f) for f in dev_frames]
    out_refs = [L.FrameRef(o) for o in out_ring]

    def dev_step(i):
        if settings_lookahead and i + 1 < n_frames:
            s.prefetch(dev_refs[i + 1], L.BGR)
        return s.submit(dev_refs[i], out_refs[i % 16], L.BGR, i)

    for i in range(warmup):
        dev_step(i)
    s.sync()
    s.stage_totals_us(reset=True)
    launches0 = L._capi.load().lvkb200_kernel_launch_count()
    if sampler is not None:
        sampler.start()
    dev_ms, wall_ms, outputs = [], [], 0
    # the K frames of a window go through lvkb200_stream_submit_batch (VideoFilter::stream for frames already in memory:
    # frame i+1 announced, frame i submitted, for the whole window in ONE call); the pointer tables are built beforehand
    plans = [L.BatchPlan([dev_refs[i] for i in range(warmup + w * K, warmup + (w + 1) * K)],
                         [out_refs[i % 16] for i in range(warmup + w * K, warmup + (w + 1) * K)],
                         list(range(warmup + w * K, warmup + (w + 1) * K))) for w in range(R)] if settings_lookahead else None
    for w in range(R):
        barrier()
        s.event_record(0)
        t0 = time.perf_counter()
        if plans is not None:
            outputs += sum(r.has_output for r in s.submit_batch(plans[w], None, L.BGR))
        else:
            for i in range(warmup + w * K, warmup + (w + 1) * K):
                outputs += dev_step(i).has_output
        s.event_record(1)
        s.sync()
        wall_ms.append(1e3 * (time.perf_counter() - t0))
        dev_ms.append(s.event_elapsed_ms(0, 1))
    launches = L._capi.load().lvkb200_kernel_launch_count() - launches0
    totals, counts = s.stage_totals_us(reset=True)  # default mode: only the remap kernel is event-timed
    last_dev_out = out_ring[(n_frames - 1) % 16].cpu().numpy().copy()
    stage_us = None
    if stage_pass:
        # untimed extra pass with per-stage CUDA events (eager launches instead of the tracking graph)
        prof = wl.make_filter(L, local)
        prof.stream.set_profiling(True)
        n_prof = min(n_frames, 90)
        for i in range(n_prof):
            prof.stream.submit(dev_frames[i], out_ring[i % 16], L.BGR, i)
            if i == min(29, n_prof // 2):
                prof.stream.stage_totals_us(reset=True)
        ptotals, pcounts = prof.stream.stage_totals_us(reset=True)
        prof.stream.close()
        stage_us = {k: (ptotals[k] / pcounts[k] if pcounts[k] else 0.0) for k in ptotals}
    s.close()
    del dev_frames, dev_refs, dev_all
    torch.cuda.empty_cache()

    # ======== pass 2: end to end through the synchronous per-frame public API (one window, informational) ========
    parity_fail, apply_fps = 0, None
    if apply_pass:
        flt2 = wl.make_filter(L, local)
        for i in range(warmup):
            flt2.apply(L.VideoFrame(pinned_in[i], i, L.BGR), output=pinned_out[i % 4])
        torch.cuda.synchronize()
        ta = time.perf_counter()
        for i in range(warmup, n_frames):
            flt2.apply(L.VideoFrame(pinned_in[i], i, L.BGR), output=pinned_out[i % 4])
        flt2.stream.sync()
        apply_fps = (n_frames - warmup) / (time.perf_counter() - ta)
        parity_fail += int(not np.array_equal(pinned_out[(n_frames - 1) % 4].numpy(), last_dev_out))
        flt2.stream.close()

    # ======== pass 3: end to end through the pipelined public API (VideoFilter::stream analogue) ========
    # upload of frame t+1 and download of output t-1 overlap the processing of frame t; still one H2D of the input and
    # one D2H of the result per step, all inside the window.  The window's clock runs on this rank alone: from its own
    # synchronised start to its own stream synchronisation (no barrier, no collective inside).
    flt3 = wl.make_filter(L, local)
    pin_refs = [L.FrameRef(t) for t in pinned_in]
    pout_refs = [L.FrameRef(t) for t in pinned_out]
    flt3.stream([L.VideoFrame(pin_refs[i], i, L.BGR) for i in range(warmup)], lambda vf: False, pout_refs[:3])
    e2e_ms, delivered = [], 0
    # lvkb200_stream_submit_batch with HOST frames and HOST outputs = the pipelined path in one FFI call per window (frame
    # i+1 uploading, output i-1 downloading while frame i is filtered; every output has landed when it returns); the
    # outputs cycle through four pinned buffers
    e2e_plans = [L.BatchPlan([pin_refs[i] for i in range(warmup + w * K, warmup + (w + 1) * K)],
                             [pout_refs[j % 4] for j in range(K)], list(range(warmup + w * K, warmup + (w + 1) * K)))
                 for w in range(R)]
    for w in range(R):
        barrier()
        tp = time.perf_counter()
        delivered += sum(r.has_output for r in flt3.stream.submit_batch(e2e_plans[w], None, L.BGR))
        flt3.stream.sync()
        e2e_ms.append(1e3 * (time.perf_counter() - tp))
    parity_fail += int(not np.array_equal(pinned_out[(K - 1) % 4].numpy(), last_dev_out))
    parity_fail += int(delivered != R * K)
    flt3.stream.close()

    # ======== pass 4: the same, frames crossing PCIe as NV12 (the OBS plugin's own layout, FrameIngest) ========
    # lvkb200_stream_prefetch_obs / _submit_obs_async: 1.5 B/px in each direction instead of 3; the plane <-> packed
    # conversions (to_ocl / to_obs) run on the copy streams.  Same clip, converted to NV12 on the host beforehand.
    nv12_ms = None
    if nv12_pass:
        import cv2
        h, w = wl.height, wl.width
        nv = torch.empty((n_frames + 4, h * 3 // 2, w), dtype=torch.uint8).pin_memory()
        for i in range(n_frames):
            i420 = cv2.cvtColor(pinned_in[i].numpy(), cv2.COLOR_BGR2YUV_I420)
            dst = nv[i].numpy()
            dst[:h] = i420[:h]
            uv = dst[h:].reshape(h // 2, w // 2, 2)
            uv[:, :, 0] = i420[h:h + h // 4].reshape(h // 2, w // 2)
            uv[:, :, 1] = i420[h + h // 4:].reshape(h // 2, w // 2)

        def obs(t, ts):
            a = t.numpy()
            return L.ObsFrame("NV12", w, h, [a[:h], a[h:]], timestamp=ts)

        srcs = [obs(nv[i], i) for i in range(n_frames)]
        outs = [obs(nv[n_frames + i], 0) for i in range(4)]
        flt4 = wl.make_filter(L, local)
        flt4.stream.stream_obs(srcs[:warmup], lambda o: False, outs[:3])
        nv12_ms, got = [], 0
        for wi in range(R):
            window = srcs[warmup + wi * K: warmup + (wi + 1) * K]
            wouts = [outs[j % 4] for j in range(K)]
            barrier()
            tp = time.perf_counter()
            got += sum(r.has_output for r in flt4.stream.submit_obs_batch(window, wouts))
            flt4.stream.sync()
            nv12_ms.append(1e3 * (time.perf_counter() - tp))
        parity_fail += int(got != R * K)
        flt4.stream.close()
        del nv, srcs, outs
    del pinned, pinned_in, pinned_out
    return {"dev_ms": dev_ms, "e2e_ms": e2e_ms, "nv12_ms": nv12_ms, "wall_ms": wall_ms, "launches": launches, "outputs": outputs,
            "parity_fail": parity_fail, "apply_fps": apply_fps, "stage_us": stage_us, "n_frames": n_frames,
            "remap_us": totals["remap"] / max(counts["remap"], 1), "clip": clip}


def _kernel_facts(res):
    """Per-launch constants of the remap kernel from the committed ncu captures (profiles/remap_traffic.json)."""
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "remap_traffic.json")))
        return prof.get(f"dram_bytes_per_launch_{res}"), prof.get("thread_instructions_per_pixel"), prof.get("source")
    except Exception:
        return None, None, None


def rooflines(wl, remap_us, sm_mhz):
    peaks, peak_kind = _peaks()
    peak = float(peaks.get("hbm_gbs", 6650.0))
    alg_bytes = 6.0 * wl.width * wl.height
    achieved = alg_bytes / (remap_us * 1e-6) / 1e9 if remap_us > 0 else 0.0
    traffic, instr_px, src = _kernel_facts(wl.res)
    # the 16x16 "Vector Field" mesh is warped by the kernel's mesh variant (per-pixel bilinear lookup of the offset field
    # in front of the same EASU body); the committed instruction count is the homography variant's, a lower bound for it
    kernel = "k_easu_remap_fast<mesh>" if wl.preset == "F" else "k_easu_remap_fast<homography>"
    hbm = {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "GB/s",
           "frac": achieved / peak, "traffic": traffic, "peak_source": peak_kind, "avg_kernel_us": remap_us,
           "algorithmic_bytes_per_launch": alg_bytes,
           "note": "EASU is bound by instruction issue / FP32 pipe / shared-memory bandwidth, not by HBM (DRAM ~2 % busy): "
                   "see roofline_issue and DESIGN.md 5.1"}
    issue = None
    if instr_px and remap_us > 0:
        clk = (sm_mhz or float(peaks.get("sm_max_mhz", 1965.0))) * 1e6
        warp_instr = instr_px * wl.width * wl.height / 32.0
        peak_ips = 148 * 4 * clk  # warp-instructions per second: SMs x schedulers x clock
        ach = warp_instr / (remap_us * 1e-6)
        issue = {"bound": "issue", "kernel": kernel, "achieved": ach / 1e9, "peak": peak_ips / 1e9,
                 "unit": "Gwarp-instr/s", "frac": ach / peak_ips, "thread_instructions_per_pixel": instr_px,
                 "warp_instructions_per_launch": warp_instr, "sm_mhz": clk / 1e6, "source": src}
    return hbm, issue


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=30)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--windows", type=int, default=0,
                    help="how many times the K-step window is measured (median reported); 0 = 300 // K clamped to 1..15")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra-configs", action="store_true",
                    help="skip the `configs` block (4K60 and the 720p30-pan CPU line, N=1 only)")
    ap.add_argument("--no-lookahead", action="store_true",
                    help="device-resident pass without announcing frame t+1 before submitting frame t")
    ap.add_argument("--resolution", default="1080p", choices=["1080p", "4k"])
    ap.add_argument("--preset", default="H", choices=["H", "D", "F"])
    ap.add_argument("--deblock", action="store_true",
                    help="BASELINE configs[4]: DeblockingFilter -> StabilizationFilter chained (per stream)")
    args = ap.parse_args()
    # the stabilizer delivers frame t-10: the look-ahead queue must be full before the timed region starts, so that
    # every timed step does a step's whole work (tracking AND a remap) and produces an output; the JSON line reports
    # the warm-up actually run
    args.warmup = max(args.warmup, 12)
    wl = Workload(args.resolution, args.preset, args.deblock)

    if args.impl == "reference":
        run_reference(args, wl)
        return

    rank, local, world = _dist_env()
    cores = _pin_rank(local, world)  # before torch / CUDA / NCCL start their threads: they inherit the mask

    import torch
    import torch.distributed as dist
    import livevisionkit_b200 as L

    # stdout carries exactly ONE JSON line: anything a library prints on fd 1 meanwhile (e.g. NCCL's version banner)
    # is diverted to stderr, and the line is written to the saved descriptor at the end.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    assert L.device_count() > 0, "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    K = args.steps
    R = args.windows if args.windows > 0 else default_windows(K)
    sampler = ClockSampler(local)
    m = measure_gpu(wl, K, args.warmup, R, local, world, lookahead=not args.no_lookahead, sampler=sampler,
                    nv12_pass=not args.deblock)
    clocks = sampler.stop()

    # ======== reduce over ranks: ONE all_gather of the per-rank counter struct (NCCL over NVLink) ========
    from tools import scaling
    d, e = _stats(m["dev_ms"]), _stats(m["e2e_ms"])
    nv = _stats(m["nv12_ms"]) if m["nv12_ms"] else {"min": 0.0, "median": 0.0, "max": 0.0}
    mine = torch.tensor([float(K), d["median"], e["median"], float(m["launches"]) / R, float(m["parity_fail"]),
                         float(np.median(m["wall_ms"])), e["median"], float(m["outputs"]) / R,
                         d["min"], d["max"], e["min"], e["max"], nv["median"]], dtype=torch.float64, device=dev)
    agg = scaling.aggregate(scaling.gather_counters(mine))
    # copy-only ceiling of THIS box with all ranks copying at once (tools/pcie_ceiling.py): what bounds `e2e` at N > 1
    ceiling = None
    try:
        from tools import pcie_ceiling as PC
        rows = PC.gather(PC.probe(local, world, wl.frame_bytes(), copies=40, dist=dist if world > 1 else None), local, world, dist)
        ceiling = PC.summarise(rows, wl.frame_bytes())
    except Exception as ex:
        ceiling = {"failed": repr(ex)}
    if rank == 0:
        t_dev, t_e2e = agg["dev_ms"], agg["e2e_ms"]
        if ceiling and "ceiling_fps_whole_job" in ceiling and agg.get("per_rank"):
            ceiling["e2e_sum_of_ranks_over_ceiling"] = agg["per_rank"]["sum_of_rank_e2e_fps"] / ceiling["ceiling_fps_whole_job"]
        hbm, issue = rooflines(wl, m["remap_us"], clocks.get("sm_mhz"))
        line = {
            "metric": wl.metric, "value": agg["value_fps"], "unit": "frames/s", "n_gpus": world,
            "steps": K, "warmup": args.warmup, "ms_per_step": t_dev / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl.workload, "resolution": [wl.width, wl.height], "preset": wl.preset_name,
                       "streams_per_gpu": 1, "windows": R,
                       "l2": f"{m['n_frames']} distinct frames ({m['n_frames'] * wl.frame_bytes() / 1e9:.2f} GB per GPU) "
                             f"streamed once each: inputs larger than L2, no flush needed",
                       "timing": f"a window = exactly {K} steps between barrier+synchronize and the rank's own stream "
                                 f"synchronize; {R} windows back to back on fresh frames; per rank the median window, over "
                                 f"ranks the max (slowest rank); value: CUDA events on the library's stream, e2e: host clock",
                       "lookahead": not args.no_lookahead, "host_cores_per_rank": len(cores) if cores else None,
                       "remap_build": "exact" if L.remap_exact() else "contract"},
            "e2e": {"value": agg["e2e_fps"], "unit": "frames/s", "h2d_bytes_per_step": wl.frame_bytes(),
                    "d2h_bytes_per_step": wl.frame_bytes(), "ms_per_step": t_e2e / K,
                    "api": "Stream.submit_batch(host frames, host outputs) = lvkb200_stream_submit_batch — the pipelined "
                           "VideoFilter::stream analogue in one FFI call per window: pinned host input -> H2D -> filter -> "
                           "D2H -> pinned host output for every step, uploads and downloads overlapped with the "
                           "neighbouring frames' processing, all outputs landed when the call returns",
                    "apply_fps_rank0": m["apply_fps"],
                    "apply_note": "same, through the synchronous per-frame StabilizationFilter.apply (no overlap)"},
            "e2e_nv12": None if not agg.get("nv12_ms") else {
                "value": agg["frames"] / (agg["nv12_ms"] * 1e-3), "unit": "frames/s", "ms_per_step": agg["nv12_ms"] / K,
                "h2d_bytes_per_step": wl.frame_bytes() // 2, "d2h_bytes_per_step": wl.frame_bytes() // 2,
                "per_rank_fps": agg.get("nv12_fps_per_rank"),
                "api": "Stream.submit_obs_batch = lvkb200_stream_submit_obs_batch (prefetch_obs / submit_obs_async inside): the same clip as NV12 planes in "
                       "pinned host memory (the OBS plugin's native layout, FrameIngest.cpp:566-604); plane upload + to_ocl on "
                       "the copy-in stream, to_obs + plane download on the copy-out stream; max over ranks of the median window"},
            "per_rank": agg.get("per_rank"),
            "pcie_ceiling": ceiling,
            "gpu_launches": agg["launches"],
            "roofline": hbm,
            "roofline_issue": issue,
            "stage_us": m["stage_us"],
            "stage_us_note": "separate untimed pass with per-stage CUDA events (profiling mode, eager launches)",
            "clocks": clocks,
            "outputs": agg["outputs"],
            "parity_failures": agg["parity_failures"],
            "parity_note": "consistency of the three GPU passes with each other (device / apply / stream); parity with "
                           "the reference is what tests/ assert",
            "wall_ms_per_step": agg["wall_ms"] / K,
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                host_frames = [m["clip"][i] for i in range(min(m["n_frames"], 72))]
                line["cpu_baseline"] = cpu_baseline_sample(wl, host_frames)
            except Exception as ex:  # the baseline must never take the bench down
                line["cpu_baseline"] = {"value": None, "unit": "frames/s", "cores": os.cpu_count(), "kind": "port",
                                        "sample": f"failed: {ex}"}
        if world == 1 and not args.no_extra_configs and args.resolution == "1080p" and not args.deblock and args.preset == "H":
            line["configs"] = extra_configs(K, args.warmup, local)
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def extra_configs(K, warmup, local):
    """The other single-GPU BASELINE configurations, same timing method with fewer steps (N=1 only):
    configs[2] 4K60 hand-shake on the GPU, configs[0] 720p30 pan with library defaults on the CPU path."""
    out = {}
    try:
        wl4 = Workload("4k", "H")
        k4 = min(K, 40)
        m4 = measure_gpu(wl4, k4, warmup, 3, local, 1, apply_pass=False, stage_pass=False)
        hbm, issue = rooflines(wl4, m4["remap_us"], None)
        d, e = _stats(m4["dev_ms"]), _stats(m4["e2e_ms"])
        out["4k60_handshake_1xB200"] = {
            "metric": wl4.metric, "workload": wl4.workload, "steps": k4, "windows": 3,
            "value": k4 / (d["median"] * 1e-3), "ms_per_step": d["median"] / k4,
            "e2e": {"value": k4 / (e["median"] * 1e-3), "h2d_bytes_per_step": wl4.frame_bytes(),
                    "d2h_bytes_per_step": wl4.frame_bytes()},
            "realtime_60fps_margin": k4 / (e["median"] * 1e-3) / 60.0,
            "roofline": hbm, "roofline_issue": issue, "parity_failures": m4["parity_fail"]}
        try:
            cb = cpu_baseline_sample(wl4, [m4["clip"][i] for i in range(32)], warm=12, count=20)
            out["4k60_handshake_1xB200"]["cpu_baseline"] = cb
        except Exception as ex:
            out["4k60_handshake_1xB200"]["cpu_baseline"] = {"value": None, "sample": f"failed: {ex}"}
    except Exception as ex:
        out["4k60_handshake_1xB200"] = {"failed": repr(ex)}
    try:
        from tools.synth import Clip
        wl0 = Workload("720p", "D", kind="pan", fps=30)
        clip = Clip("720p", "pan", frames=72, seed=42, fps=30)
        cb = cpu_baseline_sample(wl0, [clip[i] for i in range(72)], warm=12, count=60)
        out["720p30_pan_defaults_cpu"] = {"metric": wl0.metric, "workload": wl0.workload.replace("1 stream per GPU", "single "
                                          "stream, CPU/OpenCV path (no GPU)"), "value": cb["value"], "unit": "frames/s",
                                          "cpu_baseline": cb}
        # the same clip and settings on the GPU path, for the ratio
        m0 = measure_gpu(wl0, 60, 12, 1, local, 1, apply_pass=False, stage_pass=False)
        out["720p30_pan_defaults_cpu"]["gpu_value"] = 60 / (m0["dev_ms"][0] * 1e-3)
        out["720p30_pan_defaults_cpu"]["gpu_e2e"] = 60 / (m0["e2e_ms"][0] * 1e-3)
    except Exception as ex:
        out.setdefault("720p30_pan_defaults_cpu", {})["failed"] = repr(ex)
    return out


if __name__ == "__main__":
    main()
