#!/usr/bin/env python
"""bench.py — stabilized frames/s of the LiveVisionKit stabilization path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA, one stream per GPU)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port) on host cores
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...   # N > 1: one rank per GPU

A "step" is one frame through the whole hot path (ingest -> FAST grid -> pyramidal LK -> homography RANSAC ->
path smoother -> EASU remap of the 10-frames-older frame).  Workload at N=1: BASELINE.json configs[1]
(1920x1080, 60 fps synthetic hand-shake clip, OBS "Homography" preset).  Per-GPU work is fixed as N grows
(one independent stream per GPU, seeds 42+rank) -> weak scaling; NCCL only gathers the counters.

value : whole-job fps with every input frame already resident in HBM and outputs left in HBM.
e2e   : the same metric through the public API with pinned HOST buffers (H2D of the frame and D2H of the result
        inside the timed region every step).
roofline: the dominant kernel (EASU remap): algorithmic bytes (6 B/px) / its average duration measured with CUDA
        events on the library's own CUDA stream inside the timed region, against MEASURED_PEAKS.json.
cpu_baseline: the oracle port (cv2 + scalar C EASU, all host cores) on a bounded sample of the same workload.
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

RES = "1080p"
WIDTH, HEIGHT = 1920, 1080
METRIC = "stabilized_frames_per_second_1080p"
WORKLOAD = ("1080p60 synthetic hand-shake sequence, OBS Homography preset (480x270 detection, FAST grid -> pyramidal LK "
            "-> homography RANSAC -> path smoother -> FSR-EASU remap), 1 stream per GPU")


PRESET, PRESET_NAME = "H", "OBS Homography"


def _select_workload(res, preset="H"):
    """BASELINE.json configs[1] (1080p60, the default and the configuration the metric is quoted on) or configs[2]
    (4K60, `--resolution 4k`); tracking preset H (OBS "Homography", default) or D (library defaults: 256x256
    detection, local motions -> LSCG mesh), the two presets SURVEY 8(d) asks for."""
    global RES, WIDTH, HEIGHT, METRIC, WORKLOAD, PRESET, PRESET_NAME
    if res == "4k":
        RES, WIDTH, HEIGHT = "4k", 3840, 2160
        METRIC = "stabilized_frames_per_second_4k"
        WORKLOAD = WORKLOAD.replace("1080p60", "4K60")
    if preset == "F":
        PRESET, PRESET_NAME = "F", "OBS Vector Field"
        WORKLOAD = WORKLOAD.replace("OBS Homography preset (480x270 detection, FAST grid -> pyramidal LK -> homography RANSAC",
                                    "OBS Vector Field preset (480x270 detection, FAST grid -> pyramidal LK -> 16x16 "
                                    "local-motion LSCG mesh")
    if preset == "D":
        PRESET, PRESET_NAME = "D", "library defaults"
        WORKLOAD = WORKLOAD.replace("OBS Homography preset (480x270 detection, FAST grid -> pyramidal LK -> homography RANSAC",
                                    "library-default settings (256x256 detection, FAST grid -> pyramidal LK -> local-motion "
                                    "LSCG mesh")


DEBLOCK = False


def _make_filter(L, settings, device):
    """The timed filter: lvk::StabilizationFilter, or with --deblock CompositeFilter{DeblockingFilter,
    StabilizationFilter} (BASELINE configs[4]) fused on the device."""
    flt = L.StabilizationFilter(settings, device=device)
    if DEBLOCK:
        flt.stream.set_deblocking(L.DeblockingFilterSettings())
    return flt


class _OracleChain:
    """The same chain on the CPU for the baseline legs."""

    def __init__(self, O, remap_threads):
        self.flt = O.StabilizationFilter(_oracle_settings(O), remap_threads=remap_threads)
        self.deblock = None
        if DEBLOCK:
            from oracle import deblock_oracle as D
            self.deblock = D.DeblockingFilter()

    def apply(self, frame, fmt, ts):
        if self.deblock is not None:
            frame = self.deblock.apply(frame, fmt)
        return self.flt.apply(frame, fmt, ts)


def _gpu_settings(L):
    S = L.StabilizationFilterSettings
    return {"H": S.obs_homography_preset, "F": S.obs_field_preset, "D": S}[PRESET]()


def _oracle_settings(O):
    S = O.StabilizationSettings
    return {"H": S.obs_homography_preset, "F": S.obs_field_preset, "D": S}[PRESET]()


def _peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons of one GPU during the timed region."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None
        self.halt = threading.Event()
        # NVML is initialised HERE (before the timed region starts), so the thread samples from its first millisecond
        self.nvml = None
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(index)
            self.nvml = (nv, h, nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None

    def _run_nvml(self):
        """NVML polled every ~2 ms: the timed region is only tens of milliseconds long, far below nvidia-smi's period."""
        if self.nvml is None:
            raise RuntimeError("NVML unavailable")
        nv, h, smax = self.nvml
        bits = (("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap))
        while not self.halt.is_set():
            sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
            mask = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
            self.rows.append([str(sm), str(smax), "0"] + ["Active" if mask & b else "Not Active" for _, b in bits])
            time.sleep(0.001)

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
                "samples": len(sm)}


def _dist_env():
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    return rank, local, world


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port: cv2 4.13 + scalar C EASU) on the host cores."""
    rank, _, world = _dist_env()
    if rank != 0:
        return  # rank 0 alone runs and prints; the others exit 0 without work
    import cv2
    from oracle import lvk_oracle as O
    from tools.synth import Clip
    O.build_native()
    cores = os.cpu_count() or 1
    cv2.setNumThreads(cores)
    clip = Clip(RES, "shake", frames=args.warmup + args.steps, seed=42)
    frames = [clip[i] for i in range(len(clip))]
    flt = _OracleChain(O, cores)
    for i in range(args.warmup):
        flt.apply(frames[i], O.BGR, i)
    t0 = time.perf_counter()
    for i in range(args.warmup, args.warmup + args.steps):
        flt.apply(frames[i], O.BGR, i)
    dt = time.perf_counter() - t0
    fps = args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "resolution": [WIDTH, HEIGHT], "preset": PRESET_NAME},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} frames after {args.warmup} warm-up frames, single stream, "
                                   f"cv2 {cv2.__version__} (reference pins 4.8.0) + scalar C EASU on {cores} threads"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_sample(frames, warm=12, count=60):
    """Bounded CPU sample of the same workload (oracle port), rank 0 at N=1 only."""
    import cv2
    from oracle import lvk_oracle as O
    O.build_native()
    cores = os.cpu_count() or 1
    cv2.setNumThreads(cores)
    flt = _OracleChain(O, cores)
    count = min(count, len(frames) - warm)
    for i in range(warm):
        flt.apply(frames[i], O.BGR, i)
    t0 = time.perf_counter()
    for i in range(warm, warm + count):
        flt.apply(frames[i], O.BGR, i)
    dt = time.perf_counter() - t0
    return {"value": count / dt, "unit": "frames/s", "cores": cores, "kind": "port",
            "sample": f"{count} frames of the same {RES} clip after {warm} warm-up frames, oracle port "
                      f"(cv2 {cv2.__version__} + scalar C EASU, {cores} threads)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=30)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
    _select_workload(args.resolution, args.preset)
    if args.deblock:
        global DEBLOCK, WORKLOAD
        DEBLOCK = True
        WORKLOAD = "DeblockingFilter (defaults) -> " + WORKLOAD

    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    import livevisionkit_b200 as L
    from tools.synth import Clip

    rank, local, world = _dist_env()
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

    # ---- synthetic clip (distinct frames: 330 x 6.2 MB = 2.05 GB > L2, so consecutive steps never re-hit lines)
    n_frames = args.warmup + args.steps
    from tools.scaling import stream_seed
    clip = Clip(RES, "shake", frames=n_frames, seed=stream_seed(rank))
    host_frames = [clip[i] for i in range(n_frames)]
    settings = _gpu_settings(L)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ======== pass 1: device-resident (value + roofline) ========
    dev_frames = [torch.from_numpy(f).to(dev) for f in host_frames]
    out_ring = [torch.empty_like(dev_frames[0]) for _ in range(16)]
    torch.cuda.synchronize()
    flt = _make_filter(L, settings, local)
    s = flt.stream
    # Each frame is announced one step ahead (lvkb200_stream_prefetch_frame, the input thread of VideoFilter::stream
    # running ahead of the filter thread): its copy into the stream's ring and its detection image + pyramid are queued
    # behind the current frame's tracking chain.  All of a frame's work still happens inside the timed region.
    lookahead = not args.no_lookahead
    # pointer / pitch / geometry of the (reused) buffers are looked up once, not per call (L.FrameRef)
    dev_refs = [L.FrameRef(f) for f in dev_frames]
    out_refs = [L.FrameRef(o) for o in out_ring]
    for i in range(args.warmup):
        if lookahead and i + 1 < n_frames:
            s.prefetch(dev_refs[i + 1], L.BGR)
        s.submit(dev_refs[i], out_refs[i % 16], L.BGR, i)
    s.sync()
    s.stage_totals_us(reset=True)
    launches0 = L._capi.load().lvkb200_kernel_launch_count()
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    s.event_record(0)
    t0 = time.perf_counter()
    outputs = 0
    for i in range(args.warmup, n_frames):
        if lookahead and i + 1 < n_frames:
            s.prefetch(dev_refs[i + 1], L.BGR)
        r = s.submit(dev_refs[i], out_refs[i % 16], L.BGR, i)
        outputs += r.has_output
    s.event_record(1)
    s.sync()
    barrier()
    wall = time.perf_counter() - t0
    dev_ms = s.event_elapsed_ms(0, 1)
    clocks = sampler.stop()
    launches = L._capi.load().lvkb200_kernel_launch_count() - launches0
    totals, counts = s.stage_totals_us(reset=True)  # default mode: only the remap kernel is event-timed
    last_dev_out = out_ring[(n_frames - 1) % 16].cpu().numpy().copy()
    # untimed extra pass with per-stage CUDA events (eager launches instead of the tracking graph) for stage_us
    prof = _make_filter(L, settings, local)
    prof.stream.set_profiling(True)
    n_prof = min(n_frames, 90)
    for i in range(n_prof):
        prof.stream.submit(dev_frames[i], out_ring[i % 16], L.BGR, i)
        if i == 29:
            prof.stream.stage_totals_us(reset=True)
    ptotals, pcounts = prof.stream.stage_totals_us(reset=True)
    prof.stream.close()
    del dev_frames, dev_refs
    torch.cuda.empty_cache()

    # ======== pass 2: end to end through the public API with pinned host buffers ========
    pinned_in = [torch.from_numpy(f).pin_memory() for f in host_frames]
    pinned_out = [torch.empty_like(pinned_in[0]).pin_memory() for _ in range(4)]
    flt2 = _make_filter(L, settings, local)
    for i in range(args.warmup):
        flt2.apply(L.VideoFrame(pinned_in[i], i, L.BGR), output=pinned_out[i % 4])
    barrier()
    flt2.stream.event_record(0)
    t1 = time.perf_counter()
    for i in range(args.warmup, n_frames):
        flt2.apply(L.VideoFrame(pinned_in[i], i, L.BGR), output=pinned_out[i % 4])
    flt2.stream.event_record(1)
    flt2.stream.sync()
    barrier()
    wall_e2e = time.perf_counter() - t1
    apply_ms = flt2.stream.event_elapsed_ms(0, 1)
    parity_fail = int(not np.array_equal(pinned_out[(n_frames - 1) % 4].numpy(), last_dev_out))

    # ======== pass 2b: end to end through the pipelined public API (VideoFilter::stream analogue) ========
    # upload of frame t+1 and download of output t-1 overlap the processing of frame t; still one H2D of the input and
    # one D2H of the result per step, all inside the timed region.
    flt3 = _make_filter(L, settings, local)
    pin_refs = [L.FrameRef(t) for t in pinned_in]
    pout_refs = [L.FrameRef(t) for t in pinned_out[:3]]
    warm = [L.VideoFrame(pin_refs[i], i, L.BGR) for i in range(args.warmup)]
    timed = [L.VideoFrame(pin_refs[i], i, L.BGR) for i in range(args.warmup, n_frames)]
    sink = []
    flt3.stream(warm, lambda vf: False, pout_refs)
    barrier()
    tp = time.perf_counter()
    flt3.stream.event_record(0)
    delivered = flt3.stream(timed, lambda vf: sink.append(vf.timestamp), pout_refs)
    flt3.stream.event_record(1)
    flt3.stream.sync()
    barrier()
    wall_pipe = time.perf_counter() - tp
    e2e_ms = wall_pipe * 1e3  # host wall clock: the last download completes on the copy-out stream, not on `cs`
    parity_fail += int(not np.array_equal(pinned_out[(len(timed) - 1) % 3].numpy(), last_dev_out))
    parity_fail += int(delivered != args.steps)

    # ======== reduce over ranks: max time, summed frames (one all_gather of the counter struct over NCCL) ========
    from tools import scaling
    mine = torch.tensor([float(args.steps), dev_ms, e2e_ms, float(launches), float(parity_fail), wall * 1e3,
                         wall_e2e * 1e3, float(outputs)], dtype=torch.float64, device=dev)
    agg = scaling.aggregate(scaling.gather_counters(mine))
    if rank == 0:
        t_dev, t_e2e = agg["dev_ms"], agg["e2e_ms"]
        value, e2e = agg["value_fps"], agg["e2e_fps"]
        peaks, peak_kind = _peaks()
        peak = float(peaks.get("hbm_gbs", 6650.0))
        remap_us = totals["remap"] / max(counts["remap"], 1)
        alg_bytes = 6.0 * WIDTH * HEIGHT
        achieved = alg_bytes / (remap_us * 1e-6) / 1e9 if remap_us > 0 else 0.0
        traffic = None
        try:  # per-launch DRAM bytes of the remap kernel from the committed ncu capture, if present
            prof = json.load(open(os.path.join(ROOT, "profiles", "remap_traffic.json")))
            traffic = prof.get(f"dram_bytes_per_launch_{RES}")
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_dev / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "resolution": [WIDTH, HEIGHT], "preset": PRESET_NAME,
                       "streams_per_gpu": 1,
                       "l2": f"{n_frames} distinct frames ({n_frames * WIDTH * HEIGHT * 3 / 1e9:.2f} GB per GPU) "
                             f"streamed once each: inputs larger than L2, no flush needed",
                       "timing": "CUDA events on the library's CUDA stream around the K timed submits, max over ranks",
                       "lookahead": not args.no_lookahead},
            "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": WIDTH * HEIGHT * 3,
                    "d2h_bytes_per_step": WIDTH * HEIGHT * 3, "ms_per_step": t_e2e / args.steps,
                    "api": "StabilizationFilter.stream(frames, callback) — the pipelined VideoFilter::stream analogue: "
                           "pinned host input -> H2D -> filter -> D2H -> pinned host output for every step, uploads and "
                           "downloads overlapped with the neighbouring frames' processing; host wall clock between "
                           "barrier+synchronize pairs, max over ranks",
                    "apply_fps_rank0": args.steps / (apply_ms * 1e-3),
                    "apply_note": "same, through the synchronous per-frame StabilizationFilter.apply (no overlap)"},
            "gpu_launches": agg["launches"],
            "roofline": {"bound": "hbm", "kernel": "k_easu_remap<homography>", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_kind,
                         "avg_kernel_us": remap_us, "algorithmic_bytes_per_launch": alg_bytes,
                         "note": "EASU is instruction-issue bound (471 executed thread-instructions/px, 61% issue "
                                 "utilisation, DRAM 2% busy: profiles/r01_remap_1080p_committed_ncu_full.json), not HBM "
                                 "bound; see DESIGN.md 5.1"},
            "stage_us": {k: (ptotals[k] / pcounts[k] if pcounts[k] else 0.0) for k in ptotals},
            "stage_us_note": "separate untimed pass with per-stage CUDA events (profiling mode, eager launches)",
            "clocks": clocks,
            "outputs": agg["outputs"],
            "parity_failures": agg["parity_failures"],
            "wall_ms_per_step": agg["wall_ms"] / args.steps,
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = cpu_baseline_sample(host_frames)
            except Exception as e:  # the baseline must never take the bench down
                line["cpu_baseline"] = {"value": None, "unit": "frames/s", "cores": os.cpu_count(), "kind": "port",
                                        "sample": f"failed: {e}"}
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
