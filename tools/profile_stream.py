"""Drives one device-resident stabilization stream for a few frames: the target of the ncu captures in profiles/.

    ncu --set full --clock-control none --import-source on -k regex:'k_(ingest|pyramid|lk_track|ransac)' -s 150 -c 12 \
        -o gpurun_out/chain python tools/profile_stream.py --frames 40

Nothing here is timed: numbers printed under a profiler are never bench values.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402  (device memory plumbing)

import livevisionkit_b200 as L  # noqa: E402
from tools.synth import Clip  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=40)
    ap.add_argument("--resolution", default="1080p")
    ap.add_argument("--eager", action="store_true", help="per-stage events, eager launches instead of the graph")
    args = ap.parse_args()
    clip = Clip(args.resolution, "shake", frames=args.frames, seed=7)
    dev = torch.device("cuda", 0)
    frames = [torch.from_numpy(clip[i]).to(dev) for i in range(args.frames)]
    outs = [torch.empty_like(frames[0]) for _ in range(4)]
    flt = L.StabilizationFilter(L.StabilizationFilterSettings.obs_homography_preset(), device=0)
    if args.eager:
        flt.stream.set_profiling(True)
    for i, f in enumerate(frames):
        flt.stream.submit(f, outs[i % 4], L.BGR, i)
    flt.stream.sync()
    print("frames", args.frames, "launches", L._capi.load().lvkb200_kernel_launch_count())


if __name__ == "__main__":
    main()
