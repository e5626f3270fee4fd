"""Micro-benchmarks of the rows either side of the stabilizer (device-resident buffers, CUDA events on the library's
stream, a ring of distinct frames larger than L2): the stand-alone DeblockingFilter and the OBS plane-layout
ingest / egress kernels.  Prints one JSON line per kernel group with its algorithmic bytes and the HBM fraction."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import numpy as np
    import torch
    import livevisionkit_b200 as L
    from tools.synth import Clip, RESOLUTIONS
    from tools.profile_deblock import blocky

    ap = argparse.ArgumentParser()
    ap.add_argument("--res", default="4k")
    ap.add_argument("--iters", type=int, default=100)
    ap.add_argument("--ring", type=int, default=6)
    a = ap.parse_args()
    w, h = RESOLUTIONS[a.res]
    try:
        peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0
    s = L.Stream(None, 0)
    clip = Clip(a.res, "shake", frames=a.ring, seed=3)
    frames = [torch.from_numpy(blocky(clip[i], seed=i)).cuda() for i in range(a.ring)]

    def timed(fn, alg_bytes, name):
        for i in range(6):
            fn(i % a.ring)
        s.sync()
        s.event_record(0)
        for i in range(a.iters):
            fn(i % a.ring)
        s.event_record(1)
        ms = s.event_elapsed_ms(0, 1) / a.iters
        gbs = alg_bytes / (ms * 1e-3) / 1e9
        print(json.dumps({"kernel": name, "res": a.res, "us": ms * 1e3, "algorithmic_GBps": gbs, "peak_GBps": peak,
                          "frac": gbs / peak}))

    # DeblockingFilter in place: analysis reads 3 B/px, blend reads + writes 3 B/px
    timed(lambda i: s.deblock(frames[i], None, L.BGR, out=frames[i]), 9.0 * w * h,
          "DeblockingFilter (k_deblock_analyse + _median + _blend)")

    # FrameIngest: NV12 / I420 planes -> packed YUV (1.5 B/px in, 3 B/px out) and back
    for fmt, shapes in (("NV12", [(h, w), (h // 2, w)]), ("I420", [(h, w), (h // 2, w // 2), (h // 2, w // 2)])):
        planes = [[torch.randint(0, 256, sh, dtype=torch.uint8, device="cuda") for sh in shapes] for _ in range(a.ring)]
        packed = [torch.empty((h, w, 3), dtype=torch.uint8, device="cuda") for _ in range(a.ring)]
        ing = L.FrameIngest.Select(fmt, stream=s)
        obs = [L.ObsFrame(fmt, w, h, planes[i]) for i in range(a.ring)]
        timed(lambda i: ing.upload_obs_frame(obs[i], out=packed[i]), 4.5 * w * h, f"FrameIngest upload {fmt} (k_planes_to_packed)")
        vfs = [L.VideoFrame(packed[i], 0, L.YUV) for i in range(a.ring)]
        timed(lambda i: ing.download_ocl_frame(vfs[i], obs[i]), 4.5 * w * h, f"FrameIngest download {fmt} (k_packed_to_planes)")


if __name__ == "__main__":
    main()
