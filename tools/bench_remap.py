"""Micro-benchmark of the EASU remap kernel alone (device-resident frames, CUDA events on the library's stream).
Rotates through a ring of distinct source/destination frames larger than L2 in total (11-frame ring like the
stabilizer's queue) so consecutive launches do not re-hit the same lines."""
import argparse
import json
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import livevisionkit_b200 as L
    from tools.synth import Clip, RESOLUTIONS

    ap = argparse.ArgumentParser()
    ap.add_argument("--res", default="4k")
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--ring", type=int, default=11)
    ap.add_argument("--exact", action="store_true", help="time the exact arithmetic build (remap.cu) instead of the default")
    a = ap.parse_args()
    w, h = RESOLUTIONS[a.res]
    L.set_remap_exact(a.exact)
    clip = Clip(a.res, "shake", frames=a.ring)
    srcs = [torch.from_numpy(clip[i]).cuda() for i in range(a.ring)]
    dsts = [torch.empty_like(srcs[0]) for _ in range(a.ring)]
    s = L.Stream(L.StabilizationFilterSettings.obs_homography_preset(), 0)
    torch.cuda.synchronize()
    ang = math.radians(0.3)
    c, sn = math.cos(ang), math.sin(ang)
    t = np.array([[c, -sn, 3.7], [sn, c, -2.2], [1e-7, -1e-7, 1.0]])
    for i in range(10):
        s.remap_homography(srcs[i % a.ring], t, out=dsts[i % a.ring])
    s.sync()
    s.event_record(0)
    for i in range(a.iters):
        s.remap_homography(srcs[i % a.ring], t, out=dsts[i % a.ring])
    s.event_record(1)
    ms = s.event_elapsed_ms(0, 1) / a.iters
    bytes_alg = 6.0 * w * h
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = peaks.get("hbm_gbs", 6650.0)
    gbs = bytes_alg / (ms * 1e-3) / 1e9
    print(json.dumps({"kernel": "k_easu_remap<homography>" if a.exact else "k_easu_remap_fast<homography>", "res": a.res, "us": ms * 1e3, "algorithmic_GBps": gbs,
                      "peak_GBps": peak, "frac": gbs / peak, "Mpx_per_s": w * h / (ms * 1e-3) / 1e6}))


if __name__ == "__main__":
    main()
