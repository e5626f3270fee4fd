"""Micro-benchmark of the ScalingFilter kernels alone (device-resident frames, CUDA events on the library's stream):
k_rcas (lvk::sharpen) and the EASU upscale (lvk::upscale, MODE 2 of k_easu_remap), plus the two chained
(ScalingFilter::filter).  A ring of distinct frames larger than L2 in total keeps consecutive launches from re-hitting
the same lines.  Algorithmic bytes: rcas 6 B/px of the frame; upscale 3 B/px of the source + 3 B/px of the output;
the filter adds the intermediate frame's write + read (6 B/px of the output)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import livevisionkit_b200 as L
    from tools.synth import Clip, RESOLUTIONS

    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="1080p")
    ap.add_argument("--dst", default="4k")
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--ring", type=int, default=8)
    a = ap.parse_args()
    sw, sh = RESOLUTIONS[a.src]
    dw, dh = RESOLUTIONS[a.dst]
    clip = Clip(a.src, "shake", frames=a.ring)
    srcs = [torch.from_numpy(clip[i]).cuda() for i in range(a.ring)]
    ups = [torch.empty((dh, dw, 3), dtype=torch.uint8, device="cuda") for _ in range(a.ring)]
    outs = [torch.empty((dh, dw, 3), dtype=torch.uint8, device="cuda") for _ in range(a.ring)]
    s = L.Stream(None, 0)
    st = L.ScalingFilterSettings((dw, dh), 0.8, True)
    try:
        peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0

    def timed(fn, alg_bytes, name):
        for i in range(10):
            fn(i % a.ring)
        s.sync()
        s.event_record(0)
        for i in range(a.iters):
            fn(i % a.ring)
        s.event_record(1)
        ms = s.event_elapsed_ms(0, 1) / a.iters
        gbs = alg_bytes / (ms * 1e-3) / 1e9
        print(json.dumps({"kernel": name, "src": a.src, "dst": a.dst, "us": ms * 1e3, "algorithmic_GBps": gbs,
                          "peak_GBps": peak, "frac": gbs / peak}))

    timed(lambda i: s.upscale(srcs[i], (dw, dh), True, ups[i]), 3.0 * sw * sh + 3.0 * dw * dh, "k_easu_remap<scale> (lvk::upscale)")
    timed(lambda i: s.sharpen(ups[i], 0.8, outs[i]), 6.0 * dw * dh, "k_rcas (lvk::sharpen)")
    timed(lambda i: s.scaling_filter(srcs[i], st, outs[i]), 3.0 * sw * sh + 9.0 * dw * dh, "ScalingFilter::filter (upscale + sharpen)")


if __name__ == "__main__":
    main()
