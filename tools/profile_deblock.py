"""Drives the stand-alone DeblockingFilter on a device-resident frame: the target of the ncu captures in profiles/.

    ncu --set full --clock-control none --import-source on -k regex:k_deblock -c 3 -o gpurun_out/deblock \
        python tools/profile_deblock.py --resolution 1080p

The frame gets synthetic compression artefacts (a share of the macroblocks pulled to their mean) so that all three
blend cases occur: untouched, partially and fully smoothed blocks.  Nothing here is timed.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402
import torch  # noqa: E402  (device memory plumbing)

import livevisionkit_b200 as L  # noqa: E402
from tools.synth import Clip  # noqa: E402


def blocky(frame, block=16, seed=0):
    rng = np.random.default_rng(seed)
    h, w = frame.shape[:2]
    ey, ex = h // block, w // block
    out = frame.astype(np.float32)
    roi = out[:ey * block, :ex * block].reshape(ey, block, ex, block, -1)
    mean = roi.mean(axis=(1, 3), keepdims=True)
    wgt = np.where(rng.random((ey, 1, ex, 1, 1)) < 0.4, 1.0, rng.random((ey, 1, ex, 1, 1)) * 0.6).astype(np.float32)
    roi[...] = roi * (1 - wgt) + mean * wgt
    return np.clip(np.rint(out), 0, 255).astype(np.uint8)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--resolution", default="1080p")
    ap.add_argument("--repeat", type=int, default=3)
    args = ap.parse_args()
    frame = blocky(Clip(args.resolution, "shake", frames=1, seed=3)[0])
    s = L.Stream(None, 0)
    src = torch.from_numpy(frame).cuda()
    for _ in range(args.repeat):
        work = src.clone()
        s.deblock(work, None, L.BGR, out=work)
    s.sync()
    changed = float((work.cpu().numpy() != frame).mean())
    print(f"{args.resolution}: filter changed {100 * changed:.1f}% of the bytes")


if __name__ == "__main__":
    main()
