"""Generates tests/golden/scaling_golden.npz with the oracle (run from the repo root:
python tests/golden/make_scaling_golden.py).  lvk::upscale / lvk::sharpen / ScalingFilter on a small textured frame with
all-zero and all-one patches (the RCAS limiter's 0 * inf cases).  The reference ships no vectors for this filter and
cannot be built here, so this pins the oracle against itself and gives the GPU test fixed inputs and outputs."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import lvk_oracle as O  # noqa: E402
from tools.synth import make_canvas  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    O.build_native()
    h, w = 72, 100
    c = make_canvas(w, h, 11)
    y0, x0 = (c.shape[0] - h) // 2, (c.shape[1] - w) // 2
    src = c[y0:y0 + h, x0:x0 + w].copy()
    src[5:14, 8:30] = 0
    src[40:52, 60:90] = 255
    src[20:30, 40:50] = np.random.default_rng(3).integers(0, 256, (10, 10, 3), dtype=np.uint8)
    size = (170, 123)  # x1.7 / x1.708: non-integer factors
    up_bgr, up_yuv = O.upscale(src, size, False), O.upscale(src, size, True)
    np.savez_compressed(os.path.join(OUT, "scaling_golden.npz"), src=src, size=np.array(size), up_bgr=up_bgr, up_yuv=up_yuv,
                        sharp_08=O.sharpen(src, 0.8), sharp_00=O.sharpen(src, 0.0), sharp_10=O.sharpen(src, 1.0),
                        filter_yuv_08=O.ScalingFilter(O.ScalingFilterSettings(size, 0.8, True)).apply(src))


if __name__ == "__main__":
    main()
