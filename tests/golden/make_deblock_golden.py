"""Generates tests/golden/deblock_golden.npz with the deblocking oracle (run from the repo root:
python tests/golden/make_deblock_golden.py).  Same status as the other goldens: the reference ships no vectors and
cannot be built here, so this pins the oracle's restated arithmetic against itself (the cv2 calls the reference makes
agree with it to <= 1 LSB, tests/test_deblock_cpu.py) and gives the GPU tests an input/output pair that needs no cv2."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import cv2  # noqa: E402

from oracle import deblock_oracle as D  # noqa: E402
from tools.synth import Clip  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    frame = D.blocky_frame(Clip((170, 120), "shake", frames=2, seed=11)[1], 16, 0.8, seed=4)  # region 160 x 112
    np.savez_compressed(os.path.join(OUT, "deblock_golden.npz"), frame=frame, out_bgr=D.deblock_restated(frame, 0),
                        out_yuv=D.deblock_restated(frame, 4), cv2=np.array(cv2.__version__))
    print("deblock_golden.npz:", frame.shape, "changed", float((D.deblock_restated(frame, 0) != frame).mean()))


if __name__ == "__main__":
    main()
