"""Debug: list the pixels where the CUDA remap differs from the oracle for the test transforms."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))

import livevisionkit_b200 as L  # noqa: E402
from oracle import lvk_oracle as O  # noqa: E402
from test_remap_gpu import _textured, _transforms  # noqa: E402

s = L.Stream(L.StabilizationFilterSettings.obs_homography_preset(), 0)
for (w, h) in [(480, 270), (1280, 720)]:
    src = _textured(h, w, seed=w + h)
    for name, t in _transforms(w, h).items():
        ref = O.remap_homography(src, t, (255, 0, 255), False)
        got = s.remap_homography(src, t, (255, 0, 255), False)
        d = np.abs(ref.astype(np.int16) - got.astype(np.int16)).max(axis=2)
        ys, xs = np.nonzero(d)
        print(f"{w}x{h} {name}: {len(ys)} px differ, max {d.max()}")
        ti = np.linalg.inv(np.eye(3))  # t is already dst->src
        for y, x in list(zip(ys, xs))[:24]:
            p = t @ np.array([x, y, 1.0])
            sx, sy = p[0] / p[2], p[1] / p[2]
            print(f"   x={x} y={y} tile=({x // 32},{y // 16}) in-tile=({x % 32},{y % 16}) src=({sx:.4f},{sy:.4f}) "
                  f"ref={ref[y, x].tolist()} got={got[y, x].tolist()}")
