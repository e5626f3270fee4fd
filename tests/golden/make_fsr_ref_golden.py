"""Generates tests/golden/fsr_ref_golden.npz with the REFERENCE ITSELF: oracle/_ref/libfsrcl_ref_{strict,contract}.so are
the reference's FSR.cl compiled for the CPU where it lies (oracle/ref_build/build_ref.sh; needs /root/reference, so this
script only runs in the build container).  These vectors pin oracle/easu_ref.c (CPU tests) and the CUDA kernels
(GPU tests) to the reference's arithmetic.  Run from the repo root: python tests/golden/make_fsr_ref_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import fsr_ref as R  # noqa: E402
from tools.synth import make_canvas  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def textured(h, w, seed):
    c = make_canvas(w, h, seed)
    y0, x0 = (c.shape[0] - h) // 2, (c.shape[1] - w) // 2
    rng = np.random.default_rng(seed)
    img = c[y0:y0 + h, x0:x0 + w].copy()
    img[:, :, 1] = np.roll(img[:, :, 1], 2, axis=1)
    img[:, :, 2] = np.roll(img[:, :, 2], -3, axis=0)
    return np.clip(img.astype(np.int32) + rng.integers(-9, 10, size=img.shape), 0, 255).astype(np.uint8)


def transforms():
    a = np.radians(1.3)
    return np.stack([
        np.eye(3),
        np.array([[1, 0, 0.37], [0, 1, -0.81], [0, 0, 1.0]]),
        np.array([[0.9995, -0.012, 1.9], [0.012, 0.9995, -1.3], [3e-5, -2e-5, 1.0]]),
        np.array([[1.15, 0, -7.0], [0, 1.15, -5.0], [0, 0, 1.0]]),
        np.array([[0.93 * np.cos(a), -np.sin(a), 6.2], [np.sin(a), 0.93 * np.cos(a), -3.4], [-4e-5, 6e-5, 1.0]]),
        np.array([[1, 0, -30.5], [0, 1, 21.25], [0, 0, 1.0]]),  # a large part of the output is background / border band
    ])


def main():
    R.build()
    src = textured(72, 104, 7)
    ts = transforms()
    rec = {"src": src, "transforms": ts}
    for fl in R.FLAVOURS:
        rec[f"homography_{fl}"] = np.stack([R.remap_homography(src, t, (255, 0, 255), yuv, fl) for t in ts for yuv in (False, True)])
    rng = np.random.default_rng(3)
    coarse = rng.standard_normal((5, 6, 2)).astype(np.float32) * 2.5
    import cv2
    omap = cv2.resize(coarse, (90, 60), interpolation=cv2.INTER_LINEAR).astype(np.float32)  # 60x90 offsets -> dst 60x90
    rec["offset_map"] = omap
    for fl in R.FLAVOURS:
        rec[f"map_{fl}"] = np.stack([R.remap_map(src, omap, (0, 0, 0), yuv, fl) for yuv in (False, True)])
    small = textured(40, 56, 9)
    rec["scale_src"] = small
    sizes = [(112, 80), (83, 59), (56, 41), (150, 131)]
    rec["scale_sizes"] = np.array(sizes)
    for fl in R.FLAVOURS:
        for k, sz in enumerate(sizes):
            rec[f"scale{k}_{fl}"] = np.stack([R.upscale(small, sz, yuv, fl) for yuv in (False, True)])
    sharp = [0.0, 0.8, 1.0]
    rec["rcas_sharpness"] = np.array(sharp, dtype=np.float32)
    flat = src.copy()
    flat[10:30, 10:40] = 0      # all-zero and all-255 rings: the limiter's 0 * inf cases
    flat[40:60, 50:90] = 255
    rec["rcas_src"] = flat
    for fl in R.FLAVOURS:
        rec[f"rcas_{fl}"] = np.stack([R.sharpen(flat, s, fl) for s in sharp])
    np.savez_compressed(os.path.join(OUT, "fsr_ref_golden.npz"), **rec)
    print("fsr_ref_golden.npz written:", {k: v.shape for k, v in rec.items()})


if __name__ == "__main__":
    main()
