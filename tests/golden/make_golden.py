"""Generates the committed golden fixtures with the oracle (run from the repo root: python tests/golden/make_golden.py).

The reference ships no golden vectors, and it cannot be compiled here, so these pin the ORACLE against
itself (regression) and give the GPU tests fixed inputs/outputs that do not need cv2 at test time.
Seeds are fixed; cv2 version is recorded in each file."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import cv2  # noqa: E402

from oracle import lvk_oracle as O  # noqa: E402
from tools.synth import Clip, make_canvas  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def textured(h, w, seed):
    c = make_canvas(w, h, seed)
    y0, x0 = (c.shape[0] - h) // 2, (c.shape[1] - w) // 2
    rng = np.random.default_rng(seed)
    img = c[y0:y0 + h, x0:x0 + w].copy()
    img[:, :, 1] = np.roll(img[:, :, 1], 2, axis=1)
    return np.clip(img.astype(np.int32) + rng.integers(-5, 6, size=img.shape), 0, 255).astype(np.uint8)


def main():
    O.build_native()
    ver = np.array(cv2.__version__)

    # ---- remap (FSR-EASU) ------------------------------------------------------------------------------------------
    src = textured(64, 96, 7)
    ts = np.stack([np.eye(3), np.array([[1, 0, 0.37], [0, 1, -0.81], [0, 0, 1.0]]),
                   np.array([[0.9995, -0.012, 1.9], [0.012, 0.9995, -1.3], [3e-5, -2e-5, 1.0]]),
                   np.array([[1.15, 0, -7.0], [0, 1.15, -5.0], [0, 0, 1.0]])])
    outs = np.stack([O.remap_homography(src, t, (255, 0, 255), yuv) for t in ts for yuv in (False, True)])
    mesh = (np.random.default_rng(1).standard_normal((4, 5, 2)) * 0.01).astype(np.float32)
    out_mesh = O.warp_mesh_apply(mesh, src, (0, 0, 0), False)
    np.savez_compressed(os.path.join(OUT, "remap_golden.npz"), src=src, transforms=ts, outputs=outs, mesh=mesh,
                        out_mesh=out_mesh, cv2=ver)

    # ---- detection image ---------------------------------------------------------------------------------------------
    rng = np.random.default_rng(2)
    f1 = rng.integers(0, 256, (180, 320, 3), dtype=np.uint8)
    f2 = rng.integers(0, 256, (187, 333, 3), dtype=np.uint8)
    np.savez_compressed(os.path.join(OUT, "detimg_golden.npz"), f1=f1, f2=f2,
                        d1_bgr=O.detection_image(f1, O.BGR, (80, 45)), d1_yuv=O.detection_image(f1, O.YUV, (80, 45)),
                        d1_rgb=O.detection_image(f1, O.RGB, (80, 45)), d2_bgr=O.detection_image(f2, O.BGR, (100, 60)),
                        d1_half=O.detection_image(f1, O.BGR, (160, 90)), cv2=ver)

    # ---- FAST ----------------------------------------------------------------------------------------------------------
    g = cv2.cvtColor(textured(96, 128, 3), cv2.COLOR_BGR2GRAY)
    fast = cv2.FastFeatureDetector_create(20, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    kps = fast.detect(g, None)
    kp_full = np.array([(k.pt[0], k.pt[1], k.response) for k in kps], dtype=np.float32).reshape(-1, 3)
    kps = fast.detect(g[8:80, 16:112], None)
    kp_roi = np.array([(k.pt[0], k.pt[1], k.response) for k in kps], dtype=np.float32).reshape(-1, 3)
    np.savez_compressed(os.path.join(OUT, "fast_golden.npz"), image=g, threshold=20, kp_full=kp_full,
                        roi=np.array([16, 8, 96, 72]), kp_roi=kp_roi, cv2=ver)

    # ---- LK (first and second call on one tracker object: the epsilon quirk) ----------------------------------------------
    clip = Clip((320, 240), "shake", frames=3, seed=5)
    a = cv2.cvtColor(clip[0], cv2.COLOR_BGR2GRAY)
    b = cv2.cvtColor(clip[1], cv2.COLOR_BGR2GRAY)
    kps = cv2.FastFeatureDetector_create(25, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16).detect(a, None)
    pts = (np.array([k.pt for k in kps], dtype=np.float32)[:400] + np.float32(0.25))
    pts = np.concatenate([pts, np.float32([[0, 0], [319, 239], [-4, 10], [160.5, 3.25]])])
    lk = cv2.SparsePyrLKOpticalFlow_create(winSize=(11, 11), maxLevel=3,
                                           crit=(cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 5, 0.01))
    o0, s0, _ = lk.calc(a, b, pts.reshape(-1, 1, 2), None)
    o1, s1, _ = lk.calc(a, b, pts.reshape(-1, 1, 2), None)
    np.savez_compressed(os.path.join(OUT, "lk_golden.npz"), prev=a, next=b, points=pts, out_call0=o0.reshape(-1, 2),
                        status_call0=s0.reshape(-1), out_call1=o1.reshape(-1, 2), status_call1=s1.reshape(-1), cv2=ver)

    # ---- pipeline trace (both presets, 480x270 frames, 16 frames) ---------------------------------------------------------
    for name, so in (("H", O.StabilizationSettings.obs_homography_preset()), ("D", O.StabilizationSettings())):
        clip = Clip((480, 270), "shake", frames=16, seed=11)
        flt = O.StabilizationFilter(so)
        flt.restart()
        rec = {"n_detected": [], "trust": [], "scene_quality": [], "stability": [], "has_output": [], "checksum": [],
               "correction": [], "fast_counts": []}
        last = None
        for i in range(16):
            out, ts_ = flt.apply(clip[i], O.BGR, i)
            tr = flt.trace
            rec["n_detected"].append(len(tr.get("detected", [])))
            rec["trust"].append(float(tr["trust"]))
            rec["scene_quality"].append(float(tr["scene_quality"]))
            rec["stability"].append(float(tr["stability"]))
            rec["has_output"].append(out is not None)
            rec["checksum"].append(int(out.astype(np.uint64).sum()) if out is not None else 0)
            rec["correction"].append(tr["correction"].reshape(-1))
            fc = list(tr.get("fast_counts", []))
            rec["fast_counts"].append(fc + [-2] * (4 - len(fc)))
            if out is not None:
                last = out
        np.savez_compressed(os.path.join(OUT, f"pipeline_{name}_golden.npz"), last_output=last, cv2=ver,
                            **{k: np.asarray(v) for k, v in rec.items()})
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
