"""End-to-end parity, asserted (SURVEY 8(c) last row; VERDICT r1 'next' 1a/1b): >= 120 free-running output frames per
preset and resolution, the product's DEFAULT configuration (contract build of the EASU kernel, own RANSAC estimator,
device chain) against the oracle's FINAL pixels (cv2 USAC homography / Eigen-LSCG restatement + FSR restatement).

Per run the histogram of |delta| over all output bytes, the per-frame worst case and the inlier-mask agreement go to
gpurun_out/r02_parity_e2e_<preset>_<res>.json (the builder copies them to profiles/).  What is asserted is stated next
to each constant below; the bounds are the contract (<= 1 LSB) relaxed ONLY by what was measured to be the estimator
difference: our homography differs from cv2's USAC model by <= 0.25 px at detection resolution (different sampling),
which moves a few source positions across a texel boundary.
"""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

cv2 = pytest.importorskip("cv2")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# Fraction of the output bytes of EVERY frame that must lie within 1 LSB of the oracle's final pixels.  Measured on B200
# (profiles/r02_parity_e2e_*.json): worst frame H 0.99966 (1080p) / 0.99786 (4K), D 0.9998 / 0.99945 (2x2 mesh solved on
# the device: float32 summation order differs from the sequential restatement of Eigen), F 0.99196 / 0.99969; whole run
# 99.97-99.997 %.  The remainder are texel flips at high-contrast edges caused by the sub-0.01-px difference
# between our homography and cv2's USAC model (masks are identical on every frame) - not remap arithmetic: given the
# SAME transform the remap is within 1 LSB everywhere (tests/test_remap_gpu.py, tests/test_pipeline_gpu.py).
MIN_FRAC_WITHIN_1LSB = {"H": 0.997, "D": 0.999, "F": 0.99}
# fraction of all output bytes of the run that must be IDENTICAL (measured: H 0.9980 / 0.9907, D 0.998 / 0.9934, F 0.9876 / 0.9960)
MIN_FRAC_IDENTICAL = {"H": 0.985, "D": 0.99, "F": 0.98}
OUTPUT_FRAMES = 120


def _settings(oracle, L, preset):
    so = {"H": oracle.StabilizationSettings.obs_homography_preset, "D": oracle.StabilizationSettings,
          "F": oracle.StabilizationSettings.obs_field_preset}[preset]()
    sg = {"H": L.StabilizationFilterSettings.obs_homography_preset, "D": L.StabilizationFilterSettings,
          "F": L.StabilizationFilterSettings.obs_field_preset}[preset]()
    return so, sg


def _write(name, rec):
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, name), "w") as f:
        json.dump(rec, f, indent=1)


@pytest.mark.parametrize("preset,res", [("H", "1080p"), ("D", "1080p"), ("F", "1080p"), ("H", "4k"), ("D", "4k"), ("F", "4k")])
def test_free_running_final_pixels(oracle, preset, res):
    import livevisionkit_b200 as L
    from livevisionkit_b200 import _capi as K
    from tools.synth import Clip

    assert not L.remap_exact(), "this test measures the default (contract) build"
    so, sg = _settings(oracle, L, preset)
    delay = 10
    n_frames = OUTPUT_FRAMES + delay
    clip = Clip(res, "shake", frames=n_frames)
    ref, flt = oracle.StabilizationFilter(so), L.StabilizationFilter(sg, device=0)
    ref.restart()
    flt.restart()
    flt.stream.set_debug_capture(True)
    hist = np.zeros(8, dtype=np.int64)
    per_frame, mask_rows = [], []
    outputs = 0
    for i in range(n_frames):
        frame = clip[i]
        out_ref, _ = ref.apply(frame, oracle.BGR, i)
        vf = flt.apply(L.VideoFrame(frame, i, L.BGR))
        assert (out_ref is None) == vf.empty(), f"frame {i}: output cadence differs"
        tr = ref.trace
        if "inliers" in tr:
            inl = flt.stream.debug_fetch(K.DBG_INLIERS, np.uint8)
            a = np.asarray(tr["inliers"], dtype=np.uint8)
            same_len = len(inl) == len(a)
            mask_rows.append({"frame": i, "points_oracle": int(len(a)), "points_gpu": int(len(inl)),
                              "mismatches": int((inl != a).sum()) if same_len else None})
        if vf.empty():
            continue
        outputs += 1
        d = np.abs(out_ref.astype(np.int16) - vf.data.astype(np.int16))
        h = np.bincount(np.minimum(d.ravel(), 7), minlength=8)
        hist += h
        per_frame.append({"frame": i, "max": int(d.max()), "within_1lsb": float((h[0] + h[1]) / d.size),
                          "identical": float(h[0] / d.size)})
    total = int(hist.sum())
    rec = {"what": f"free-running {res} preset {preset}: final pixels of the default GPU path vs the oracle's final pixels, "
                   f"{outputs} output frames, histogram of |delta| over all output bytes (last bin: >= 7)",
           "preset": preset, "resolution": res, "output_frames": outputs, "bytes": total,
           "histogram_abs_delta": [int(v) for v in hist],
           "frac_identical": float(hist[0] / total), "frac_within_1lsb": float((hist[0] + hist[1]) / total),
           "worst_frame_within_1lsb": min(p["within_1lsb"] for p in per_frame),
           "max_abs_delta": max(p["max"] for p in per_frame),
           "mask_frames_compared": sum(1 for m in mask_rows if m["mismatches"] is not None),
           "mask_frames_with_mismatch": sum(1 for m in mask_rows if m["mismatches"]),
           "mask_mismatching_points": sum(m["mismatches"] or 0 for m in mask_rows),
           "mask_points": sum(m["points_oracle"] for m in mask_rows),
           "asserted": {"min_frac_within_1lsb_every_frame": MIN_FRAC_WITHIN_1LSB[preset],
                        "min_frac_identical_run": MIN_FRAC_IDENTICAL[preset]},
           "per_frame": per_frame, "masks": mask_rows}
    _write(f"r02_parity_e2e_{preset}_{res}.json", rec)
    print(f"[{preset} {res}] {outputs} frames: identical {rec['frac_identical']:.5f}, within 1 LSB {rec['frac_within_1lsb']:.6f} "
          f"(worst frame {rec['worst_frame_within_1lsb']:.6f}), max |d| {rec['max_abs_delta']}, hist {rec['histogram_abs_delta']}; "
          f"masks: {rec['mask_mismatching_points']} mismatching of {rec['mask_points']} points in "
          f"{rec['mask_frames_with_mismatch']}/{rec['mask_frames_compared']} frames")
    assert outputs >= OUTPUT_FRAMES
    # inlier masks: bit-exact on every frame whose point sets are still the same on both sides
    assert rec["mask_frames_compared"] >= 100 and rec["mask_mismatching_points"] == 0
    assert rec["worst_frame_within_1lsb"] >= MIN_FRAC_WITHIN_1LSB[preset], rec["worst_frame_within_1lsb"]
    assert rec["frac_identical"] >= MIN_FRAC_IDENTICAL[preset], rec["frac_identical"]


def _err(H, p, q):
    ph = np.concatenate([p, np.ones((len(p), 1))], axis=1) @ np.asarray(H, dtype=np.float64).T
    return np.linalg.norm(ph[:, :2] / ph[:, 2:] - q, axis=1)


def test_occluder_clip_inlier_masks(gpu_stream, oracle):
    """SURVEY 8(d) occluder variant (15 % independently moving block, seed 7): the oracle tracks the clip free-running;
    every frame's (tracked, matched) point set goes through cv2.findHomography(USAC_MAGSAC) AND through the GPU
    estimator.  Asserted: masks agree except for points whose reprojection error lies within 0.25 px of the acceptance
    threshold under either model (borderline), with at most 1e-3 of the points outside that band - measured: 132 of
    64 038 points differ, all in ONE frame where the block moves ~3 px (= the threshold) against the background and the
    two estimators settle on different sides of it (cv2 keeps 141 of 143 block points, we keep 12); 14 of them lie
    0.25-0.5 px from the threshold."""
    import livevisionkit_b200 as L
    from tools.synth import Clip
    so = oracle.StabilizationSettings.obs_homography_preset()
    clip = Clip("1080p", "occluder", frames=70)
    ref = oracle.StabilizationFilter(so)
    ref.restart()
    thr = float(so.acceptance_threshold)
    sxy = np.array([so.detection_resolution[0] / clip.width, so.detection_resolution[1] / clip.height])
    rows, frames_with_model = [], 0
    for i in range(70):
        ref.apply(clip[i], oracle.BGR, i)
        tr = ref.trace
        if "inliers" not in tr or "H" not in tr:
            continue
        p, q = np.asarray(tr["tracked"], dtype=np.float32), np.asarray(tr["matched"], dtype=np.float32)
        mc = np.asarray(tr["inliers"], dtype=np.uint8)
        try:
            Hg, mg = gpu_stream.find_homography(p, q, thr)
        except L.LvkB200Error:
            rows.append({"frame": i, "gpu_model": False})
            continue
        frames_with_model += 1
        e_c, e_g = _err(tr["H"], p, q), _err(Hg, p, q)
        borderline = (np.abs(e_c - thr) < 0.25) | (np.abs(e_g - thr) < 0.25)
        mism = mc != mg
        x, y, bw, bh = clip.occluder_rect(i - 1)  # tracked points live in the previous frame
        on_block = (p[:, 0] >= x * sxy[0]) & (p[:, 0] < (x + bw) * sxy[0]) & (p[:, 1] >= y * sxy[1]) & (p[:, 1] < (y + bh) * sxy[1])
        rows.append({"frame": i, "points": int(len(p)), "on_block": int(on_block.sum()), "inliers_cv2": int(mc.sum()),
                     "inliers_gpu": int(mg.sum()), "mismatches": int(mism.sum()), "borderline": int((mism & borderline).sum()),
                     "hard": int((mism & ~borderline).sum()), "block_inliers_cv2": int(mc[on_block].sum()),
                     "block_inliers_gpu": int(mg[on_block].sum())})
    _write("r02_parity_occluder_masks.json", {"what": "occluder clip (15 % moving block, seed 7), 1080p, OBS Homography "
           "preset: cv2 USAC_MAGSAC mask vs GPU estimator mask on the oracle's per-frame point sets", "frames": rows})
    hard = sum(r.get("hard", 0) for r in rows)
    mism = sum(r.get("mismatches", 0) for r in rows)
    pts = sum(r.get("points", 0) for r in rows)
    blk = sum(r.get("on_block", 0) for r in rows)
    print(f"occluder clip: {frames_with_model} frames, {pts} points ({blk} on the block), {mism} mask mismatches, {hard} not borderline; "
          f"block inliers cv2 {sum(r.get('block_inliers_cv2', 0) for r in rows)} gpu {sum(r.get('block_inliers_gpu', 0) for r in rows)}")
    assert frames_with_model >= 50 and blk > 0.05 * pts
    assert hard <= 1e-3 * pts and sum(1 for r in rows if r.get("hard")) <= 2
