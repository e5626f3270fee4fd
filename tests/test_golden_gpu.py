"""CUDA path (through the C-ABI) against the committed golden fixtures (tests/golden, made by the oracle with fixed
seeds).  No oracle or cv2 call at test time; plus size-independent properties at BASELINE.json's full sizes."""
import os

import numpy as np
import pytest

# bit-exact comparisons with the oracle restatement: these modules run the EXACT arithmetic build of the EASU kernels
# (tests that exercise the default contract build say so and switch it back on)
pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("exact_build")]

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    return np.load(os.path.join(G, name), allow_pickle=False)


def test_remap_golden(gpu_stream):
    g = _load("remap_golden.npz")
    k = 0
    for t in g["transforms"]:
        for yuv in (False, True):
            got = gpu_stream.remap_homography(g["src"], t, (255, 0, 255), yuv)
            assert (got == g["outputs"][k]).all(), f"transform {k // 2} yuv={yuv}"
            k += 1
    got = gpu_stream.remap_mesh(g["src"], g["mesh"], (0, 0, 0), False)
    assert np.abs(got.astype(np.int16) - g["out_mesh"].astype(np.int16)).max() <= 1


def test_detection_image_golden(gpu_stream):
    import livevisionkit_b200 as L
    g = _load("detimg_golden.npz")
    assert (gpu_stream.detection_image(g["f1"], L.BGR, (80, 45)) == g["d1_bgr"]).all()
    assert (gpu_stream.detection_image(g["f1"], L.YUV, (80, 45)) == g["d1_yuv"]).all()
    assert (gpu_stream.detection_image(g["f1"], L.RGB, (80, 45)) == g["d1_rgb"]).all()
    assert (gpu_stream.detection_image(g["f2"], L.BGR, (100, 60)) == g["d2_bgr"]).all()
    assert (gpu_stream.detection_image(g["f1"], L.BGR, (160, 90)) == g["d1_half"]).all()


def test_fast_golden(gpu_stream):
    g = _load("fast_golden.npz")
    h, w = g["image"].shape
    got = gpu_stream.fast_detect(g["image"], (0, 0, w, h), int(g["threshold"]))
    assert (np.stack([got["x"], got["y"], got["response"]], axis=1) == g["kp_full"]).all()
    roi = tuple(int(v) for v in g["roi"])
    got = gpu_stream.fast_detect(g["image"], roi, int(g["threshold"]))
    assert (np.stack([got["x"], got["y"], got["response"]], axis=1) == g["kp_roi"]).all()


def test_lk_golden(gpu_stream):
    g = _load("lk_golden.npz")
    w, h = g["prev"].shape[1], g["prev"].shape[0]
    for call in (0, 1):
        out, status = gpu_stream.lk_track(g["prev"], g["next"], g["points"], call_index=call)
        ref, rstat = g[f"out_call{call}"], g[f"status_call{call}"]
        p = out - 5.0
        final_oob = (np.floor(p[:, 0]) < -11) | (np.floor(p[:, 0]) >= w) | (np.floor(p[:, 1]) < -11) | (np.floor(p[:, 1]) >= h)
        assert (((status != rstat) & ~((rstat == 0) & (status == 1) & final_oob))).sum() == 0
        ok = (status == 1) & (rstat == 1)
        assert np.abs(out[ok] - ref[ok]).max() <= 0.01


@pytest.mark.parametrize("preset", ["H", "D"])
def test_pipeline_golden(preset):
    import livevisionkit_b200 as L
    from livevisionkit_b200 import _capi as K
    from tools.synth import Clip
    g = _load(f"pipeline_{preset}_golden.npz")
    s = L.StabilizationFilterSettings.obs_homography_preset() if preset == "H" else L.StabilizationFilterSettings()
    clip = Clip((480, 270), "shake", frames=16, seed=11)
    flt = L.StabilizationFilter(s, 0)
    flt.restart()
    flt.stream.set_debug_capture(True)
    for i in range(16):
        v = flt.apply(L.VideoFrame(clip[i], i, L.BGR))
        r = flt.last_result
        assert (not v.empty()) == bool(g["has_output"][i])
        kd = flt.stream.debug_fetch(K.DBG_DETECTED, np.dtype([("x", "f4"), ("y", "f4"), ("r", "f4"), ("c", "i4")]))
        assert (0 if kd is None else len(kd)) == int(g["n_detected"][i]), f"frame {i}"
        assert abs(r.trust_factor - g["trust"][i]) < 1e-6 and abs(r.tracking_stability - g["stability"][i]) < 1e-6
        c = flt.stream.debug_fetch(K.DBG_CORRECTION, np.float32)
        # corner displacement of the correction at frame resolution (normalized units x 480 px)
        assert np.abs(c - g["correction"][i]).max() * 480 <= 0.05, f"frame {i}"
    d = np.abs(v.data.astype(np.int16) - g["last_output"].astype(np.int16))
    print(f"[{preset}] last output vs golden: max |d| {int(d.max())}, within 1 LSB {float((d <= 1).mean()):.6f}")
    assert (d <= 1).mean() > 0.999


# ---- size-independent properties at full size (BASELINE.json configs 1-2: 1080p and 4K) -------------------------------


@pytest.mark.parametrize("res", ["1080p", "4k"])
def test_full_size_properties(gpu_stream, res):
    import livevisionkit_b200 as L
    from tools.synth import Clip
    frame = Clip(res, "shake", frames=1)[0]
    h, w = frame.shape[:2]
    # (1) remap: border band of an identity warp is a nearest-neighbour copy, interior is dering-clamped
    out = gpu_stream.remap_homography(frame, np.eye(3))
    assert (out[0] == frame[0]).all() and (out[:, 0] == frame[:, 0]).all()
    assert (out[-4:] == frame[-4:]).all() and (out[:, -4:] == frame[:, -4:]).all()
    f = frame.astype(np.int16)
    lo = np.minimum(np.minimum(f[1:-4, 1:-4], f[1:-4, 2:-3]), np.minimum(f[2:-3, 1:-4], f[2:-3, 2:-3]))
    hi = np.maximum(np.maximum(f[1:-4, 1:-4], f[1:-4, 2:-3]), np.maximum(f[2:-3, 1:-4], f[2:-3, 2:-3]))
    core = out[1:-4, 1:-4].astype(np.int16)
    assert (core >= lo - 1).all() and (core <= hi).all()
    # (2) remap: integer translation far from the border reproduces the identity-warp result shifted
    t = np.array([[1, 0, 8.0], [0, 1, 6.0], [0, 0, 1.0]])
    sh = gpu_stream.remap_homography(frame, t)
    assert (sh[16:-32, 16:-32] == out[22:-26, 24:-24]).all()
    # (3) remap is channel-wise equivariant for the BGR build (luma = channel 0 only steers the kernel)
    perm = np.ascontiguousarray(frame[:, :, [0, 2, 1]])
    outp = gpu_stream.remap_homography(perm, t)
    assert (outp[:-16, :-16][:, :, [0, 2, 1]] == sh[:-16, :-16]).all()  # (the background colour is not permuted)
    # (4) detection image: linear in a constant offset of the gray level (block mean of gray + c)
    det = gpu_stream.detection_image(frame, L.YUV, (480, 270))
    dim = (frame // 2).astype(np.uint8)
    det2 = gpu_stream.detection_image(dim, L.YUV, (480, 270))
    assert np.abs(det.astype(np.int16) // 2 - det2.astype(np.int16)).max() <= 1
    # (5) detection image is exactly the mean of integer blocks
    k = w // 480
    blocks = frame[:, :, 0].astype(np.int64).reshape(270, k, 480, k).sum(axis=(1, 3))
    assert (np.rint(blocks.astype(np.float32) * np.float32(1.0 / (k * k))).astype(np.uint8) == det).all()


def test_scaling_golden(gpu_stream):
    import livevisionkit_b200 as L
    g = _load("scaling_golden.npz")
    size = (int(g["size"][0]), int(g["size"][1]))
    assert (gpu_stream.upscale(g["src"], size, False) == g["up_bgr"]).all()
    assert (gpu_stream.upscale(g["src"], size, True) == g["up_yuv"]).all()
    for key, sharpness in (("sharp_08", 0.8), ("sharp_00", 0.0), ("sharp_10", 1.0)):
        assert (gpu_stream.sharpen(g["src"], sharpness) == g[key]).all(), key
    got = gpu_stream.scaling_filter(g["src"], L.ScalingFilterSettings(size, 0.8, True))
    assert (got == g["filter_yuv_08"]).all()
