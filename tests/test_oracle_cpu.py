"""Oracle pinning (CPU): the oracle reproduces the committed golden fixtures and satisfies size-independent
properties of the domain.  The reference's own tests hold NO vectors for this path and the reference
cannot be built here, so the oracle is 'parity unpinned' against the real reference; these tests pin it against
itself and against first-principles properties."""
import os

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    return np.load(os.path.join(G, name), allow_pickle=False)


def test_remap_golden(oracle):
    g = _load("remap_golden.npz")
    k = 0
    for t in g["transforms"]:
        for yuv in (False, True):
            assert (oracle.remap_homography(g["src"], t, (255, 0, 255), yuv) == g["outputs"][k]).all()
            k += 1
    assert (oracle.warp_mesh_apply(g["mesh"], g["src"], (0, 0, 0), False) == g["out_mesh"]).all()


def test_remap_properties(oracle):
    rng = np.random.default_rng(0)
    # a constant image stays constant under any warp that stays inside it (weights are normalised, dering clamps)
    const = np.full((80, 120, 3), 137, dtype=np.uint8)
    out = oracle.remap_homography(const, np.array([[1, 0.01, 1.5], [-0.01, 1, -0.7], [0, 0, 1.0]]), (9, 9, 9))
    inside = out[6:-8, 6:-8]
    assert np.abs(inside.astype(int) - 137).max() <= 1
    # the border band of an identity warp is nearest-neighbour (FSR.cl:436-448), the interior is filtered
    img = rng.integers(0, 256, (64, 96, 3), dtype=np.uint8)
    idt = oracle.remap_homography(img, np.eye(3))
    assert (idt[0] == img[0]).all() and (idt[:, 0] == img[:, 0]).all() and (idt[-4:] == img[-4:]).all()
    assert (idt[:, -4:] == img[:, -4:]).all() and (idt[1:-4, 1:-4] != img[1:-4, 1:-4]).any()
    # the output of EASU is clamped to the min/max of the 4 nearest texels (dering)
    lo = np.minimum(np.minimum(img[1:-4, 1:-4], img[1:-4, 2:-3]), np.minimum(img[2:-3, 1:-4], img[2:-3, 2:-3]))
    hi = np.maximum(np.maximum(img[1:-4, 1:-4], img[1:-4, 2:-3]), np.maximum(img[2:-3, 1:-4], img[2:-3, 2:-3]))
    core = idt[1:-4, 1:-4].astype(int)
    assert (core >= lo.astype(int) - 1).all() and (core <= hi).all()
    # everything outside the source is background
    far = oracle.remap_homography(img, np.array([[1, 0, 500.0], [0, 1, 0], [0, 0, 1.0]]), (1, 2, 3))
    assert (far == np.array([1, 2, 3], dtype=np.uint8)).all()
    # the threaded row split does not change results
    t = np.array([[1.001, 0.004, -2.2], [-0.004, 1.001, 1.4], [1e-6, 0, 1.0]])
    assert (oracle.remap_homography(img, t, threads=1) == oracle.remap_homography(img, t, threads=5)).all()


def test_detection_image_golden(oracle):
    g = _load("detimg_golden.npz")
    assert (oracle.detection_image(g["f1"], oracle.BGR, (80, 45)) == g["d1_bgr"]).all()
    assert (oracle.detection_image(g["f1"], oracle.YUV, (80, 45)) == g["d1_yuv"]).all()
    assert (oracle.detection_image(g["f1"], oracle.RGB, (80, 45)) == g["d1_rgb"]).all()
    assert (oracle.detection_image(g["f2"], oracle.BGR, (100, 60)) == g["d2_bgr"]).all()
    assert (oracle.detection_image(g["f1"], oracle.BGR, (160, 90)) == g["d1_half"]).all()
    # first-principles restatement of the integer-factor path: BGR->Y fixed point, block mean, round-half-even
    f = g["f1"].astype(np.int64)
    gray = (f[..., 0] * 3735 + f[..., 1] * 19235 + f[..., 2] * 9798 + (1 << 14)) >> 15
    blocks = gray.reshape(45, 4, 80, 4).sum(axis=(1, 3))
    assert (np.rint(blocks.astype(np.float32) * np.float32(1 / 16)).astype(np.uint8) == g["d1_bgr"]).all()
    half = gray.reshape(90, 2, 160, 2).sum(axis=(1, 3))
    assert (((half + 2) >> 2).astype(np.uint8) == g["d1_half"]).all()  # OpenCV's 2x2 special case


def test_fast_and_lk_goldens():
    g = _load("fast_golden.npz")
    det = cv2.FastFeatureDetector_create(int(g["threshold"]), True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    kp = np.array([(k.pt[0], k.pt[1], k.response) for k in det.detect(g["image"], None)], dtype=np.float32)
    assert (kp.reshape(-1, 3) == g["kp_full"]).all()
    assert (np.diff(g["kp_full"][:, 1]) >= 0).all()  # (y, x) emission order
    assert g["kp_full"][:, 0].min() >= 3 and g["kp_full"][:, 1].min() >= 3  # 3-px border never fires
    l = _load("lk_golden.npz")
    lk = cv2.SparsePyrLKOpticalFlow_create(winSize=(11, 11), maxLevel=3,
                                           crit=(cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 5, 0.01))
    o0, s0, _ = lk.calc(l["prev"], l["next"], l["points"].reshape(-1, 1, 2), None)
    o1, s1, _ = lk.calc(l["prev"], l["next"], l["points"].reshape(-1, 1, 2), None)
    assert (s0.reshape(-1) == l["status_call0"]).all() and (s1.reshape(-1) == l["status_call1"]).all()
    assert np.abs(o0.reshape(-1, 2) - l["out_call0"]).max() <= 1e-4
    assert np.abs(o1.reshape(-1, 2) - l["out_call1"]).max() <= 1e-4
    ok = (l["status_call0"] == 1) & (l["status_call1"] == 1)
    assert np.abs(l["out_call0"][ok] - l["out_call1"][ok]).max() > 1e-4  # the in-place squared epsilon is real


@pytest.mark.parametrize("preset", ["H", "D"])
def test_pipeline_golden(oracle, preset):
    from tools.synth import Clip
    g = _load(f"pipeline_{preset}_golden.npz")
    so = oracle.StabilizationSettings.obs_homography_preset() if preset == "H" else oracle.StabilizationSettings()
    clip = Clip((480, 270), "shake", frames=16, seed=11)
    flt = oracle.StabilizationFilter(so)
    flt.restart()
    for i in range(16):
        out, ts = flt.apply(clip[i], oracle.BGR, i)
        tr = flt.trace
        assert (out is not None) == bool(g["has_output"][i])
        assert len(tr.get("detected", [])) == int(g["n_detected"][i])
        assert abs(float(tr["trust"]) - g["trust"][i]) < 1e-6
        assert np.abs(tr["correction"].reshape(-1) - g["correction"][i]).max() < 1e-6
        if out is not None:
            assert ts == i - 10
            assert int(out.astype(np.uint64).sum()) == int(g["checksum"][i])
    assert (out == g["last_output"]).all()
    # cadence: predictive_samples empty frames, then one output per input
    assert list(g["has_output"]) == [False] * 10 + [True] * 6


def test_smoother_is_gaussian_lowpass_of_the_path(oracle):
    """First-principles check of PathSmoother::next: correction = (Gaussian-weighted path) - (current position)."""
    s = oracle.StabilizationSettings.obs_homography_preset()
    sm = oracle.PathSmoother(s)
    rng = np.random.default_rng(1)
    motions = [(rng.standard_normal((2, 2, 2)) * 1e-3).astype(np.float32) for _ in range(60)]
    for m in motions:
        sigma = sm.base_smoothing + sm.smoothing_factor
        c = sm.next(m)
        traj = np.stack(sm.traj).astype(np.float64)  # 21 newest motions, oldest first
        path = np.cumsum(traj, axis=0)                # cumulative path over the window
        g = cv2.getGaussianKernel(21, sigma, cv2.CV_64F).reshape(-1)
        smooth = np.tensordot(g, path, axes=(0, 0))
        expect = smooth - path[10]
        assert np.abs(np.clip(expect, -0.05, 0.05) - c).max() < 5e-6


def test_ring_buffer_delay_semantics(oracle):
    """StreamBuffer / m_FrameQueue: the output at call t is the frame pushed at t - predictive_samples."""
    s = oracle.StabilizationSettings.obs_homography_preset()
    s.stabilize_output = False
    f = oracle.StabilizationFilter(s)
    frames = [np.full((32, 48, 3), i, dtype=np.uint8) for i in range(14)]
    for i, fr in enumerate(frames):
        out, ts = f.apply(fr, oracle.BGR, 100 + i)
        if i < 10:
            assert out is None
        else:
            assert (out == frames[i - 10]).all() and ts == 100 + i - 10
