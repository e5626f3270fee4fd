"""Stage-isolated parity of the tracking kernels (through the C-ABI) against the cv2-based oracle:
K0 detection image (bit-exact), K2 FAST keypoints (coordinates, order, response: bit-exact),
K1+K4 pyramidal LK (status bit-exact, positions <= 0.01 px), K6a homography RANSAC (mask + H contract)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

cv2 = pytest.importorskip("cv2")


def _clip(res, n=3, kind="shake"):
    from tools.synth import Clip
    return Clip(res, kind, frames=n)


# ---------------------------------------------------------------------------------------------------------------------
# K0


@pytest.mark.parametrize("res,det", [("1080p", (480, 270)), ("4k", (480, 270)), ("720p", (480, 270)),
                                     ("1080p", (256, 256)), ("720p", (256, 256)), ("4k", (256, 256)),
                                     ((960, 540), (480, 270)), ((1440, 810), (480, 270)), ((480, 270), (480, 270))])
def test_detection_image_bit_exact(gpu_stream, oracle, res, det):
    frame = _clip(res, 1)[0]
    for fmt in (oracle.BGR, oracle.RGB, oracle.YUV):
        ref = oracle.detection_image(frame, fmt, det)
        got = gpu_stream.detection_image(frame, fmt, det)
        bad = int((ref != got).sum())
        print(f"{res}->{det} fmt={fmt}: mismatching pixels {bad}")
        assert bad == 0


def test_detection_image_odd_sizes(gpu_stream, oracle):
    rng = np.random.default_rng(3)
    for (w, h, dw, dh) in [(963, 541, 480, 270), (1366, 768, 256, 256), (1000, 600, 333, 200)]:
        frame = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        ref = oracle.detection_image(frame, oracle.BGR, (dw, dh))
        got = gpu_stream.detection_image(frame, oracle.BGR, (dw, dh))
        assert (ref == got).all(), f"{w}x{h}->{dw}x{dh}: {(ref != got).sum()} mismatches"


def test_detection_image_device_memory(gpu_stream, oracle):
    torch = pytest.importorskip("torch")
    frame = _clip("1080p", 1)[0]
    d = torch.from_numpy(frame).cuda()
    torch.cuda.synchronize()
    got = gpu_stream.detection_image(d, oracle.BGR, (480, 270))
    assert (got == oracle.detection_image(frame, oracle.BGR, (480, 270))).all()


# ---------------------------------------------------------------------------------------------------------------------
# K2


def _fast_ref(img, roi, thr):
    x, y, w, h = roi
    det = cv2.FastFeatureDetector_create(int(thr), True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    kps = det.detect(img[y:y + h, x:x + w], None)
    return np.array([(k.pt[0], k.pt[1], k.response) for k in kps], dtype=np.float32).reshape(-1, 3)


@pytest.mark.parametrize("thr", [10, 25, 60])
def test_fast_bit_exact(gpu_stream, oracle, thr):
    frame = _clip("1080p", 1)[0]
    det = oracle.detection_image(frame, oracle.BGR, (480, 270))
    for roi in [(0, 0, 240, 270), (240, 0, 240, 270), (0, 0, 480, 270), (17, 9, 101, 77)]:
        ref = _fast_ref(det, roi, thr)
        got = gpu_stream.fast_detect(det, roi, thr)
        g = np.stack([got["x"], got["y"], got["response"]], axis=1) if len(got) else np.zeros((0, 3), np.float32)
        print(f"thr={thr} roi={roi}: cv2 {len(ref)} gpu {len(g)}")
        assert len(ref) == len(g)
        assert (ref == g).all()  # coordinates, emission order and response


def test_fast_random_noise(gpu_stream):
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (131, 203), dtype=np.uint8)
    ref = _fast_ref(img, (0, 0, 203, 131), 20)
    got = gpu_stream.fast_detect(img, (0, 0, 203, 131), 20)
    g = np.stack([got["x"], got["y"], got["response"]], axis=1)
    assert len(ref) == len(g) and (ref == g).all()


def test_fast_flat_image_is_empty(gpu_stream):
    img = np.full((64, 64), 128, dtype=np.uint8)
    assert len(gpu_stream.fast_detect(img, (0, 0, 64, 64), 10)) == 0


# ---------------------------------------------------------------------------------------------------------------------
# K1 + K4


def _lk_ref(prev, nxt, pts):
    lk = cv2.SparsePyrLKOpticalFlow_create(winSize=(11, 11), maxLevel=3,
                                           crit=(cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 5, 0.01))
    out, status, _ = lk.calc(prev, nxt, pts.reshape(-1, 1, 2).astype(np.float32), None)
    return out.reshape(-1, 2), status.reshape(-1)


def _final_oob(points, w, h):
    """cv2-python always requests the error output, which adds one bounds test on the FINAL position
    (lkpyramid.cpp, err branch); the reference's C++ call passes no err (FrameTracker.cpp:140-146), and so do we."""
    p = points - 5.0
    ix, iy = np.floor(p[:, 0]), np.floor(p[:, 1])
    return (ix < -11) | (ix >= w) | (iy < -11) | (iy >= h)


@pytest.mark.parametrize("det", [(480, 270), (256, 256)])
def test_lk_parity(gpu_stream, oracle, det):
    clip = _clip("1080p", 4)
    a = oracle.detection_image(clip[0], oracle.BGR, det)
    b = oracle.detection_image(clip[2], oracle.BGR, det)
    kps = cv2.FastFeatureDetector_create(15, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16).detect(a, None)
    pts = np.array([k.pt for k in kps], dtype=np.float32)
    rng = np.random.default_rng(1)
    # add border / flat-area probes and sub-pixel positions
    extra = np.array([[0.0, 0.0], [det[0] - 1.0, det[1] - 1.0], [2.5, 100.25], [det[0] - 2.0, 3.0], [-3.0, 5.0]],
                     dtype=np.float32)
    pts = np.concatenate([pts + rng.uniform(-0.5, 0.5, pts.shape).astype(np.float32), extra])
    ref, rstat = _lk_ref(a, b, pts)
    got, gstat = gpu_stream.lk_track(a, b, pts)
    explained = (rstat == 0) & (gstat == 1) & _final_oob(got, det[0], det[1])
    mism = (rstat != gstat) & ~explained
    both = (rstat == 1) & (gstat == 1)
    err = np.abs(ref[both] - got[both]).max() if both.any() else 0.0
    print(f"det={det}: {len(pts)} pts, status mismatches {int(mism.sum())} (+{int(explained.sum())} explained by the "
          f"python-only err check), tracked {int(both.sum())}, max |dpos| = {err:.2e} px, "
          f"bit-identical positions {float((ref[both] == got[both]).all(axis=1).mean()):.4f}")
    assert mism.sum() == 0
    assert err <= 0.01


@pytest.mark.parametrize("size", [(480, 270), (256, 256), (333, 201), (97, 64), (50, 37)])
def test_pyramid_builders_agree(gpu_stream, size, monkeypatch):
    """The tile pyramid builder (default, no grid-wide barriers) against the per-level builder: a dense grid of
    probes, borders included, must track bit-identically, i.e. every level and derivative plane the windows touch
    (all of them, with an 11-px window every 3 px) is the same.  The per-level builder is the one pinned against
    cv2 by test_lk_parity before the tile builder existed."""
    w, h = size
    rng = np.random.default_rng(w * 1000 + h)
    base = cv2.GaussianBlur(rng.integers(0, 256, (h + 8, w + 8), dtype=np.uint8), (0, 0), 1.5)
    a = np.ascontiguousarray(base[4:4 + h, 4:4 + w])
    b = np.ascontiguousarray(base[3:3 + h, 2:2 + w])  # shifted by (2, 1)
    xs, ys = np.meshgrid(np.arange(-4, w + 4, 3, dtype=np.float32), np.arange(-4, h + 4, 3, dtype=np.float32))
    pts = np.stack([xs.ravel() + 0.37, ys.ravel() + 0.61], axis=1).astype(np.float32)
    results = {}
    for builder in ("tiles", "levels", "coop"):
        monkeypatch.setenv("LVKB200_PYRAMID", builder)
        results[builder] = gpu_stream.lk_track(a, b, pts)
    monkeypatch.delenv("LVKB200_PYRAMID")
    for other in ("levels", "coop"):
        assert np.array_equal(results["tiles"][1], results[other][1]), other
        assert np.array_equal(results["tiles"][0], results[other][0]), other
    assert results["tiles"][1].sum() > 0.3 * len(pts)


def test_lk_repeated_calls_square_epsilon(gpu_stream, oracle):
    """Upstream OpenCV squares TermCriteria::epsilon IN PLACE on every calc() of one SparsePyrLKOpticalFlow object;
    the reference reuses one m_OpticalTracker for all frames (FrameTracker.cpp:41-48), so its n-th frame iterates
    until |delta|^2 <= 0.01^(2^n).  The C-ABI reproduces that through call_index."""
    det = (480, 270)
    clip = _clip("1080p", 3)
    a = oracle.detection_image(clip[0], oracle.BGR, det)
    b = oracle.detection_image(clip[1], oracle.BGR, det)
    kps = cv2.FastFeatureDetector_create(20, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16).detect(a, None)
    rng = np.random.default_rng(3)
    pts = (np.array([k.pt for k in kps], dtype=np.float32) + rng.uniform(-0.5, 0.5, (len(kps), 2))).astype(np.float32)
    lk = cv2.SparsePyrLKOpticalFlow_create(winSize=(11, 11), maxLevel=3,
                                           crit=(cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 5, 0.01))
    first = None
    for call in range(11):
        out, status, _ = lk.calc(a, b, pts.reshape(-1, 1, 2), None)
        ref, rstat = out.reshape(-1, 2), status.reshape(-1)
        if call == 0:
            first = ref.copy()
        if call in (0, 1, 2, 3, 10):
            got, gstat = gpu_stream.lk_track(a, b, pts, call_index=call)
            explained = (rstat == 0) & (gstat == 1) & _final_oob(got, det[0], det[1])
            assert ((rstat != gstat) & ~explained).sum() == 0
            both = (rstat == 1) & (gstat == 1)
            err = float(np.abs(ref[both] - got[both]).max())
            moved = float(np.abs(ref[both] - first[both]).max())
            print(f"call {call}: max |gpu - cv2| = {err:.2e} px; cv2 drift vs its first call = {moved:.2e} px")
            assert err <= 1e-3
    assert moved > 1e-3  # the quirk is real: later calls iterate further than the first one


def test_lk_large_motion_and_flat(gpu_stream):
    rng = np.random.default_rng(5)
    a = cv2.GaussianBlur(rng.integers(0, 256, (270, 480)).astype(np.uint8), (0, 0), 2.0)
    M = np.float32([[1, 0, 6.3], [0, 1, -4.1]])
    b = cv2.warpAffine(a, M, (480, 270), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT)
    a[100:160, 200:300] = 77  # flat patch -> minEig rejection
    b[100:160, 200:300] = 77
    ys, xs = np.mgrid[8:262:12, 8:472:12]
    pts = np.stack([xs.ravel(), ys.ravel()], axis=1).astype(np.float32)
    ref, rstat = _lk_ref(a, b, pts)
    got, gstat = gpu_stream.lk_track(a, b, pts)
    explained = (rstat == 0) & (gstat == 1) & _final_oob(got, 480, 270)
    assert ((rstat != gstat) & ~explained).sum() == 0
    both = (rstat == 1) & (gstat == 1)
    assert both.sum() > 100 and (gstat == 0).sum() > 5
    assert np.abs(ref[both] - got[both]).max() <= 0.01


# ---------------------------------------------------------------------------------------------------------------------
# K6a


def _corner_disp(H1, H2, w, h):
    c = np.array([[0, 0, 1], [w, 0, 1], [0, h, 1], [w, h, 1]], dtype=np.float64).T
    a = H1 @ c
    b = H2 @ c
    return float(np.abs(a[:2] / a[2] - b[:2] / b[2]).max())


def _err2_f32(H, p, q):
    m = H.astype(np.float32).reshape(9)
    x, y = p[:, 0], p[:, 1]
    z = np.float32(1.0) / (m[6] * x + m[7] * y + m[8])
    dx = q[:, 0] - (m[0] * x + m[1] * y + m[2]) * z
    dy = q[:, 1] - (m[3] * x + m[4] * y + m[5]) * z
    return dx * dx + dy * dy


@pytest.mark.parametrize("noise,outliers", [(0.0, 0.0), (0.05, 0.0), (0.1, 0.15), (0.3, 0.3)])
def test_homography_contract(gpu_stream, oracle, noise, outliers):
    rng = np.random.default_rng(int(noise * 100) + int(outliers * 1000))
    n, w, h, thr = 1200, 480, 270, 3.0
    p = np.stack([rng.uniform(5, w - 5, n), rng.uniform(5, h - 5, n)], axis=1).astype(np.float32)
    Ht = np.array([[1.002, -0.006, 2.4], [0.0055, 0.999, -1.7], [2e-6, -1e-6, 1.0]])
    ph = np.concatenate([p, np.ones((n, 1), np.float32)], axis=1).astype(np.float64) @ Ht.T
    q = (ph[:, :2] / ph[:, 2:]).astype(np.float32) + rng.normal(0, noise, (n, 2)).astype(np.float32)
    n_out = int(outliers * n)
    q[:n_out] += rng.uniform(-40, 40, (n_out, 2)).astype(np.float32)
    Hc, mc = cv2.findHomography(p.reshape(-1, 1, 2), q.reshape(-1, 1, 2), oracle.usac_params(thr))
    Hg, mg = gpu_stream.find_homography(p, q, thr)
    mc = mc.reshape(-1).astype(np.uint8)
    # (1) our mask is exactly the thresholded float32 forward error of OUR returned H
    assert ((_err2_f32(Hg, p, q) < np.float32(thr * thr)).astype(np.uint8) == mg).all()
    # (2) masks agree with cv2 except points within epsilon of the threshold under either model
    e_c, e_g = np.sqrt(_err2_f32(Hc, p, q)), np.sqrt(_err2_f32(Hg, p, q))
    borderline = (np.abs(e_c - thr) < 0.25) | (np.abs(e_g - thr) < 0.25)
    hard = (mc != mg) & ~borderline
    disp = _corner_disp(Hc, Hg, w, h)
    print(f"noise={noise} outliers={outliers}: inliers cv2 {int(mc.sum())} gpu {int(mg.sum())}, mask mismatches "
          f"{int((mc != mg).sum())} ({int(hard.sum())} not borderline), corner displacement vs cv2 {disp:.4f} px, "
          f"vs truth: cv2 {_corner_disp(Hc, Ht, w, h):.4f} gpu {_corner_disp(Hg, Ht, w, h):.4f}")
    assert hard.sum() == 0
    # (3) H within the estimator's own input-order variance (<= ~0.2 px on noisy data)
    assert disp <= (1e-3 if noise == 0.0 else 0.25)
    assert abs(Hg[2, 2] - 1.0) < 1e-12


def test_homography_degenerate_is_no_model(gpu_stream):
    import livevisionkit_b200 as L
    t = np.linspace(0, 400, 100, dtype=np.float32)
    p = np.stack([t, 0.5 * t + 3], axis=1).astype(np.float32)  # collinear (cv2 returns None: SURVEY App. B12)
    with pytest.raises(L.LvkB200Error) as e:
        gpu_stream.find_homography(p, p + 1.0, 3.0)
    assert e.value.status == 4  # LVKB200_ERR_NO_MODEL


def test_homography_pure_translation(gpu_stream):
    rng = np.random.default_rng(9)
    p = np.stack([rng.uniform(0, 480, 75), rng.uniform(0, 270, 75)], axis=1).astype(np.float32)
    q = p + np.float32([3.0, -2.0])
    H, m = gpu_stream.find_homography(p, q, 3.0)
    assert m.all()
    assert np.allclose(H, [[1, 0, 3], [0, 1, -2], [0, 0, 1]], atol=1e-4)


@pytest.mark.parametrize("noise,outliers", [(0.0, 0.0), (0.1, 0.2), (0.3, 0.35)])
def test_affine_partial_contract(gpu_stream, noise, outliers):
    """cv::estimateAffinePartial2D(RANSAC, thr, 50) branch (FrameTracker.cpp:362-373)."""
    rng = np.random.default_rng(int(noise * 100) + int(outliers * 1000) + 7)
    n, thr = 900, 3.0
    p = np.stack([rng.uniform(5, 475, n), rng.uniform(5, 265, n)], axis=1).astype(np.float32)
    ang, sc = np.deg2rad(0.6), 1.004
    A = np.array([[sc * np.cos(ang), -sc * np.sin(ang), 2.2], [sc * np.sin(ang), sc * np.cos(ang), -1.4]])
    q = (p @ A[:, :2].T + A[:, 2]).astype(np.float32) + rng.normal(0, noise, (n, 2)).astype(np.float32)
    n_out = int(outliers * n)
    q[:n_out] += rng.uniform(-50, 50, (n_out, 2)).astype(np.float32)
    Ac, mc = cv2.estimateAffinePartial2D(p.reshape(-1, 1, 2), q.reshape(-1, 1, 2), None, cv2.RANSAC, thr, 50)
    Hg, mg = gpu_stream.estimate_affine_partial(p, q, thr)
    mc = mc.reshape(-1).astype(np.uint8)
    # a similarity: [a -b tx; b a ty; 0 0 1]
    assert abs(Hg[0, 0] - Hg[1, 1]) < 1e-12 and abs(Hg[0, 1] + Hg[1, 0]) < 1e-12 and (Hg[2] == [0, 0, 1]).all()
    Hc = np.eye(3)
    Hc[:2] = Ac
    e_c = np.linalg.norm(p @ Hc[:2, :2].T + Hc[:2, 2] - q, axis=1)
    e_g = np.linalg.norm(p @ Hg[:2, :2].T + Hg[:2, 2] - q, axis=1)
    borderline = (np.abs(e_c - thr) < 0.5) | (np.abs(e_g - thr) < 0.5)
    hard = (mc != mg) & ~borderline
    disp = _corner_disp(Hc, Hg, 480, 270)
    truth = np.eye(3)
    truth[:2] = A
    print(f"affine noise={noise} outliers={outliers}: inliers cv2 {int(mc.sum())} gpu {int(mg.sum())}, mismatches "
          f"{int((mc != mg).sum())} ({int(hard.sum())} not borderline), corner disp vs cv2 {disp:.4f} px; vs truth cv2 "
          f"{_corner_disp(Hc, truth, 480, 270):.4f} gpu {_corner_disp(Hg, truth, 480, 270):.4f}")
    assert hard.sum() == 0
    assert disp <= (1e-3 if noise == 0.0 else 0.25)


def test_pipeline_takes_the_affine_branch_on_clustered_features(oracle):
    """Features confined to one corner -> distribution <= 0.6 -> estimateAffinePartial2D branch, vs the oracle."""
    import livevisionkit_b200 as L
    from livevisionkit_b200 import _capi as K
    from tools.synth import Clip
    clip = Clip((960, 540), "shake", frames=6, seed=3)
    so = oracle.StabilizationSettings.obs_homography_preset()
    so.uniformity_threshold = 0.0
    sg = L.StabilizationFilterSettings.obs_homography_preset()
    sg.uniformity_threshold = 0.0
    ref, flt = oracle.StabilizationFilter(so), L.StabilizationFilter(sg, 0)
    flt.stream.set_debug_capture(True)
    took_affine = False
    for i in range(6):
        f = clip[i].copy()
        f[:, 400:] = 90   # flatten everything except the left part of the frame: corners only on the left
        f[300:, :] = 90
        ref.apply(f, oracle.BGR, i)
        flt.apply(L.VideoFrame(f, i, L.BGR))
        tr = ref.trace
        if "H" in tr and tr["distribution"] <= np.float32(0.6):
            took_affine = True
            Hg = flt.stream.debug_fetch(K.DBG_HOMOGRAPHY, np.float64).reshape(3, 3)
            assert (Hg[2] == [0, 0, 1]).all() and abs(Hg[0, 0] - Hg[1, 1]) < 1e-12  # a similarity, like the oracle's
            assert (tr["H"][2] == [0, 0, 1]).all()
            assert _corner_disp(Hg, tr["H"], 200, 150) <= 0.25
            inl = flt.stream.debug_fetch(K.DBG_INLIERS, np.uint8)
            assert len(inl) == len(tr["inliers"]) and (inl != tr["inliers"]).mean() < 0.02
    assert took_affine, "the clip never produced a badly distributed feature set"


def test_local_motions_vs_oracle(gpu_stream, oracle, monkeypatch):
    """The HOST restatement of the LSCG solve (host_mesh.hpp; the fallback for meshes too large for one CTA, selected
    here with the debug knob) against the oracle: same summation order as Eigen -> <= 1e-3 px."""
    import livevisionkit_b200 as L
    monkeypatch.setenv("LVKB200_MESH_DEVICE_MIN", "1000000")
    s = L.Stream(L.StabilizationFilterSettings(), 0)  # defaults: 256x256, 2x2 mesh, local motions
    rng = np.random.default_rng(4)
    n = 900
    p = np.stack([rng.uniform(0, 256, n), rng.uniform(0, 256, n)], axis=1).astype(np.float32)
    q = (p * np.float32(1.003) + np.float32([1.2, -0.8]) + rng.normal(0, 0.05, (n, 2))).astype(np.float32)
    q[:40] += 25.0
    trk = oracle.FrameTracker(oracle.StabilizationSettings())
    state = np.zeros(8, dtype=np.float32)
    for it in range(3):  # warm-started over consecutive "frames"
        motion_ref, inl_ref = trk.estimate_local_motions(p.tolist(), q.tolist())
        state, offsets, mask = s.estimate_local_motions(p, q, state)
        assert (mask == inl_ref).all()
        assert np.abs(offsets.reshape(2, 2, 2) - motion_ref).max() * 256 <= 1e-3  # corner displacement in px
        assert np.abs(state - trk.optimized_mesh).max() <= 1e-3
    s.close()


@pytest.mark.parametrize("mesh", ["field16x16", "field16x16_round1_kernel", "field16x16_one_slot", "field24x24_two_slots",
                                  "default2x2_on_device"])
def test_local_motions_device_solver_vs_oracle(gpu_stream, oracle, mesh, monkeypatch):
    """K6c k_mesh_cgls2 / k_mesh_cgls (one CTA, whole LSCG solve in shared memory) against the sequential CPU restatement
    of Eigen's LeastSquaresConjugateGradient (oracle/lscg_ref.c): same iteration, different float32 summation order.
    The 16x16 mesh runs on the re-tiled 1024-thread kernel by default and on the round-1 kernel (which remains the
    solver of the 2x2 mesh and of meshes beyond the re-tiled kernel's slot limits) behind LVKB200_MESH_V1=1."""
    import livevisionkit_b200 as L
    if mesh == "field16x16_round1_kernel":
        monkeypatch.setenv("LVKB200_MESH_V1", "1")
    n = 1100
    if mesh.startswith("field"):
        sg, so = L.StabilizationFilterSettings.obs_field_preset(), oracle.StabilizationSettings.obs_field_preset()
        # k_mesh_cgls2 is instantiated per (unknowns, similarity rows, features) a thread may own: the preset's capacity
        # (1 856 features) takes <1,1,3>; a sparser detection grid (capacity 704) <1,1,1>; 1 152 unknowns <2,2,4>
        if mesh == "field16x16_one_slot":
            n = 700
            for cfg in (sg, so):
                cfg.max_feature_density, cfg.min_feature_density = 0.06, 0.03
        elif mesh == "field24x24_two_slots":
            for cfg in (sg, so):
                cfg.motion_resolution = (24, 24)
    else:  # the library-default 2x2 mesh: on the device by default
        sg, so = L.StabilizationFilterSettings(), oracle.StabilizationSettings()
    w, h = so.detection_resolution
    mc, mr = so.motion_resolution
    s = L.Stream(sg, 0)
    rng = np.random.default_rng(11)
    p = np.stack([rng.uniform(0, w, n), rng.uniform(0, h, n)], axis=1).astype(np.float32)
    # a smooth non-rigid field (what the mesh is for) + noise + gross outliers
    flow = np.stack([1.5 + 2.0 * np.sin(p[:, 1] / h * 3.0), -0.8 + 1.5 * np.cos(p[:, 0] / w * 2.0)], axis=1)
    q = (p * np.float32(1.002) + flow + rng.normal(0, 0.05, (n, 2))).astype(np.float32)
    q[:50] += 30.0
    trk = oracle.FrameTracker(so)
    state = np.zeros(2 * mc * mr, dtype=np.float32)
    worst_state = worst_off = 0.0
    for it in range(4):  # warm-started over consecutive "frames"
        motion_ref, inl_ref = trk.estimate_local_motions(p.tolist(), q.tolist())
        state, offsets, mask = s.estimate_local_motions(p, q, state)
        assert (mask != inl_ref).sum() <= 1, f"frame {it}: {(mask != inl_ref).sum()} inlier flags differ"
        worst_off = max(worst_off, float(np.abs(offsets.reshape(mr, mc, 2) - motion_ref).max() * max(w, h)))
        worst_state = max(worst_state, float(np.abs(state - np.asarray(trk.optimized_mesh).ravel()).max()))
        q = (q + rng.normal(0, 0.02, (n, 2))).astype(np.float32)
    print(f"[mesh {mesh}] max |vertex - oracle| {worst_state:.2e} px, max offset deviation {worst_off:.2e} px")
    assert worst_state <= 5e-3 and worst_off <= 5e-3
    s.close()
