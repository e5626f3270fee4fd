"""DeblockingFilter oracle pinning (CPU): the restated OpenCV arithmetic (what the CUDA kernels implement) against the
cv2 calls the reference makes (Filters/DeblockingFilter.cpp:48-118), stage by stage, plus the committed golden vector
and first-principles properties.  The reference holds no vectors for this filter and cannot be built here."""
import os

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def D():
    from oracle import deblock_oracle
    return deblock_oracle


def test_restated_primitives_match_cv2(D):
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (64, 96, 3), dtype=np.uint8)
    assert (D.area_integer(img, 4, 4) == cv2.resize(img, None, fx=0.25, fy=0.25, interpolation=cv2.INTER_AREA)).all()
    assert (D.area_integer(img[..., 0], 16, 16) == cv2.resize(img[..., 0], (6, 4), interpolation=cv2.INTER_AREA)).all()
    for k in (3, 5, 7):
        assert (D.median_8u(img, k) == cv2.medianBlur(img, k)).all()
    for (dw, dh) in ((384, 256), (960, 512), (200, 131)):
        assert (D.linear_8u(img, dw, dh) == cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR)).all()
    f = rng.choice(np.array([0, 1 / 3., 2 / 3., 1], dtype=np.float32), (9, 13)).astype(np.float32)
    use = cv2.ipp.useIPP()
    try:
        cv2.ipp.setUseIPP(False)  # OpenCV's own kernels: bit-exact
        assert (D.linear_32f(f, 208, 144) == cv2.resize(f, (208, 144), interpolation=cv2.INTER_LINEAR)).all()
    finally:
        cv2.ipp.setUseIPP(use)
    # the IPP build of the same call differs by at most one ulp
    ipp = cv2.resize(f, (208, 144), interpolation=cv2.INTER_LINEAR)
    assert np.abs(D.linear_32f(f, 208, 144) - ipp).max() <= 1.2e-7
    a, b = img, rng.integers(0, 256, img.shape, dtype=np.uint8)
    w1 = rng.random(img.shape[:2], dtype=np.float32)
    w2 = np.abs(w1 - np.float32(1)).astype(np.float32)
    assert (D.blend_linear(a, b, w1, w2) == cv2.blendLinear(a, b, w1, w2)).all()


@pytest.mark.parametrize("size,levels,fmt", [((640, 360), 3, 0), ((1280, 720), 3, 4), ((963, 541), 2, 0),
                                             ((640, 360), 6, 2), ((160, 96), 1, 0)])
def test_restated_filter_matches_reference_calls(D, oracle, size, levels, fmt):
    from tools.synth import Clip
    frame = D.blocky_frame(Clip(size, "shake", frames=2)[1], 16, 0.8, seed=levels)
    s = D.DeblockingSettings(detection_levels=levels)
    flt = D.DeblockingFilter(s)
    ref = flt.apply(frame, fmt)
    stages = {}
    got = D.deblock_restated(frame, fmt, s, stages)
    diff = np.abs(got.astype(int) - ref.astype(int))
    changed = float((ref != frame).mean())
    print(f"{size} levels={levels}: {int((diff > 0).sum())} bytes differ (max {int(diff.max())}), filter changed "
          f"{100 * changed:.1f}% of the bytes, keep levels {np.unique(stages['fbuf'])}")
    assert diff.max() <= 1 and (diff > 0).sum() <= 1e-5 * diff.size  # IPP's float resize: <= 1 ulp in the weights
    assert changed > 0.2
    assert len(np.unique(stages["fbuf"])) == levels + 1
    # outside the whole-macroblock region the frame passes through (DeblockingFilter.cpp:67-73)
    x, y, rw, rh = flt.filter_region
    assert (ref[rh:] == frame[rh:]).all() and (ref[:, rw:] == frame[:, rw:]).all()


def test_properties(D):
    rng = np.random.default_rng(2)
    # a frame whose macroblocks are all textured (deviation >= levels) is returned unchanged
    noisy = rng.integers(0, 256, (96, 160, 3), dtype=np.uint8)
    assert (D.DeblockingFilter().apply(noisy) == noisy).all()
    assert (D.deblock_restated(noisy) == noisy).all()
    # a frame of flat macroblocks is fully replaced by the smooth frame: block edges get blurred, block centres stay
    blocks = np.repeat(np.repeat(rng.integers(0, 256, (6, 10, 1), dtype=np.uint8), 16, 0), 16, 1).repeat(3, 2)
    out = D.deblock_restated(blocks)
    assert (out != blocks).any()
    assert np.abs(out.astype(int) - blocks.astype(int)).max() <= 255
    # settings preconditions (DeblockingFilter.cpp:38-42)
    for bad in (dict(block_size=0), dict(filter_size=4), dict(filter_size=1), dict(detection_levels=0), dict(filter_scaling=1.0)):
        with pytest.raises(AssertionError):
            D.DeblockingFilter(D.DeblockingSettings(**bad))


def test_golden(D):
    g = np.load(os.path.join(G, "deblock_golden.npz"), allow_pickle=False)
    for fmt, key in ((0, "out_bgr"), (4, "out_yuv")):
        assert (D.deblock_restated(g["frame"], fmt) == g[key]).all()
        ref = D.DeblockingFilter().apply(g["frame"], fmt)
        assert np.abs(ref.astype(int) - g[key].astype(int)).max() <= 1
