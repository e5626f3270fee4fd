"""FrameIngest oracle pinning (no GPU): the cv2 form of oracle/ingest_oracle.py (the reference's own OpenCV calls) against
the written-out integer arithmetic that livevisionkit_b200/csrc/formats.cu implements, plus first-principles properties."""
import cv2
import numpy as np
import pytest

from oracle import ingest_oracle as I

SIZES = [(64, 36), (1280, 720), (482, 270), (1920, 1080)]
YUV_FORMATS = [f for f in I.FORMATS if f not in ("Y800", "BGR3")]


@pytest.mark.parametrize("fmt", YUV_FORMATS)
@pytest.mark.parametrize("size", SIZES[:3])
def test_upload_restatement_matches_cv2(fmt, size):
    w, h = size
    frame = I.random_frame(fmt, w, h, seed=w + h)
    ref = I.upload_obs_frame(frame)
    got = I.upload_restated(frame)
    assert ref.shape == (h, w, 3) and ref.dtype == np.uint8
    assert (ref == got).all(), f"{fmt} {size}: {(ref != got).sum()} bytes differ"


@pytest.mark.parametrize("fmt", YUV_FORMATS)
@pytest.mark.parametrize("size", SIZES[:3])
def test_download_restatement_matches_cv2(fmt, size):
    w, h = size
    rng = np.random.default_rng(w * h)
    img = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
    ref = I.download_ocl_frame(img, fmt)
    got = I.download_restated(img, fmt)
    assert len(ref) == len(got)
    for a, b, (rows, rb) in zip(ref, got, I.plane_shapes(fmt, w, h)):
        assert a.reshape(rows, rb).shape == (rows, rb)
        assert (a.reshape(rows, rb) == b.reshape(rows, rb)).all(), f"{fmt} {size}"


def test_restatement_is_independent_of_ipp():
    """cv2's IPP build and its plain C++ path agree on these two resizes; the restatement matches both."""
    rng = np.random.default_rng(5)
    src = rng.integers(0, 256, size=(270, 480, 2), dtype=np.uint8)
    try:
        for use in (False, True):
            cv2.ipp.setUseIPP(use)
            assert (cv2.resize(src, (960, 540), interpolation=cv2.INTER_LINEAR) == I.resize_linear_u8(src, 960, 540)).all()
            assert (cv2.resize(src, None, fx=0.5, fy=0.5, interpolation=cv2.INTER_AREA) == I.area_half(src, 2, 2)).all()
            assert (cv2.resize(src, None, fx=0.5, fy=1.0, interpolation=cv2.INTER_AREA) == I.area_half(src, 2, 1)).all()
    finally:
        cv2.ipp.setUseIPP(True)


@pytest.mark.parametrize("fmt", ["I444", "YUVA", "AYUV", "BGR3", "Y800"])
def test_lossless_layouts_round_trip(fmt):
    w, h = 96, 54
    frame = I.random_frame(fmt, w, h, seed=9)
    img = I.upload_obs_frame(frame)
    back = I.download_ocl_frame(img, fmt)
    for a, b in zip(frame["planes"], back):
        a = a.copy()
        if fmt == "AYUV":
            a.reshape(h, w, 4)[:, :, 0] = 255  # the reference rewrites alpha as 255 (FrameIngest.cpp:709)
        assert (a.reshape(b.shape) == b).all()


@pytest.mark.parametrize("fmt", ["I420", "NV12", "YUY2", "I422"])
def test_subsampled_layouts_preserve_luma_and_flat_chroma(fmt):
    w, h = 128, 72
    frame = I.random_frame(fmt, w, h, seed=3)
    kind = I.FORMATS[fmt][0]
    if kind in ("planar", "semiplanar"):
        for p in frame["planes"][1:]:
            p[:] = 77  # flat chroma survives up- and down-sampling exactly
    img = I.upload_obs_frame(frame)
    back = I.download_ocl_frame(img, fmt)
    if kind in ("planar", "semiplanar"):
        assert (img[:, :, 1] == 77).all() and (img[:, :, 2] == 77).all()
        for a, b in zip(frame["planes"], back):
            assert (a.reshape(b.shape) == b).all()
    else:
        yo = I.FORMATS[fmt][3][0]
        assert (back[0].reshape(h, w, 2)[:, :, yo] == frame["planes"][0].reshape(h, w, 2)[:, :, yo]).all()
