"""
ORACLE — TEST INFRASTRUCTURE ONLY (same rules as lvk_oracle.py: tests/, smoke() and bench.py's CPU legs only).

CPU restatement of lvk::DeblockingFilter (reference commit 2f7bb70), the SURVEY 8(f)-1 "next" row:
  * DeblockingFilter::configure / filter      Filters/DeblockingFilter.cpp:36-118
  * DeblockingFilterSettings defaults         Filters/DeblockingFilter.hpp:26-32

Two forms:
  * `DeblockingFilter.apply`   the reference's body statement by statement through the cv2 wheel in this image
                               (cv2 4.13.0 with IPP; the reference pins 4.8.0) — the checker;
  * `deblock_restated`         OpenCV's own (non-IPP) CPU arithmetic written out in NumPy integer / float32 steps —
                               what the CUDA kernels implement.  It is pinned against cv2 by tests/test_oracle_cpu.py:
                               every 8-bit stage is bit-exact; the float32 blend maps differ from the IPP build by
                               <= 1 ulp on ~20 % of the pixels (cv2.ipp.setUseIPP(False) reproduces them exactly),
                               which never reached an output byte in the tested frames.

PARITY UNPINNED: the reference ships no tests or golden vectors for this filter and cannot be compiled here.
"""
from __future__ import annotations

from dataclasses import dataclass

import cv2
import numpy as np

from .lvk_oracle import BGR, to_gray


@dataclass
class DeblockingSettings:  # Filters/DeblockingFilter.hpp:26-32
    detection_levels: int = 3
    block_size: int = 16
    filter_size: int = 5
    filter_scaling: float = 4.0


class DeblockingFilter:
    def __init__(self, settings: DeblockingSettings | None = None):
        self.configure(settings or DeblockingSettings())
        self.filter_region = (0, 0, 0, 0)

    def configure(self, s: DeblockingSettings):  # DeblockingFilter.cpp:36-45
        assert s.block_size > 0
        assert s.filter_size >= 3 and s.filter_size % 2 == 1
        assert s.detection_levels > 0
        assert s.filter_scaling > 1.0
        self.s = s

    def apply(self, frame: np.ndarray, fmt: int = BGR) -> np.ndarray:  # filter(), :49-118
        assert frame.size > 0
        s = self.s
        bs = int(s.block_size)
        h, w = frame.shape[:2]
        ex, ey = w // bs, h // bs  # macroblock_extent (cv::Size / int)
        rw, rh = ex * bs, ey * bs
        self.filter_region = (0, 0, rw, rh)
        out = frame.copy()
        if rw == 0 or rh == 0:
            return out
        roi = out[:rh, :rw]

        # smooth frame (:76-79)
        area_scaling = np.float32(1.0) / np.float32(s.filter_scaling)
        small = cv2.resize(roi, None, fx=float(area_scaling), fy=float(area_scaling), interpolation=cv2.INTER_AREA)
        small = cv2.medianBlur(small, int(s.filter_size))
        smooth = cv2.resize(small, (rw, rh), interpolation=cv2.INTER_LINEAR)

        # reference frame (:82-86)
        det = to_gray(roi, fmt)
        grid = cv2.resize(det, (ex, ey), interpolation=cv2.INTER_AREA)
        ref = cv2.resize(grid, (rw, rh), interpolation=cv2.INTER_NEAREST)
        det = cv2.absdiff(det, ref)
        grid = cv2.resize(det, (ex, ey), interpolation=cv2.INTER_AREA)

        # blend maps (:89-100)
        fbuf = np.zeros((ey, ex), np.float32)
        level_step = 1.0 / s.detection_levels
        for level in range(int(s.detection_levels)):
            _, mask = cv2.threshold(grid, level, 255, cv2.THRESH_BINARY)
            fbuf[mask != 0] = np.float32((level + 1.0) * level_step)
        keep = cv2.resize(fbuf, (rw, rh), interpolation=cv2.INTER_LINEAR)
        deblock = cv2.absdiff(keep, np.full_like(keep, 1.0)).astype(np.float32)  # absdiff(map, Scalar(1.0))
        self.keep_map = keep

        # adaptive blend, in place on the region (:103-109)
        out[:rh, :rw] = cv2.blendLinear(np.ascontiguousarray(roi), smooth, keep, deblock)
        return out


# ---------------------------------------------------------------------------------------------------------------------
# OpenCV's CPU arithmetic, restated (upstream imgproc/resize.cpp, median_blur, blend.cpp — not under /root/reference)

def area_integer(img: np.ndarray, sx: int, sy: int) -> np.ndarray:
    """INTER_AREA with integer factors: exact integer block sum, one float32 multiply by 1/area, round-half-even."""
    h, w = img.shape[:2]
    c = 1 if img.ndim == 2 else img.shape[2]
    v = img.reshape(h // sy, sy, w // sx, sx, c).astype(np.int64).sum(axis=(1, 3))
    if sx == 2 and sy == 2:
        out = ((v + 2) >> 2).astype(np.uint8)  # OpenCV's 2x2 special case rounds half up
    else:
        out = np.rint(v.astype(np.float32) * np.float32(1.0 / (sx * sy))).astype(np.uint8)
    return out[..., 0] if img.ndim == 2 else out


def linear_tables(ssize: int, dsize: int, clamp_fraction: bool):
    """resize()'s per-axis source index + weights: fx = (float)((d + 0.5) * scale - 0.5).  Horizontally the fraction
    is zeroed where the index clamps; vertically the two ROW indices are clipped instead and the weights stay."""
    scale = ssize / dsize
    i0 = np.zeros(dsize, np.int64)
    i1 = np.zeros(dsize, np.int64)
    f = np.zeros(dsize, np.float32)
    for d in range(dsize):
        v = np.float32((d + 0.5) * scale - 0.5)
        s = int(np.floor(v))
        fr = np.float32(v - np.float32(s))
        if clamp_fraction:
            if s < 0:
                s, fr = 0, np.float32(0)
            if s >= ssize - 1:
                s, fr = ssize - 1, np.float32(0)
        i0[d] = min(max(s, 0), ssize - 1)
        i1[d] = min(max(s + 1, 0), ssize - 1)
        f[d] = fr
    return i0, i1, f


def linear_8u(img: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """INTER_LINEAR on 8-bit data: 11-bit fixed-point weights, horizontal pass in int32 (x2048), vertical pass
    ((b0*(r0>>4))>>16 + (b1*(r1>>4))>>16 + 2) >> 2."""
    h, w = img.shape[:2]
    x0, x1, fx = linear_tables(w, dw, True)
    y0, y1, fy = linear_tables(h, dh, False)
    one = np.float32(1.0)
    ax0 = np.rint((one - fx) * np.float32(2048)).astype(np.int64)
    ax1 = np.rint(fx * np.float32(2048)).astype(np.int64)
    ay0 = np.rint((one - fy) * np.float32(2048)).astype(np.int64)
    ay1 = np.rint(fy * np.float32(2048)).astype(np.int64)
    s = img.astype(np.int64)
    hor = s[:, x0] * ax0[None, :, None] + s[:, x1] * ax1[None, :, None]
    out = (((ay0[:, None, None] * (hor[y0] >> 4)) >> 16) + ((ay1[:, None, None] * (hor[y1] >> 4)) >> 16) + 2) >> 2
    return out.astype(np.uint8)


def linear_32f(img: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """INTER_LINEAR on float32 data, OpenCV's own kernels: a*w0 + b*w1 with separate roundings, rows then columns."""
    h, w = img.shape[:2]
    x0, x1, fx = linear_tables(w, dw, True)
    y0, y1, fy = linear_tables(h, dh, False)
    one = np.float32(1.0)
    hor = ((img[:, x0] * (one - fx)[None, :]).astype(np.float32) + (img[:, x1] * fx[None, :]).astype(np.float32)).astype(np.float32)
    return ((hor[y0] * (one - fy)[:, None]).astype(np.float32) + (hor[y1] * fy[:, None]).astype(np.float32)).astype(np.float32)


def median_8u(img: np.ndarray, k: int) -> np.ndarray:
    """medianBlur: exact per-channel median of the k x k window, BORDER_REPLICATE."""
    r = k // 2
    p = np.pad(img, ((r, r), (r, r), (0, 0)), mode="edge")
    h, w = img.shape[:2]
    win = np.stack([p[j:j + h, i:i + w] for j in range(k) for i in range(k)], axis=0)
    return np.partition(win, (k * k) // 2, axis=0)[(k * k) // 2]


def blend_linear(a: np.ndarray, b: np.ndarray, w1: np.ndarray, w2: np.ndarray) -> np.ndarray:
    """cv::blendLinear: (a*w1 + b*w2) / (w1 + w2 + 1e-5f) in float32, round-half-even, saturate."""
    den = ((w1 + w2).astype(np.float32) + np.float32(1e-5)).astype(np.float32)[..., None]
    num = ((a.astype(np.float32) * w1[..., None]).astype(np.float32) + (b.astype(np.float32) * w2[..., None]).astype(np.float32))
    return np.clip(np.rint((num.astype(np.float32) / den).astype(np.float32)), 0, 255).astype(np.uint8)


def level_values(levels: int) -> np.ndarray:
    """value written for 'block deviation > l' is (float)((l + 1.0) * (1.0 / levels)); index = number of levels passed."""
    step = 1.0 / levels
    return np.array([0.0] + [np.float32((level + 1.0) * step) for level in range(levels)], dtype=np.float32)


def deblock_restated(frame: np.ndarray, fmt: int = BGR, s: DeblockingSettings | None = None, stages: dict | None = None):
    """The filter with every OpenCV call replaced by its restated arithmetic.  Supports what the CUDA path supports:
    integer filter_scaling that divides block_size."""
    s = s or DeblockingSettings()
    bs, sc, k = int(s.block_size), int(s.filter_scaling), int(s.filter_size)
    assert float(sc) == float(s.filter_scaling) and bs % sc == 0
    h, w = frame.shape[:2]
    ex, ey = w // bs, h // bs
    rw, rh = ex * bs, ey * bs
    out = frame.copy()
    if rw == 0 or rh == 0:
        return out
    roi = out[:rh, :rw]
    small = median_8u(area_integer(roi, sc, sc), k)
    smooth = linear_8u(small, rw, rh)
    det = to_gray(roi, fmt)
    grid = area_integer(det, bs, bs)
    dev = np.abs(det.astype(np.int16) - np.repeat(np.repeat(grid, bs, axis=0), bs, axis=1).astype(np.int16)).astype(np.uint8)
    grid2 = area_integer(dev, bs, bs)
    passed = np.minimum(grid2.astype(np.int64), int(s.detection_levels))  # thresholds 0 .. levels-1, each "> l"
    fbuf = level_values(int(s.detection_levels))[passed]
    keep = linear_32f(fbuf, rw, rh)
    deblock = np.abs(keep - np.float32(1.0)).astype(np.float32)
    out[:rh, :rw] = blend_linear(roi, smooth, keep, deblock)
    if stages is not None:
        stages.update(small=small, smooth=smooth, grid=grid, grid2=grid2, fbuf=fbuf, keep=keep)
    return out


def blocky_frame(frame: np.ndarray, block: int = 16, strength: float = 0.7, seed: int = 0) -> np.ndarray:
    """Synthetic compression artefacts: pulls each macroblock towards its own mean with a per-block random weight, so
    that flat blocks, textured blocks and every detection level in between occur."""
    rng = np.random.default_rng(seed)
    h, w = frame.shape[:2]
    ey, ex = h // block, w // block
    out = frame.astype(np.float32)
    roi = out[:ey * block, :ex * block].reshape(ey, block, ex, block, -1)
    mean = roi.mean(axis=(1, 3), keepdims=True)
    wgt = (rng.random((ey, 1, ex, 1, 1)) ** 0.5 * strength).astype(np.float32)
    kind = rng.random((ey, 1, ex, 1, 1))
    wgt = np.where(kind < 0.25, np.float32(1.0), wgt)  # a quarter of the blocks entirely flat (deviation 0)
    wgt = np.where((kind >= 0.25) & (kind < 0.5), (0.9 + 0.1 * rng.random(kind.shape)).astype(np.float32), wgt)  # nearly flat
    roi[...] = roi * (1 - wgt) + mean * wgt
    return np.clip(np.rint(out), 0, 255).astype(np.uint8)
