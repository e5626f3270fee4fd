"""
ORACLE — TEST INFRASTRUCTURE ONLY (same rules as lvk_oracle.py: tests/, smoke() and bench.py's CPU legs only).

CPU restatement of the OBS frame ingest / egress of LiveVisionKit (reference commit 2f7bb70), SURVEY 8(f)-3:
  * FrameIngest::Select                              Modules/OBS-Plugin/Interop/FrameIngest.cpp:38-76
  * I4XXIngest::to_ocl / to_obs                      :479-560   (cv::resize INTER_LINEAR up, INTER_AREA down)
  * NV12Ingest::to_ocl / to_obs                      :566-604
  * P422Ingest::to_ocl / to_obs  (YUY2, YVYU, UYVY)  :608-678
  * P444Ingest::to_ocl / to_obs  (AYUV)              :684-716
  * DirectIngest::to_ocl / to_obs (Y800, BGR3)       :722-757

Two forms:
  * `upload_obs_frame` / `download_ocl_frame`   the reference's statements through the cv2 wheel of this image (cv2 4.13.0;
                                               the reference pins 4.8.0 and runs the same calls on OpenCL UMats) — the checker;
  * `resize_linear_u8` / `area_half`            OpenCV's CPU arithmetic for the two cv::resize calls written out in NumPy
                                               integer steps — what formats.cu implements.  tests/test_formats_cpu.py pins the
                                               restatement against cv2 bit-exactly (with and without IPP).

Frames are dicts {"format": name, "width": w, "height": h, "planes": [np.uint8 arrays]} with tightly packed planes, the
only layout the reference handles (upload_planes copies width*height*channels contiguous bytes per plane).

PARITY UNPINNED: the reference ships no tests or golden vectors for this path and cannot be compiled here.
"""
from __future__ import annotations

import cv2
import numpy as np

# name -> (kind, chroma sub_x, sub_y, packed byte offsets (y, u, v))
FORMATS = {
    "I420": ("planar", 2, 2, None), "I422": ("planar", 2, 1, None), "I444": ("planar", 1, 1, None),
    "I40A": ("planar", 2, 2, None), "I42A": ("planar", 2, 1, None), "YUVA": ("planar", 1, 1, None),
    "NV12": ("semiplanar", 2, 2, None),
    "YUY2": ("packed422", 2, 1, (0, 1, 3)), "YVYU": ("packed422", 2, 1, (0, 3, 1)), "UYVY": ("packed422", 2, 1, (1, 0, 2)),
    "AYUV": ("packed444", 1, 1, (1, 2, 3)),
    "Y800": ("direct", 1, 1, None), "BGR3": ("direct", 1, 1, None),
}
# numbering of lvkb200_video_format (include/lvkb200.h)
FORMAT_IDS = {"I420": 0, "I422": 1, "I444": 2, "I40A": 3, "I42A": 4, "YUVA": 5, "NV12": 6, "YVYU": 7, "YUY2": 8,
              "UYVY": 9, "AYUV": 10, "Y800": 11, "BGR3": 12}


def plane_shapes(fmt: str, w: int, h: int):
    """(rows, row_bytes) of every plane LVK touches."""
    kind, sx, sy, _ = FORMATS[fmt]
    cw, ch = w // sx, h // sy
    if kind == "planar":
        return [(h, w), (ch, cw), (ch, cw)]
    if kind == "semiplanar":
        return [(h, w), (ch, 2 * cw)]
    if kind == "packed422":
        return [(h, 2 * w)]
    if kind == "packed444":
        return [(h, 4 * w)]
    return [(h, w * (1 if fmt == "Y800" else 3))]


def random_frame(fmt: str, w: int, h: int, seed: int = 0):
    rng = np.random.default_rng(seed)
    planes = []
    for rows, rb in plane_shapes(fmt, w, h):
        # smooth + noise so that interpolation has structure to chew on
        base = rng.integers(0, 256, size=(rows // 8 + 2, rb // 8 + 2)).astype(np.float32)
        up = cv2.resize(base, (rb, rows), interpolation=cv2.INTER_CUBIC)
        planes.append(np.clip(up + rng.integers(-12, 13, size=(rows, rb)), 0, 255).astype(np.uint8))
    return {"format": fmt, "width": w, "height": h, "planes": planes}


def upload_obs_frame(frame) -> np.ndarray:
    """FrameIngest::upload_obs_frame -> to_ocl: returns the packed 8UC3 frame (8UC1 for Y800)."""
    fmt, w, h, p = frame["format"], frame["width"], frame["height"], frame["planes"]
    kind, sx, sy, offs = FORMATS[fmt]
    if kind == "planar":  # FrameIngest.cpp:479-522
        y, u, v = p[0], p[1], p[2]
        if (sx, sy) != (1, 1):
            u = cv2.resize(u, (w, h), interpolation=cv2.INTER_LINEAR)
            v = cv2.resize(v, (w, h), interpolation=cv2.INTER_LINEAR)
        return cv2.merge([y, u, v])
    if kind == "semiplanar":  # :566-584
        uv = p[1].reshape(h // 2, w // 2, 2)
        uv = cv2.resize(uv, (w, h), interpolation=cv2.INTER_LINEAR)
        return np.dstack([p[0], uv[:, :, 0], uv[:, :, 1]])
    if kind == "packed422":  # :618-645
        plane = p[0].reshape(h, w, 2)
        y_first, u_first = fmt != "UYVY", fmt != "YVYU"
        chroma = plane[:, :, 1 if y_first else 0]                      # extractChannel
        uv = np.ascontiguousarray(chroma).reshape(h, w // 2, 2)        # reshape(2, rows)
        uv = cv2.resize(uv, (w, h), interpolation=cv2.INTER_LINEAR)
        luma = plane[:, :, 0 if y_first else 1]
        return np.dstack([luma, uv[:, :, 0], uv[:, :, 1]] if u_first else [luma, uv[:, :, 1], uv[:, :, 0]])
    if kind == "packed444":  # :689-698
        q = p[0].reshape(h, w, 4)
        return np.ascontiguousarray(q[:, :, 1:4])
    return p[0].reshape(h, w) if fmt == "Y800" else p[0].reshape(h, w, 3)  # :738-747


def download_ocl_frame(img: np.ndarray, fmt: str, into=None):
    """FrameIngest::download_ocl_frame -> to_obs: returns the planes (tightly packed).  `into` = existing planes whose
    untouched parts (none for the supported layouts) would be preserved."""
    h, w = img.shape[:2]
    kind, sx, sy, offs = FORMATS[fmt]
    if kind == "planar":  # :526-560
        y, u, v = cv2.split(img)
        if (sx, sy) != (1, 1):
            u = cv2.resize(u, None, fx=1.0 / sx, fy=1.0 / sy, interpolation=cv2.INTER_AREA)
            v = cv2.resize(v, None, fx=1.0 / sx, fy=1.0 / sy, interpolation=cv2.INTER_AREA)
        return [y, u, v]
    if kind == "semiplanar":  # :588-604
        uv = np.ascontiguousarray(img[:, :, 1:3])
        uv = cv2.resize(uv, None, fx=0.5, fy=0.5, interpolation=cv2.INTER_AREA)
        return [np.ascontiguousarray(img[:, :, 0]), uv.reshape(h // 2, w)]
    if kind == "packed422":  # :649-678
        y_first, u_first = fmt != "UYVY", fmt != "YVYU"
        mix = np.ascontiguousarray(img[:, :, 1:3] if u_first else img[:, :, [2, 1]])
        uv = cv2.resize(mix, None, fx=0.5, fy=1.0, interpolation=cv2.INTER_AREA)   # (h, w/2, 2)
        uv = uv.reshape(h, w)                                                     # interleaved u v u v
        out = np.empty((h, w, 2), np.uint8)
        out[:, :, 0 if y_first else 1] = img[:, :, 0]
        out[:, :, 1 if y_first else 0] = uv
        return [out.reshape(h, 2 * w)]
    if kind == "packed444":  # :702-716
        out = np.empty((h, w, 4), np.uint8)
        out[:, :, 0] = 255
        out[:, :, 1:4] = img
        return [out.reshape(h, 4 * w)]
    return [img.reshape(h, -1).copy()]  # :751-757


# ---------------------------------------------------------------------------------------------------------------------
# OpenCV's CPU arithmetic, written out (imgproc/src/resize.cpp) — what formats.cu implements
# ---------------------------------------------------------------------------------------------------------------------

def linear_taps(src: int, dst: int, horizontal: bool):
    """Source index and 11-bit weights of cv::resize(INTER_LINEAR) for 8-bit data along one axis."""
    scale = 1.0 / (dst / src)
    d = np.arange(dst)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = f - s.astype(np.float32)
    if horizontal:  # weights clamped at the borders; vertical taps keep their weights and clip the row index instead
        lo = s < 0
        f[lo] = 0
        s[lo] = 0
        hi = s >= src - 1
        f[hi] = 0
        s[hi] = src - 1
    w1 = np.rint(f * np.float32(2048)).astype(np.int32)
    w0 = np.rint((np.float32(1) - f) * np.float32(2048)).astype(np.int32)
    return s, w0, w1


def resize_linear_u8(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    sh, sw = src.shape[:2]
    s = src.astype(np.int32)
    sx, a0, a1 = linear_taps(sw, dw, True)
    sy, b0, b1 = linear_taps(sh, dh, False)
    x1 = np.minimum(sx + 1, sw - 1)
    ex = (None, slice(None)) + ((None,) if s.ndim == 3 else ())
    rows = s[:, sx] * a0[ex] + s[:, x1] * a1[ex]
    y0, y1 = np.clip(sy, 0, sh - 1), np.clip(sy + 1, 0, sh - 1)
    ey = (slice(None), None) + ((None,) if s.ndim == 3 else ())
    v = ((b0[ey] * (rows[y0] >> 4)) >> 16) + ((b1[ey] * (rows[y1] >> 4)) >> 16)
    return ((v + 2) >> 2).astype(np.uint8)


def area_half(src: np.ndarray, sub_x: int, sub_y: int) -> np.ndarray:
    """cv::resize(INTER_AREA) by 1/sub_x, 1/sub_y for sub in {1, 2}.  The rounding depends on which OpenCV code path
    the channel count selects: 2x2 on 1 channel runs ResizeAreaFastVec ((sum + 2) >> 2, half up); everything else —
    2x2 on the 2-channel NV12 plane included — runs the generic loop, saturate_cast<uchar>(sum * scale) = half to even."""
    s = src.astype(np.int32)
    if (sub_x, sub_y) == (2, 2):
        t = s[0::2, 0::2] + s[0::2, 1::2] + s[1::2, 0::2] + s[1::2, 1::2]
        if src.ndim == 2:
            return ((t + 2) >> 2).astype(np.uint8)
        q, r = t >> 2, t & 3
        return (q + ((r == 3) | ((r == 2) & ((q & 1) == 1)))).astype(np.uint8)
    if (sub_x, sub_y) == (2, 1):  # generic fast area: saturate_cast<uchar>(sum * 0.5f) = round half to even
        t = s[:, 0::2] + s[:, 1::2]
        return ((t >> 1) + ((t & 1) & ((t >> 1) & 1))).astype(np.uint8)
    return src.copy()


def upload_restated(frame) -> np.ndarray:
    """upload_obs_frame with cv2.resize replaced by the written-out arithmetic."""
    fmt, w, h, p = frame["format"], frame["width"], frame["height"], frame["planes"]
    kind, sx, sy, offs = FORMATS[fmt]
    if kind == "planar":
        u, v = p[1], p[2]
        if (sx, sy) != (1, 1):
            u, v = resize_linear_u8(u, w, h), resize_linear_u8(v, w, h)
        return np.dstack([p[0], u, v])
    if kind == "semiplanar":
        uv = resize_linear_u8(p[1].reshape(h // 2, w // 2, 2), w, h)
        return np.dstack([p[0], uv[:, :, 0], uv[:, :, 1]])
    if kind == "packed422":
        q = p[0].reshape(h, w // 2, 4)
        yo, uo, vo = offs
        luma = p[0].reshape(h, w, 2)[:, :, yo]
        u, v = resize_linear_u8(q[:, :, uo], w, h), resize_linear_u8(q[:, :, vo], w, h)
        return np.dstack([luma, u, v])
    return upload_obs_frame(frame)


def download_restated(img: np.ndarray, fmt: str):
    h, w = img.shape[:2]
    kind, sx, sy, offs = FORMATS[fmt]
    if kind == "planar":
        return [np.ascontiguousarray(img[:, :, 0]), area_half(img[:, :, 1], sx, sy), area_half(img[:, :, 2], sx, sy)]
    if kind == "semiplanar":
        uv = area_half(img[:, :, 1:3], 2, 2)
        return [np.ascontiguousarray(img[:, :, 0]), uv.reshape(h // 2, w)]
    if kind == "packed422":
        yo, uo, vo = offs
        out = np.empty((h, w // 2, 4), np.uint8)
        out[:, :, uo] = area_half(img[:, :, 1], 2, 1)
        out[:, :, vo] = area_half(img[:, :, 2], 2, 1)
        out.reshape(h, w, 2)[:, :, yo] = img[:, :, 0]
        return [out.reshape(h, 2 * w)]
    return download_ocl_frame(img, fmt)
