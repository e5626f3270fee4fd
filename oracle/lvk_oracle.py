"""
ORACLE — TEST INFRASTRUCTURE ONLY.

CPU restatement of LiveVisionKit's stabilization hot path (reference commit 2f7bb70).  Only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module, and only as the checker / the CPU baseline.  The product (livevisionkit_b200/)
never imports it.

What is restated, and from where (paths relative to /root/reference/LiveVisionKit):
  * StabilizationFilter::filter / configure / restart   Filters/StabilizationFilter.cpp:42-159
  * FrameTracker::track / estimate_* / mesh constraints  Vision/FrameTracker.cpp:33-457
  * FeatureDetector::detect / propagate / configure      Vision/FeatureDetector.cpp:28-214
  * SpatialMap key/quality, VirtualGrid                  Data/SpatialMap.tpp:241-265,588-625; Math/VirtualGrid.cpp:85-250
  * PathSmoother::next / configure                       Vision/PathSmoother.cpp:36-145
  * StreamBuffer ring semantics                          Data/StreamBuffer.tpp:37-252
  * WarpMesh (set_to(H), crop_in, clamp, apply)          Math/WarpMesh.cpp:183-223,333-342,379-427
  * Homography::transform                                Math/Homography.cpp:125-130
  * scalar helpers step / EMA / crop / barycentric_rect  Functions/Math.tpp:133-265, Logic.tpp:53-65,
    fast_erase / fast_filter / ratio_of                  Functions/Container.tpp:31-129
  * VideoFrame -> GRAY                                   Data/VideoFrame.cpp:187-301
  * lvk::remap host side + FSR-EASU kernels              Functions/Image.cpp:28-151 -> oracle/easu_ref.c

Third-party arithmetic that is NOT under /root/reference is called through the cv2 wheel in this
image (OpenCV 4.13.0; the reference pins 4.8.0, Scripts/setup_deb.sh:42 — version skew stated):
cvtColor, extractChannel, resize(INTER_AREA / INTER_LINEAR_EXACT), FastFeatureDetector(TYPE_9_16),
SparsePyrLKOpticalFlow, findHomography(UsacParams), estimateAffinePartial2D, getGaussianKernel,
getPerspectiveTransform.  Eigen's LSCG (absent) is restated in oracle/lscg_ref.c.

PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures for this path and cannot
be compiled here (no OpenCV C++ headers, Eigen, Qt5, OpenCL).  Goldens under tests/golden are
produced by THIS module (tests/golden/make_golden.py) and pin the oracle only against itself.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

try:  # cv2 is present in the build image and on the GPU box (same image)
    import cv2
except Exception:  # pragma: no cover
    cv2 = None

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

f32 = np.float32


def build_native(force: bool = False) -> str:
    """Compile oracle/easu_ref.c + lscg_ref.c -> oracle/_build/liblvkoracle.so (gcc)."""
    so = os.path.join(_HERE, "_build", "liblvkoracle.so")
    srcs = [os.path.join(_HERE, "easu_ref.c"), os.path.join(_HERE, "lscg_ref.c")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def native():
    global _LIB
    if _LIB is None:
        lib = ctypes.CDLL(build_native())
        u8p, f32p, f64p, i32p = (ctypes.POINTER(ctypes.c_uint8), ctypes.POINTER(ctypes.c_float),
                                 ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int))
        lib.oracle_easu_remap_homography.argtypes = [u8p, ctypes.c_int, ctypes.c_int, ctypes.c_int, u8p, ctypes.c_int,
                                                     f64p, u8p, ctypes.c_int, ctypes.c_int]
        lib.oracle_easu_remap_homography.restype = None
        lib.oracle_easu_remap_map.argtypes = [u8p, ctypes.c_int, ctypes.c_int, ctypes.c_int, u8p, ctypes.c_int,
                                              f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, u8p, ctypes.c_int,
                                              ctypes.c_int]
        lib.oracle_easu_remap_map.restype = None
        lib.oracle_lscg_solve.argtypes = [ctypes.c_int, ctypes.c_int, i32p, i32p, f32p, f32p, f32p, ctypes.c_int,
                                          ctypes.c_float]
        lib.oracle_lscg_solve.restype = ctypes.c_int
        lib.oracle_easu_scale.argtypes = [u8p, ctypes.c_int, ctypes.c_int, ctypes.c_int, u8p, ctypes.c_int, ctypes.c_int,
                                          ctypes.c_int, ctypes.c_int, ctypes.c_int]
        lib.oracle_easu_scale.restype = None
        lib.oracle_rcas.argtypes = [u8p, ctypes.c_int, ctypes.c_int, ctypes.c_int, u8p, ctypes.c_int, ctypes.c_float,
                                    ctypes.c_int]
        lib.oracle_rcas.restype = None
        lib.oracle_rcas_kernel_sharpness.argtypes = [ctypes.c_float]
        lib.oracle_rcas_kernel_sharpness.restype = ctypes.c_float
        lib.oracle_max_threads.restype = ctypes.c_int
        _LIB = lib
    return _LIB


def _ptr(a, ty):
    return a.ctypes.data_as(ctypes.POINTER(ty))


# ---------------------------------------------------------------------------------------------------------------------
# Frame formats (Data/VideoFrame.hpp:27)
BGR, BGRA, RGB, RGBA, YUV, GRAY, UNKNOWN = range(7)


def to_gray(frame: np.ndarray, fmt: int) -> np.ndarray:
    """VideoFrame::viewAsFormat(GRAY) — Data/VideoFrame.cpp:187-301,310-317."""
    if fmt == BGR:
        return cv2.cvtColor(frame, cv2.COLOR_BGR2GRAY)
    if fmt == RGB:
        return cv2.cvtColor(frame, cv2.COLOR_RGB2GRAY)
    if fmt == BGRA:
        return cv2.cvtColor(frame, cv2.COLOR_BGRA2GRAY)
    if fmt == RGBA:
        return cv2.cvtColor(frame, cv2.COLOR_RGBA2GRAY)
    if fmt == YUV:
        return cv2.extractChannel(frame, 0)
    if fmt == GRAY:
        return frame
    raise ValueError("unknown format")


def detection_image(frame: np.ndarray, fmt: int, det_res: tuple[int, int]) -> np.ndarray:
    """gray view + cv::resize(INTER_AREA) — StabilizationFilter.cpp:98 + FrameTracker.cpp:117."""
    return cv2.resize(to_gray(frame, fmt), det_res, interpolation=cv2.INTER_AREA)


# ---------------------------------------------------------------------------------------------------------------------
# lvk::remap (Functions/Image.cpp) -> C restatement of the OpenCL kernels


def remap_homography(src: np.ndarray, t_inv: np.ndarray, background=(255, 0, 255), yuv: bool = False,
                     threads: int = 0) -> np.ndarray:
    """lvk::remap(src, dst, homography, background, inverted=true) — Image.cpp:85-151, FSR.cl:407-452."""
    src = np.ascontiguousarray(src, dtype=np.uint8)
    rows, cols = src.shape[:2]
    dst = np.empty_like(src)
    t = np.ascontiguousarray(t_inv, dtype=np.float64).reshape(9)
    bg = np.array([int(background[0]) & 255, int(background[1]) & 255, int(background[2]) & 255], dtype=np.uint8)
    native().oracle_easu_remap_homography(_ptr(src, ctypes.c_uint8), src.strides[0], rows, cols,
                                          _ptr(dst, ctypes.c_uint8), dst.strides[0], _ptr(t, ctypes.c_double),
                                          _ptr(bg, ctypes.c_uint8), int(yuv), threads)
    return dst


def remap_map(src: np.ndarray, offset_map: np.ndarray, background=(255, 0, 255), yuv: bool = False,
              threads: int = 0) -> np.ndarray:
    """lvk::remap(src, dst, offset_map, background) — Image.cpp:28-81, FSR.cl:362-403."""
    src = np.ascontiguousarray(src, dtype=np.uint8)
    m = np.ascontiguousarray(offset_map, dtype=np.float32)
    rows, cols = m.shape[:2]
    dst = np.empty((rows, cols, 3), dtype=np.uint8)
    bg = np.array([int(background[0]) & 255, int(background[1]) & 255, int(background[2]) & 255], dtype=np.uint8)
    native().oracle_easu_remap_map(_ptr(src, ctypes.c_uint8), src.strides[0], src.shape[0], src.shape[1],
                                   _ptr(dst, ctypes.c_uint8), dst.strides[0], _ptr(m, ctypes.c_float), m.strides[0],
                                   rows, cols, _ptr(bg, ctypes.c_uint8), int(yuv), threads)
    return dst


def upscale(src: np.ndarray, size: tuple[int, int], yuv: bool = False, threads: int = 0) -> np.ndarray:
    """lvk::upscale(src, dst, size, yuv) — Image.cpp:155-201, FSR.cl:326-358.  size = (width, height)."""
    src = np.ascontiguousarray(src, dtype=np.uint8)
    w, h = int(size[0]), int(size[1])
    assert w >= src.shape[1] and h >= src.shape[0] and src.shape[2] == 3  # Image.cpp:157-160
    dst = np.empty((h, w, 3), dtype=np.uint8)
    native().oracle_easu_scale(_ptr(src, ctypes.c_uint8), src.strides[0], src.shape[0], src.shape[1],
                               _ptr(dst, ctypes.c_uint8), dst.strides[0], h, w, int(yuv), threads)
    return dst


def rcas_kernel_sharpness(sharpness: float) -> np.float32:
    """std::exp2(-2.0f * (1.0f - sharpness)) in float — Image.cpp:227."""
    assert 0.0 <= sharpness <= 1.0  # LVK_ASSERT_01, Image.cpp:209
    return f32(native().oracle_rcas_kernel_sharpness(float(f32(sharpness))))


def sharpen(src: np.ndarray, sharpness: float, threads: int = 0) -> np.ndarray:
    """lvk::sharpen(src, dst, sharpness) — Image.cpp:205-233, FSR.cl:460-535 (out of place, see easu_ref.c)."""
    src = np.ascontiguousarray(src, dtype=np.uint8)
    dst = np.empty_like(src)
    native().oracle_rcas(_ptr(src, ctypes.c_uint8), src.strides[0], src.shape[0], src.shape[1],
                         _ptr(dst, ctypes.c_uint8), dst.strides[0], float(rcas_kernel_sharpness(sharpness)), threads)
    return dst


@dataclass
class ScalingFilterSettings:
    """Filters/ScalingFilter.hpp:27-32."""
    output_size: tuple = (1920, 1080)
    sharpness: float = 0.8
    yuv_input: bool = True


class ScalingFilter:
    """lvk::ScalingFilter — Filters/ScalingFilter.cpp:28-59: upscale (EASU) then sharpen (RCAS)."""

    def __init__(self, settings: "ScalingFilterSettings | None" = None):
        self.settings = settings or ScalingFilterSettings()
        s = self.settings
        assert 0.0 <= s.sharpness <= 1.0 and s.output_size[0] > 0 and s.output_size[1] > 0  # ScalingFilter.cpp:43-45

    def apply(self, frame: np.ndarray, threads: int = 0) -> np.ndarray:
        s = self.settings
        return sharpen(upscale(frame, s.output_size, s.yuv_input, threads), s.sharpness, threads)


def mesh_to_inverse_homography(offsets: np.ndarray, width: int, height: int) -> np.ndarray:
    """2x2 branch of WarpMesh::apply — Math/WarpMesh.cpp:196-217 (dst corners -> src corners)."""
    w, h = f32(width), f32(height)
    destination = np.array([[0, 0], [w, 0], [0, h], [w, h]], dtype=np.float32)
    scaling = np.array([float(width), float(height)], dtype=np.float64)  # cv::Scalar(src.cols, src.rows)
    offs = offsets.reshape(2, 2, 2).astype(np.float32)
    corner_offsets = np.stack([offs[0, 0], offs[0, 1], offs[1, 0], offs[1, 1]])
    scaled = (corner_offsets.astype(np.float64) * scaling).astype(np.float32)  # Point2f * Scalar -> Point2f
    source = (destination + scaled).astype(np.float32)
    return cv2.getPerspectiveTransform(destination, source)


def warp_mesh_apply(offsets: np.ndarray, src: np.ndarray, background=(255, 0, 255), yuv: bool = False,
                    threads: int = 0) -> np.ndarray:
    """WarpMesh::apply — Math/WarpMesh.cpp:183-223."""
    rows, cols = src.shape[:2]
    if offsets.shape[:2] != (2, 2):
        warp_map = cv2.resize(offsets, (cols, rows), interpolation=cv2.INTER_LINEAR_EXACT)
        warp_map = cv2.multiply(warp_map, (float(cols), float(rows), 0.0, 0.0))
        return remap_map(src, warp_map, background, yuv, threads)
    t = mesh_to_inverse_homography(offsets, cols, rows)
    return remap_homography(src, t, background, yuv, threads)


# ---------------------------------------------------------------------------------------------------------------------
# Settings (field names/defaults are the reference's API)


@dataclass
class StabilizationSettings:
    # FeatureDetectorSettings — Vision/FeatureDetector.hpp:28-37
    detection_resolution: tuple = (256, 256)  # (width, height)
    detection_regions: tuple = (2, 2)
    force_detection: bool = False
    max_feature_density: float = 0.20
    min_feature_density: float = 0.05
    accumulation_rate: float = 2.0
    # FrameTrackerSettings — Vision/FrameTracker.hpp:31-44
    track_local_motions: bool = True
    temporal_smoothing: float = 1.0
    local_smoothing: float = 20.0
    min_motion_samples: int = 75
    acceptance_threshold: float = 8.0
    uniformity_threshold: float = 0.20
    # PathSmootherSettings — Vision/PathSmoother.hpp:29-39
    predictive_samples: int = 10
    corrective_limits: tuple = (0.1, 0.1)
    smoothing_steps: float = 20.0
    response_rate: float = 0.04
    # StabilizationFilterSettings — Filters/StabilizationFilter.hpp:28-39
    motion_resolution: tuple = (2, 2)
    background_colour: tuple = (255, 0, 255)
    crop_to_stable_region: bool = False
    stabilize_output: bool = True
    min_scene_quality: float = 0.8
    min_tracking_quality: float = 0.3

    @staticmethod
    def obs_homography_preset() -> "StabilizationSettings":
        """OBS 'Homography' subsystem preset — Modules/OBS-Plugin/Sources/Stabilisation/VSFilter.cpp:269-280."""
        return StabilizationSettings(detection_resolution=(480, 270), detection_regions=(2, 1),
                                     max_feature_density=0.12, min_feature_density=0.04, accumulation_rate=3.0,
                                     track_local_motions=False, acceptance_threshold=3.0, motion_resolution=(2, 2))

    @staticmethod
    def obs_field_preset() -> "StabilizationSettings":
        """OBS 'Vector Field' subsystem preset — Modules/OBS-Plugin/Sources/Stabilisation/VSFilter.cpp:257-268."""
        return StabilizationSettings(detection_resolution=(480, 270), detection_regions=(2, 2),
                                     max_feature_density=0.12, min_feature_density=0.06, accumulation_rate=3.0,
                                     track_local_motions=True, acceptance_threshold=10.0, motion_resolution=(16, 16))


def _cv_round(x: float) -> int:
    """cv::saturate_cast<int>(float) == cvRound: round half to even."""
    return int(np.rint(np.float32(x)))


FAST_MIN_THRESHOLD, FAST_MAX_THRESHOLD, FAST_THRESHOLD_STEP, FAST_FEATURE_TOLERANCE = 10, 250, 5, 150


def _step(current, target, amount):
    """lvk::step — Functions/Math.tpp:133-142."""
    if current > target:
        return max(current - amount, target)
    return min(current + amount, target)


class VirtualGrid:
    """Math/VirtualGrid.cpp:85-91,180-203."""

    def __init__(self, size, alignment):
        self.cols, self.rows = int(size[0]), int(size[1])
        self.ax, self.ay, self.aw, self.ah = (f32(alignment[0]), f32(alignment[1]), f32(alignment[2]),
                                              f32(alignment[3]))
        self.kw = f32(self.aw / f32(self.cols))
        self.kh = f32(self.ah / f32(self.rows))

    def test_point(self, x, y) -> bool:
        x, y = f32(x), f32(y)
        return bool(self.ax <= x < f32(self.ax + self.aw) and self.ay <= y < f32(self.ay + self.ah))

    def key_of(self, x, y):
        kx = int(f32(f32(f32(x) - self.ax) / self.kw))
        ky = int(f32(f32(f32(y) - self.ay) / self.kh))
        return kx, ky

    def key_to_point(self, kx, ky):
        return f32(f32(kx) * self.kw), f32(f32(ky) * self.kh)


@dataclass
class Feature:
    """cv::KeyPoint subset used by LVK: pt, response, class_id (= age)."""
    x: np.float32
    y: np.float32
    response: float
    class_id: int

    def copy(self):
        return Feature(self.x, self.y, self.response, self.class_id)


class FeatureDetector:
    """Vision/FeatureDetector.cpp."""

    def __init__(self, s: StabilizationSettings):
        self.fast = cv2.FastFeatureDetector_create(FAST_MIN_THRESHOLD, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
        self.configure(s)

    def configure(self, s: StabilizationSettings):  # FeatureDetector.cpp:48-83
        self.s = s
        W, H = s.detection_resolution
        md = f32(s.max_feature_density)
        self.grid_cols = _cv_round(f32(W) * md)
        self.grid_rows = _cv_round(f32(H) * md)
        self.grid = VirtualGrid((self.grid_cols, self.grid_rows), (0, 0, W, H))
        self.reg_cols, self.reg_rows = s.detection_regions
        self.reg_grid = VirtualGrid((self.reg_cols, self.reg_rows), (0, 0, W, H))
        # construct_detection_regions — :87-110
        self.regions = []
        for r in range(self.reg_rows):
            for c in range(self.reg_cols):
                bounds = (f32(f32(c) * self.reg_grid.kw), f32(f32(r) * self.reg_grid.kh), self.reg_grid.kw,
                          self.reg_grid.kh)
                self.regions.append({"bounds": bounds, "threshold": FAST_MIN_THRESHOLD, "load": 0})
        max_features = self.grid_cols * self.grid_rows
        max_regions = f32(self.reg_cols * self.reg_rows)
        max_region_features = f32(f32(max_features) / max_regions)
        density_ratio = f32(f32(s.min_feature_density) / f32(s.max_feature_density))
        self.min_feature_load = int(f32(max_region_features * density_ratio))
        self.fast_feature_target = int(f32(f32(s.accumulation_rate) * max_region_features))
        self.cell = {}  # key -> index into self.features   (SpatialMap<size_t>)
        self.cell_order = []  # insertion order of keys      (m_Data order, used by distribution_quality)
        self.features: list[Feature] = []
        self.last_fast_counts = []

    def max_feature_capacity(self):
        return self.grid_cols * self.grid_rows

    def reset(self):  # :209-214
        self.cell.clear()
        self.cell_order.clear()
        for r in self.regions:
            r["load"] = 0

    @staticmethod
    def region_rect(bounds):
        """cv::Rect2f -> cv::Rect conversion used by frame(bounds): saturate_cast<int> (round)."""
        return _cv_round(bounds[0]), _cv_round(bounds[1]), _cv_round(bounds[2]), _cv_round(bounds[3])

    def fast_region(self, frame, bounds, threshold):
        x, y, w, h = self.region_rect(bounds)
        self.fast.setThreshold(int(threshold))
        return self.fast.detect(frame[y:y + h, x:x + w], None)

    def detect(self, frame: np.ndarray):  # :114-178
        self.last_fast_counts = []
        for region in self.regions:
            if self.s.force_detection or region["load"] <= self.min_feature_load:
                kps = self.fast_region(frame, region["bounds"], region["threshold"])
                bx, by = region["bounds"][0], region["bounds"][1]
                for kp in kps:
                    fx, fy = f32(f32(kp.pt[0]) + bx), f32(f32(kp.pt[1]) + by)
                    feat = Feature(fx, fy, float(kp.response), 0)
                    key = self.grid.key_of(fx, fy)
                    if key not in self.cell:
                        self.cell[key] = len(self.features)
                        self.cell_order.append(key)
                        self.features.append(feat)
                    else:
                        mx = self.features[self.cell[key]]
                        if feat.response > mx.response and mx.class_id <= 0:
                            self.features[self.cell[key]] = feat
                n = len(kps)
                self.last_fast_counts.append(n)
                if n > self.fast_feature_target + FAST_FEATURE_TOLERANCE:
                    region["threshold"] = _step(region["threshold"], FAST_MAX_THRESHOLD, FAST_THRESHOLD_STEP)
                elif n < (self.fast_feature_target - FAST_FEATURE_TOLERANCE) % (1 << 64):  # size_t arithmetic
                    region["threshold"] = _step(region["threshold"], FAST_MIN_THRESHOLD, FAST_THRESHOLD_STEP)
            else:
                self.last_fast_counts.append(-1)
            region["load"] = 0
        out, self.features = self.features, []
        quality = self.distribution_quality()
        self.cell.clear()
        self.cell_order.clear()
        return out, quality

    def distribution_quality(self) -> np.float32:  # Data/SpatialMap.tpp:588-625
        n = len(self.cell_order)
        if n == 0:
            return f32(1.0)
        sectors = 4
        if self.grid_cols <= sectors or self.grid_rows <= sectors:
            return f32(f32(n) / f32(self.grid_cols * self.grid_rows))
        sg = VirtualGrid((sectors, sectors), (0, 0, self.grid_cols, self.grid_rows))
        buckets = [0] * (sectors * sectors)
        ideal = int(f32(f32(n) / f32(sectors * sectors)))
        excess = f32(0.0)
        for (kx, ky) in self.cell_order:
            if sg.test_point(kx, ky):
                sx, sy = sg.key_of(kx, ky)
                idx = sy * sectors + sx
                buckets[idx] += 1
                if buckets[idx] > ideal:
                    excess = f32(excess + f32(1.0))
        return f32(f32(1.0) - f32(excess / f32(n - ideal)))

    def propagate(self, feats: list[Feature]):  # :182-205
        for feat in feats:
            if self.grid.test_point(feat.x, feat.y):
                key = self.grid.key_of(feat.x, feat.y)
                if key not in self.cell:
                    self.cell[key] = len(self.features)
                    self.cell_order.append(key)
                    rk = self.reg_grid.key_of(feat.x, feat.y)
                    self.regions[rk[1] * self.reg_cols + rk[0]]["load"] += 1
                    self.features.append(feat.copy())
                else:
                    mx = self.features[self.cell[key]]
                    if feat.response > mx.response and feat.class_id >= mx.class_id:
                        self.features[self.cell[key]] = feat.copy()


# ---------------------------------------------------------------------------------------------------------------------
# WarpMesh helpers on float32 offset arrays of shape (rows, cols, 2)


def mesh_identity(res):
    return np.zeros((res[1], res[0], 2), dtype=np.float32)


def homography_transform_f(H: np.ndarray, x, y):
    """Homography::transform(Point2f) -> cv::perspectiveTransform (double math, float result)."""
    x, y = float(f32(x)), float(f32(y))
    w = x * H[2, 0] + y * H[2, 1] + H[2, 2]
    if abs(w) > np.finfo(np.float64).eps:
        w = 1.0 / w
        return f32((x * H[0, 0] + y * H[0, 1] + H[0, 2]) * w), f32((x * H[1, 0] + y * H[1, 1] + H[1, 2]) * w)
    return f32(0), f32(0)


def mesh_set_to_homography(H: np.ndarray, scale, res) -> np.ndarray:
    """WarpMesh::set_to(Homography, motion_scale) — Math/WarpMesh.cpp:333-342."""
    cols, rows = res
    sw, sh = f32(scale[0]), f32(scale[1])
    csx, csy = f32(sw / f32(cols - 1)), f32(sh / f32(rows - 1))
    nfx, nfy = f32(f32(1.0) / sw), f32(f32(1.0) / sh)
    out = np.zeros((rows, cols, 2), dtype=np.float32)
    for r in range(rows):
        for c in range(cols):
            px, py = f32(f32(c) * csx), f32(f32(r) * csy)
            tx, ty = homography_transform_f(H, px, py)
            out[r, c, 0] = f32(f32(px - tx) * nfx)
            out[r, c, 1] = f32(f32(py - ty) * nfy)
    return out


def mesh_crop_in(offsets: np.ndarray, region) -> np.ndarray:
    """WarpMesh::crop_in — Math/WarpMesh.cpp:379-390. region = (x, y, w, h) float."""
    rows, cols = offsets.shape[:2]
    csx = f32(f32(f32(region[2]) - f32(1.0)) / f32(cols - 1))
    csy = f32(f32(f32(region[3]) - f32(1.0)) / f32(rows - 1))
    out = offsets.copy()
    for r in range(rows):
        for c in range(cols):
            out[r, c, 0] = f32(out[r, c, 0] + f32(f32(f32(c) * csx) + f32(region[0])))
            out[r, c, 1] = f32(out[r, c, 1] + f32(f32(f32(r) * csy) + f32(region[1])))
    return out


# ---------------------------------------------------------------------------------------------------------------------


def fast_filter(lists, keep):
    """3-vector lvk::fast_filter — Functions/Container.tpp:97-121 (reverse swap-with-last erase)."""
    for k in range(len(keep) - 1, -1, -1):
        if not keep[k]:
            for data in lists:
                data[k], data[-1] = data[-1], data[k]
                data.pop()


def usac_params(threshold: float):
    """Vision/FrameTracker.cpp:337-347."""
    p = cv2.UsacParams()
    p.threshold = float(threshold)
    p.confidence = 0.99
    p.maxIterations = 50
    p.sampler = cv2.SAMPLING_UNIFORM
    p.score = cv2.SCORE_METHOD_MAGSAC
    p.loMethod = cv2.LOCAL_OPTIM_SIGMA
    p.loIterations = 10
    p.loSampleSize = 20
    p.final_polisher = cv2.MAGSAC
    p.final_polisher_iterations = 0
    return p


class FrameTracker:
    """Vision/FrameTracker.cpp."""

    def __init__(self, s: StabilizationSettings):
        self.lk = cv2.SparsePyrLKOpticalFlow_create(
            winSize=(11, 11), maxLevel=3,
            crit=(cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 5, 0.01))
        self.detector = FeatureDetector(s)
        self.mesh_res = None
        self.configure(s)
        self.restart()

    def configure(self, s: StabilizationSettings):  # :57-93
        self.detector.configure(s)
        self.region = (f32(0), f32(0), f32(s.detection_resolution[0]), f32(s.detection_resolution[1]))
        if self.mesh_res != tuple(s.motion_resolution):
            self.mesh_res = tuple(s.motion_resolution)
            self.optimized_mesh = np.zeros(2 * self.mesh_res[0] * self.mesh_res[1], dtype=np.float32)
            self.static_rows, self.static_count = self.generate_mesh_constraints(s)
        self.s = s

    def restart(self):  # :97-104
        self.stability = f32(0)
        self.features: list[Feature] = []
        self.detector.reset()
        self.initialized = False
        self.prev = None
        self.curr = None
        self.optimized_mesh = np.zeros(2 * self.mesh_res[0] * self.mesh_res[1], dtype=np.float32)
        self.trace = {}

    # ---- estimate_global_motion, :325-375
    def estimate_global_motion(self, homography: bool, tracked, matched):
        s = self.s
        src = np.asarray(tracked, dtype=np.float32).reshape(-1, 1, 2)
        dst = np.asarray(matched, dtype=np.float32).reshape(-1, 1, 2)
        if homography:
            H, mask = cv2.findHomography(src, dst, usac_params(s.acceptance_threshold))
            if H is None:
                return None, None  # the reference asserts here (Math/Homography.cpp:89-95)
        else:
            A, mask = cv2.estimateAffinePartial2D(src, dst, None, cv2.RANSAC, float(s.acceptance_threshold), 50)
            if A is None:
                return None, None
            H = np.eye(3)
            H[:2, :] = A
        motion = mesh_set_to_homography(H, (self.region[2], self.region[3]), self.mesh_res)
        self.trace["H"] = H.copy()
        return motion, mask.reshape(-1).astype(np.uint8)

    # ---- generate_mesh_constraints, :380-457.  Returns (list of rows [(col, val)...], count)
    def generate_mesh_constraints(self, s: StabilizationSettings):
        mw, mh = self.mesh_res
        gw, gh = mw - 1, mh - 1
        rw = f32(f32(f32(mw) / f32(gw)) * self.region[2])
        rh = f32(f32(f32(mh) / f32(gh)) * self.region[3])
        grid = VirtualGrid((mw, mh), (self.region[0], self.region[1], rw, rh))
        rows = []
        ts = f32(s.temporal_smoothing)
        index = 0
        for r in range(mh):
            for c in range(mw):
                rows.append([(2 * index, ts)])
                rows.append([(2 * index + 1, ts)])
                index += 1
        v1 = -(float(grid.kw) / float(grid.kh))  # cv::Size2f::aspectRatio() returns double
        v2 = -1.0 / v1
        index = 0
        for r in range(mh):
            for c in range(mw):
                idx = index
                index += 1
                quad = 1
                if c % 4 == 0 and r % 4 == 0:
                    quad = 3
                elif (c + r) % 2 != 1 and c != 0 and r != 0 and c != mw - 2 and r != mh - 2:
                    continue
                if c >= mw - quad or r >= mh - quad:
                    continue
                i00, i10 = 2 * idx, 2 * idx + 2 * quad
                i01 = 2 * (idx + quad * mw)
                i11 = i01 + 2 * quad
                weight = f32(s.local_smoothing)
                w1 = f32(v1 * float(weight))
                w2 = f32(v2 * float(weight))
                rows.append([(i00, f32(-weight)), (i01, weight), (i01 + 1, f32(-w2)), (i11 + 1, w2)])
                rows.append([(i00 + 1, f32(-weight)), (i01, w2), (i01 + 1, weight), (i11, f32(-w2))])
                rows.append([(i00, f32(-weight)), (i10, weight), (i10 + 1, f32(-w1)), (i11 + 1, w1)])
                rows.append([(i00 + 1, f32(-weight)), (i10, w1), (i10 + 1, weight), (i11, f32(-w1))])
        return rows, len(rows)

    # ---- estimate_local_motions, :200-321
    def estimate_local_motions(self, tracked, matched):
        s = self.s
        mw, mh = self.mesh_res
        gw, gh = mw - 1, mh - 1
        rw = f32(f32(f32(mw) / f32(gw)) * self.region[2])
        rh = f32(f32(f32(mh) / f32(gh)) * self.region[3])
        grid = VirtualGrid((mw, mh), (self.region[0], self.region[1], rw, rh))
        n = len(tracked)
        m_rows = self.static_count + 2 * n
        b = np.zeros(m_rows, dtype=np.float32)
        ts = f32(s.temporal_smoothing)
        for k in range(2 * mw * mh):
            b[k] = f32(ts * self.optimized_mesh[k])
        rows = list(self.static_rows)
        feat_rows = []
        off = self.static_count
        for i in range(n):
            sx, sy = f32(tracked[i][0]), f32(tracked[i][1])
            dx, dy = f32(matched[i][0]), f32(matched[i][1])
            k00x, k00y = grid.key_of(sx, sy)
            k00x = min(max(k00x, 0), gw)
            k00y = min(max(k00y, 0), gh)
            k11x, k11y = k00x + 1, k00y + 1
            # key_to_index asserts test_key (Math/VirtualGrid.cpp:146-151); the clamp to grid_size (not
            # grid_size-1) can step outside for points on the far edge — mirror the index arithmetic.
            i00 = 2 * (k00y * mw + k00x)
            i11 = 2 * (k11y * mw + k11x)
            i10, i01 = i00 + 2, i11 - 2
            p0 = grid.key_to_point(k00x, k00y)
            p1 = grid.key_to_point(k11x, k11y)
            # barycentric_rect({p0, p1}, src) — Rect_(pt1, pt2); Functions/Math.tpp:247-265
            rx, ry = min(p0[0], p1[0]), min(p0[1], p1[1])
            rwid, rhei = f32(max(p0[0], p1[0]) - rx), f32(max(p0[1], p1[1]) - ry)
            inv_area = f32(f32(1.0) / f32(rwid * rhei))
            x1, x2 = rx, f32(rx + rwid)
            y1, y2 = ry, f32(ry + rhei)
            rx1, ry1, rx2, ry2 = f32(x2 - sx), f32(y2 - sy), f32(sx - x1), f32(sy - y1)
            # cv::Scalar is double; the products are float and then widened, Triplet<float> narrows again.
            w = [f32(f32(rx1 * ry1) * inv_area), f32(f32(rx1 * ry2) * inv_area), f32(f32(rx2 * ry2) * inv_area),
                 f32(f32(rx2 * ry1) * inv_area)]
            rowx = [(i00, w[0]), (i01, w[1]), (i11, w[2]), (i10, w[3])]
            rowy = [(i00 + 1, w[0]), (i01 + 1, w[1]), (i11 + 1, w[2]), (i10 + 1, w[3])]
            rows.append(rowx)
            rows.append(rowy)
            feat_rows.append((rowx, rowy))
            b[off] = dx
            b[off + 1] = dy
            off += 2
        ncols = 2 * mw * mh
        # CSR with duplicates summed (Eigen setFromTriplets)
        rp, ci, vv = [0], [], []
        for row in rows:
            acc = {}
            for (c, v) in row:
                acc[c] = f32(acc.get(c, f32(0)) + f32(v))
            for c in sorted(acc):
                ci.append(c)
                vv.append(acc[c])
            rp.append(len(ci))
        rp = np.asarray(rp, dtype=np.int32)
        ci = np.asarray(ci, dtype=np.int32)
        vv = np.asarray(vv, dtype=np.float32)
        x = np.ascontiguousarray(self.optimized_mesh, dtype=np.float32).copy()
        iters = native().oracle_lscg_solve(m_rows, ncols, _ptr(rp, ctypes.c_int), _ptr(ci, ctypes.c_int),
                                           _ptr(vv, ctypes.c_float), _ptr(b, ctypes.c_float),
                                           _ptr(x, ctypes.c_float), -1, -1.0)
        self.trace["lscg_iters"] = iters
        self.optimized_mesh = x
        inliers = np.zeros(n, dtype=np.uint8)
        thr = f32(s.acceptance_threshold)
        for i, (rowx, rowy) in enumerate(feat_rows):
            px = f32(0)
            for k, (c, v) in enumerate(rowx):
                t = f32(v * x[c])
                px = t if k == 0 else f32(px + t)
            py = f32(0)
            for k, (c, v) in enumerate(rowy):
                t = f32(v * x[c])
                py = t if k == 0 else f32(py + t)
            bx, by = b[self.static_count + 2 * i], b[self.static_count + 2 * i + 1]
            err = f32(abs(f32(px - bx)) + abs(f32(py - by)))
            inliers[i] = 1 if err < thr else 0
        motion = np.zeros((mh, mw, 2), dtype=np.float32)
        mesh = x.reshape(mh, mw, 2)
        for r in range(mh):
            for c in range(mw):
                ax, ay = f32(f32(c) * grid.kw), f32(f32(r) * grid.kh)
                motion[r, c, 0] = f32(f32(ax - mesh[r, c, 0]) / self.region[2])
                motion[r, c, 1] = f32(f32(ay - mesh[r, c, 1]) / self.region[3])
        return motion, inliers

    # ---- track, :108-196.  next_det = the already-resized detection image (uint8, det_res)
    def track(self, next_det: np.ndarray):
        s = self.s
        self.trace = {}
        self.stability = f32(0)
        self.prev, self.curr = self.curr, next_det
        if not self.initialized or self.prev is None or self.prev.shape != self.curr.shape:
            self.initialized = True
            return None
        self.features, distribution = self.detector.detect(self.curr)
        self.trace["detected"] = [(f.x, f.y, f.response, f.class_id) for f in self.features]
        self.trace["distribution"] = distribution
        self.trace["fast_counts"] = list(self.detector.last_fast_counts)
        if len(self.features) < s.min_motion_samples or distribution < f32(s.uniformity_threshold):
            self.features = []
            return None
        tracked = [[f.x, f.y] for f in self.features]
        pts = np.asarray(tracked, dtype=np.float32).reshape(-1, 1, 2)
        nxt, status, _err = self.lk.calc(self.prev, self.curr, pts, None)
        matched = [[f32(p[0][0]), f32(p[0][1])] for p in nxt]
        status = [int(v) for v in status.reshape(-1)]
        self.trace["lk_in"] = pts.reshape(-1, 2).copy()
        self.trace["lk_out"] = nxt.reshape(-1, 2).copy()
        self.trace["lk_status"] = np.asarray(status, dtype=np.uint8)
        fast_filter([self.features, tracked, matched], status)
        if len(matched) < s.min_motion_samples:
            self.features = []
            return None
        self.trace["tracked"] = np.asarray(tracked, dtype=np.float32)
        self.trace["matched"] = np.asarray(matched, dtype=np.float32)
        if s.track_local_motions:
            motion, inliers = self.estimate_local_motions(tracked, matched)
        else:
            motion, inliers = self.estimate_global_motion(distribution > f32(0.6), tracked, matched)
            if motion is None:
                self.trace["no_model"] = True
                self.features = []
                return None
        self.trace["inliers"] = inliers.copy()
        self.stability = f32(f32(int(np.count_nonzero(inliers == 1))) / f32(len(inliers)))
        for i in range(len(inliers) - 1, -1, -1):
            if inliers[i]:
                self.features[i].class_id += 1
                self.features[i].x, self.features[i].y = f32(matched[i][0]), f32(matched[i][1])
            else:
                self.features[i], self.features[-1] = self.features[-1], self.features[i]
                self.features.pop()
        self.detector.propagate(self.features)
        self.trace["propagated"] = [(f.x, f.y, f.response, f.class_id) for f in self.features]
        return motion


# ---------------------------------------------------------------------------------------------------------------------


class PathSmoother:
    """Vision/PathSmoother.cpp.  The trajectory StreamBuffer is always full (capacity 2n+1)."""

    def __init__(self, s: StabilizationSettings):
        self.res = None
        self.window = None
        self.smoothing_factor = 0.0  # double m_SmoothingFactor
        self.configure(s)

    def configure(self, s: StabilizationSettings):  # :36-80
        res = tuple(s.motion_resolution)
        if self.res != res:
            self.res = res
            self.window = None
        window = 2 * s.predictive_samples + 1
        if self.window != window:
            old = getattr(self, "traj", [])
            if self.window is None:
                old = []
            # resize() keeps the newest elements, pad_front() fills the front with identity meshes.
            keep = old[-window:] if old else []
            self.traj = [mesh_identity(res) for _ in range(window - len(keep))] + keep
            self.window = window
            centre = (window - 1) // 2
            self.position = self.traj[0].copy()
            for i in range(1, centre + 1):
                self.position = (self.position + self.traj[i]).astype(np.float32)
            self.base_smoothing = float(window) / 12.0
        if not hasattr(self, "trace_mesh"):
            self.trace_mesh = mesh_identity(res)
        cl = s.corrective_limits
        # crop<float>({1,1}, limits) — Functions/Math.tpp:218-233
        thc, tvc = f32(f32(1.0) * f32(cl[0])), f32(f32(1.0) * f32(cl[1]))
        self.margins = (f32(thc / f32(2)), f32(tvc / f32(2)), f32(f32(1.0) - thc), f32(f32(1.0) - tvc))
        self.scene_crop = mesh_crop_in(mesh_identity(res), self.margins)
        self.s = s

    def restart(self):  # :139-145
        for m in self.traj:
            m[:] = 0
        self.position[:] = 0
        self.trace_mesh[:] = 0

    def time_delay(self):
        return self.s.predictive_samples

    def next(self, motion: np.ndarray) -> np.ndarray:  # :84-135
        s = self.s
        self.position = (self.position - self.traj[0]).astype(np.float32)
        self.traj.pop(0)
        self.traj.append(motion.astype(np.float32).copy())
        centre = (len(self.traj) - 1) // 2
        self.position = (self.position + self.traj[centre]).astype(np.float32)

        filt = cv2.getGaussianKernel(len(self.traj), self.base_smoothing + self.smoothing_factor, cv2.CV_32F)
        filt = filt.reshape(-1)
        weight = f32(1.0)
        trace = self.traj[0].copy()
        for i in range(1, len(self.traj)):
            weight = f32(weight - filt[i - 1])
            trace = (trace + self.traj[i] * weight).astype(np.float32)  # scaleAdd
        self.trace_mesh = trace
        correction = (trace - self.position).astype(np.float32)

        max_drift = f32(0.0)
        mx, my = self.margins[0], self.margins[1]
        for r in range(correction.shape[0]):
            for c in range(correction.shape[1]):
                xd = f32(abs(correction[r, c, 0]) / mx)
                yd = f32(abs(correction[r, c, 1]) / my)
                max_drift = max(max_drift, xd)
                max_drift = max(max_drift, yd)
        if max_drift > f32(1.0):
            correction[..., 0] = np.clip(correction[..., 0], -mx, mx)
            correction[..., 1] = np.clip(correction[..., 1], -my, my)
            max_drift = f32(1.0)

        d = float(max_drift)
        if d >= 0.7:
            target = 0.0
        elif d <= 0.3:
            target = float(f32(s.smoothing_steps))
        else:
            target = d
        self.smoothing_factor = self.smoothing_factor + float(f32(s.response_rate)) * (target - self.smoothing_factor)
        self.last_drift = max_drift
        return correction


# ---------------------------------------------------------------------------------------------------------------------

QA_UPDATE_RATE = f32(0.1)
QA_BLEND_STEP = f32(0.05)


class StabilizationFilter:
    """Filters/StabilizationFilter.cpp.  apply() mirrors VideoFilter::apply -> filter (Filters/VideoFilter.cpp:46-58):
    returns (output ndarray | None, timestamp | None); None == the reference's released/empty output frame."""

    def __init__(self, settings: StabilizationSettings | None = None, remap_threads: int = 0):
        s = settings or StabilizationSettings()
        self.tracker = FrameTracker(s)
        self.smoother = PathSmoother(s)
        self.scene_quality = f32(0)
        self.trust = f32(0)
        self.queue = []
        self.remap_threads = remap_threads
        self.s = None
        self.configure(s)
        self.trace = {}

    def configure(self, s: StabilizationSettings):  # :42-65
        assert 0.0 <= s.min_tracking_quality <= 1.0 and 0.0 <= s.min_scene_quality <= 1.0
        if self.s is not None and self.s.stabilize_output and not s.stabilize_output:
            self.reset_context()
        self.s = s
        self.null_motion = mesh_identity(s.motion_resolution)
        self.smoother.configure(s)
        self.queue_capacity = self.smoother.time_delay() + 1
        self.tracker.configure(s)

    def restart(self):  # :139-144
        self.scene_quality = f32(1.0)
        self.queue = []
        self.reset_context()

    def reset_context(self):  # :155-159
        self.tracker.restart()
        self.smoother.restart()

    def ready(self):
        return len(self.queue) == self.queue_capacity

    def frame_delay(self):
        return self.smoother.time_delay()

    def apply(self, frame: np.ndarray, fmt: int = BGR, timestamp: int = 0):  # filter(), :69-135
        s = self.s
        self.trace = {}
        if not s.stabilize_output:
            self._push((frame, fmt, timestamp))
            if self.ready():
                out, ofmt, ots = self.queue.pop(0)
                if s.crop_to_stable_region:
                    out = warp_mesh_apply(self.smoother.scene_crop, out, (0, 0, 0), ofmt == YUV, self.remap_threads)
                return out, ots
            return None, None

        det = detection_image(frame, fmt, tuple(s.detection_resolution))
        self.trace["det"] = det
        motion = self.tracker.track(det)
        self.trace.update(self.tracker.trace)
        self.trace["has_motion"] = motion is not None
        if motion is None:
            motion = self.null_motion.copy()
        self.trace["motion_raw"] = motion.copy()

        q = f32(self.tracker.stability)
        self.scene_quality = f32(self.scene_quality + f32(QA_UPDATE_RATE * f32(q - self.scene_quality)))
        if q < f32(s.min_tracking_quality):
            self.trust = f32(0.0)
        elif self.scene_quality < f32(s.min_scene_quality):
            self.trust = f32(_step(self.trust, f32(0.0), QA_BLEND_STEP))
        else:
            self.trust = f32(_step(self.trust, f32(1.0), QA_BLEND_STEP))
        motion = (motion * self.trust).astype(np.float32)
        self.trace["stability"] = q
        self.trace["scene_quality"] = self.scene_quality
        self.trace["trust"] = self.trust

        self._push((frame, fmt, timestamp))
        correction = self.smoother.next(motion)
        self.trace["correction"] = correction.copy()
        if self.ready():
            nxt, nfmt, nts = self.queue.pop(0)
            if s.crop_to_stable_region:
                correction = (correction + self.smoother.scene_crop).astype(np.float32)
            self.trace["applied"] = correction.copy()
            if correction.shape[:2] == (2, 2):
                self.trace["warp_T"] = mesh_to_inverse_homography(correction, nxt.shape[1], nxt.shape[0])
            out = warp_mesh_apply(correction, nxt, s.background_colour, nfmt == YUV, self.remap_threads)
            return out, nts
        return None, None

    def _push(self, item):
        # StreamBuffer::push on a full ring overwrites the oldest (Data/StreamBuffer.tpp:37-84)
        if len(self.queue) == self.queue_capacity:
            self.queue.pop(0)
        self.queue.append(item)
