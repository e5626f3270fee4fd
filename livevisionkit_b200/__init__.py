"""
livevisionkit_b200 — B200-native LiveVisionKit stabilization path.

Python mirror of the reference's filter interface for this path (lvk::VideoFilter::apply /
lvk::StabilizationFilter, LiveVisionKit/Filters/{VideoFilter,StabilizationFilter}.hpp) on top of the C-ABI
in include/lvkb200.h.  All compute happens in liblvkb200.so (hand-written CUDA, sm_100a); this module only
marshals pointers.  There is no CPU fallback: importing works without a GPU (so the ABI can be checked),
but creating a filter requires a CUDA device and a built library.

Frames are HxWx3 uint8, either numpy arrays (host memory) or torch CUDA tensors (device memory, used in place).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _capi
from ._capi import (BGR, BGRA, GRAY, RGB, RGBA, UNKNOWN, YUV, LvkB200Error, STAGE_NAMES)  # noqa: F401

__all__ = ["StabilizationFilterSettings", "StabilizationFilter", "DeblockingFilterSettings", "DeblockingFilter",
           "ScalingFilterSettings", "ScalingFilter",
           "CompositeFilter", "VideoFrame", "FrameRef", "Stream", "BGR", "RGB", "YUV", "LvkB200Error", "device_count"]


def device_count() -> int:
    return _capi.load().lvkb200_device_count()


def set_remap_exact(exact: bool) -> None:
    """Selects the arithmetic build of the EASU kernels (lvk::remap / lvk::upscale): False = contract build (default),
    True = exact build, bit-identical to oracle/easu_ref.c.  Process wide; see include/lvkb200.h."""
    _capi.load().lvkb200_set_remap_exact(1 if exact else 0)


def remap_exact() -> bool:
    return bool(_capi.load().lvkb200_remap_exact())


@dataclass
class StabilizationFilterSettings:
    """lvk::StabilizationFilterSettings (+ bases); field names and defaults are the reference's
    (Filters/StabilizationFilter.hpp:28-39, Vision/FrameTracker.hpp:31-44, Vision/FeatureDetector.hpp:28-37,
    Vision/PathSmoother.hpp:29-39).  Sizes are (width, height)."""
    detection_resolution: tuple = (256, 256)
    detection_regions: tuple = (2, 2)
    force_detection: bool = False
    max_feature_density: float = 0.20
    min_feature_density: float = 0.05
    accumulation_rate: float = 2.0
    motion_resolution: tuple = (2, 2)
    track_local_motions: bool = True
    temporal_smoothing: float = 1.0
    local_smoothing: float = 20.0
    min_motion_samples: int = 75
    acceptance_threshold: float = 8.0
    uniformity_threshold: float = 0.20
    predictive_samples: int = 10
    corrective_limits: tuple = (0.1, 0.1)
    smoothing_steps: float = 20.0
    response_rate: float = 0.04
    background_colour: tuple = (255, 0, 255)
    crop_to_stable_region: bool = False
    stabilize_output: bool = True
    min_scene_quality: float = 0.8
    min_tracking_quality: float = 0.3

    @staticmethod
    def obs_homography_preset() -> "StabilizationFilterSettings":
        """OBS 'Homography' preset — Modules/OBS-Plugin/Sources/Stabilisation/VSFilter.cpp:269-280."""
        return StabilizationFilterSettings(detection_resolution=(480, 270), detection_regions=(2, 1),
                                           max_feature_density=0.12, min_feature_density=0.04, accumulation_rate=3.0,
                                           track_local_motions=False, acceptance_threshold=3.0)

    @staticmethod
    def obs_field_preset() -> "StabilizationFilterSettings":
        """OBS 'Vector Field' preset — Modules/OBS-Plugin/Sources/Stabilisation/VSFilter.cpp:257-268."""
        return StabilizationFilterSettings(detection_resolution=(480, 270), detection_regions=(2, 2),
                                           max_feature_density=0.12, min_feature_density=0.06, accumulation_rate=3.0,
                                           track_local_motions=True, acceptance_threshold=10.0,
                                           motion_resolution=(16, 16))

    def to_c(self) -> _capi.Settings:
        s = _capi.Settings()
        s.detection_resolution_width, s.detection_resolution_height = self.detection_resolution
        s.detection_regions_width, s.detection_regions_height = self.detection_regions
        s.force_detection = int(self.force_detection)
        s.max_feature_density = self.max_feature_density
        s.min_feature_density = self.min_feature_density
        s.accumulation_rate = self.accumulation_rate
        s.motion_resolution_width, s.motion_resolution_height = self.motion_resolution
        s.track_local_motions = int(self.track_local_motions)
        s.temporal_smoothing = self.temporal_smoothing
        s.local_smoothing = self.local_smoothing
        s.min_motion_samples = self.min_motion_samples
        s.acceptance_threshold = self.acceptance_threshold
        s.uniformity_threshold = self.uniformity_threshold
        s.predictive_samples = self.predictive_samples
        s.corrective_limits_width, s.corrective_limits_height = self.corrective_limits
        s.smoothing_steps = self.smoothing_steps
        s.response_rate = self.response_rate
        bg = list(self.background_colour) + [0.0] * (4 - len(self.background_colour))
        for i in range(4):
            s.background_colour[i] = float(bg[i])
        s.crop_to_stable_region = int(self.crop_to_stable_region)
        s.stabilize_output = int(self.stabilize_output)
        s.min_scene_quality = self.min_scene_quality
        s.min_tracking_quality = self.min_tracking_quality
        return s


@dataclass
class VideoFrame:
    """lvk::VideoFrame (Data/VideoFrame.hpp:25-31): pixel buffer + timestamp + format."""
    data: object = None
    timestamp: int = 0
    format: int = BGR

    def empty(self) -> bool:
        return self.data is None


class FrameRef:
    """A frame buffer (NumPy array or torch tensor) with its pointer / pitch / geometry looked up ONCE.  Every call of
    the mirror otherwise re-derives them from the array object (~5 us of Python per buffer, three buffers per frame —
    a fifth of a 100 us frame); a caller that reuses its buffers wraps them once and passes the FrameRef wherever a
    buffer is accepted.  The wrapped object is kept alive; `.buf` returns it."""
    __slots__ = ("buf", "info")

    def __init__(self, buf):
        self.buf = buf
        self.info = _buffer_info(buf)


def _buffer_info(buf):
    """-> (pointer, pitch_bytes, height, width, channels, memspace)."""
    if type(buf) is FrameRef:
        return buf.info
    if isinstance(buf, np.ndarray):
        if buf.dtype != np.uint8 or buf.ndim not in (2, 3) or not buf.flags["C_CONTIGUOUS"] and buf.strides[-1] != 1:
            raise ValueError("frames must be uint8 HxW[xC] arrays with contiguous rows")
        ch = 1 if buf.ndim == 2 else buf.shape[2]
        if buf.ndim == 3 and buf.strides[1] != ch:
            raise ValueError("pixels must be packed")
        return buf.ctypes.data, buf.strides[0], buf.shape[0], buf.shape[1], ch, _capi.MEM_HOST
    if hasattr(buf, "data_ptr") and hasattr(buf, "is_cuda"):  # torch tensor (plumbing only)
        if str(buf.dtype) != "torch.uint8" or buf.dim() not in (2, 3):
            raise ValueError("frames must be uint8 HxW[xC] tensors")
        ch = 1 if buf.dim() == 2 else buf.shape[2]
        if buf.stride(-1) != 1 or (buf.dim() == 3 and buf.stride(1) != ch):
            raise ValueError("pixels must be packed")
        space = _capi.MEM_DEVICE if buf.is_cuda else _capi.MEM_HOST
        return buf.data_ptr(), buf.stride(0), buf.shape[0], buf.shape[1], ch, space
    raise TypeError(f"unsupported frame buffer type {type(buf)}")


class BatchPlan:
    """Pointer tables of a frame sequence for Stream.submit_batch, built once (buffers are usually reused)."""

    def __init__(self, frames, outs, timestamps=None):
        frames, outs = list(frames), list(outs)
        if len(frames) != len(outs):
            raise ValueError("one output buffer per frame")
        self.count = len(frames)
        infos, oinfos = [_buffer_info(f) for f in frames], [_buffer_info(o) for o in outs]
        ptr0, self.pitch, self.h, self.w, ch, self.space = infos[0] if infos else (0, 0, 0, 0, 3, _capi.MEM_HOST)
        optr0, self.opitch, oh, ow, och, self.ospace = oinfos[0] if oinfos else (0, 0, 0, 0, 3, _capi.MEM_HOST)
        for inf in infos:
            if inf[1:] != (self.pitch, self.h, self.w, ch, self.space):
                raise ValueError("submit_batch takes frames of one geometry, pitch and memory space")
        for inf in oinfos:
            if inf[1:] != (self.opitch, self.h, self.w, ch, self.ospace):
                raise ValueError("output buffers must match the frames")
        self.frames = (C.c_void_p * self.count)(*[inf[0] for inf in infos])
        self.outs = (C.c_void_p * self.count)(*[inf[0] for inf in oinfos])
        self.timestamps = None if timestamps is None else (C.c_uint64 * self.count)(*[int(t) for t in timestamps])
        self.results = (_capi.Result * self.count)()
        self._keep = (frames, outs)


def _f32(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a if shape is None else a.reshape(shape)


class Stream:
    """Thin object wrapper over a lvkb200_stream handle (one video stream / one CUDA stream)."""

    def __init__(self, settings: StabilizationFilterSettings | None = None, device: int = 0):
        self._lib = _capi.load()
        self._h = C.c_void_p()
        cs = (settings or StabilizationFilterSettings()).to_c()
        _capi.check(self._lib.lvkb200_stream_create(device, C.byref(cs), C.byref(self._h)))
        self.device = device

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.lvkb200_stream_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- filter-level
    def configure(self, settings: StabilizationFilterSettings):
        cs = settings.to_c()
        _capi.check(self._lib.lvkb200_stream_configure(self._h, C.byref(cs)))

    def restart(self):
        _capi.check(self._lib.lvkb200_stream_restart(self._h))

    def reset_context(self):
        _capi.check(self._lib.lvkb200_stream_reset_context(self._h))

    def ready(self) -> bool:
        return bool(self._lib.lvkb200_stream_ready(self._h))

    def frame_delay(self) -> int:
        return int(self._lib.lvkb200_stream_frame_delay(self._h))

    def stable_region(self, width: int, height: int):
        x, y, w, h = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        _capi.check(self._lib.lvkb200_stream_stable_region(self._h, width, height, C.byref(x), C.byref(y), C.byref(w),
                                                           C.byref(h)))
        return x.value, y.value, w.value, h.value

    def sync(self):
        _capi.check(self._lib.lvkb200_stream_sync(self._h))

    def submit(self, frame, out, fmt: int = BGR, timestamp: int = 0) -> _capi.Result:
        ptr, pitch, h, w, ch, space = _buffer_info(frame)
        optr, opitch, oh, ow, och, ospace = _buffer_info(out)
        if (oh, ow, och) != (h, w, ch):
            raise ValueError("output buffer must match the input frame")
        res = _capi.Result()
        _capi.check(self._lib.lvkb200_stream_submit(self._h, ptr, pitch, w, h, fmt, timestamp, space, optr, opitch,
                                                    ospace, C.byref(res)))
        return res

    def __call__(self, frames, callback, outputs=None) -> int:
        """Pipelined VideoFilter::stream: see StabilizationFilter.stream_frames."""
        frames = list(frames)
        if not frames:
            return 0
        if outputs is None:
            outputs = [np.empty_like(frames[0].data) for _ in range(3)]
        if len(outputs) < 3:
            raise ValueError("stream() needs at least 3 output buffers")
        # Two outputs stay in flight: output t's remap is launched inside submit t+1 (beside that frame's LK + RANSAC,
        # see lvkb200_stream_submit_async) and its download overlaps frame t+2, so it is collected after submit t+2.
        pending, delivered = [], 0
        for i, f in enumerate(frames):
            if i + 1 < len(frames):
                self.prefetch(frames[i + 1].data, frames[i + 1].format)
            out = outputs[i % len(outputs)]
            res, ticket = self.submit_async(f.data, out, f.format, f.timestamp)
            if res.has_output:
                pending.append((ticket, VideoFrame(out, int(res.out_timestamp), int(res.out_format))))
            while len(pending) > 2:
                tk, vf = pending.pop(0)
                self.wait_output(tk)
                delivered += 1
                if callback(vf):  # VideoFilter.cpp:180: a true return terminates the stream
                    for tk, _ in pending:  # let the queued downloads finish before the buffers are handed back
                        self.wait_output(tk)
                    return delivered
        for tk, vf in pending:
            self.wait_output(tk)
            delivered += 1
            if callback(vf):
                break
        for tk, _ in pending:
            self.wait_output(tk)
        return delivered

    def submit_batch(self, frames, outs, fmt: int = BGR, timestamps=None):
        """lvkb200_stream_submit_batch: filters `frames` back to back, frame i into outs[i] (buffers of one geometry, all
        host or all device), announcing frame i+1 before frame i is submitted - one FFI call for the whole sequence.
        Returns the list of per-step Results.  `frames` / `outs` may be a BatchPlan built once for reused buffers."""
        plan = frames if isinstance(frames, BatchPlan) else BatchPlan(frames, outs, timestamps)
        _capi.check(self._lib.lvkb200_stream_submit_batch(self._h, plan.frames, plan.pitch, plan.w, plan.h, fmt, plan.timestamps,
                                                          plan.space, plan.outs, plan.opitch, plan.ospace, plan.count, plan.results))
        return plan.results

    def prefetch(self, frame, fmt: "int | None" = None):
        """Announces the NEXT frame (call it before submitting the current one).  Without `fmt`: starts the upload of
        a host frame (lvkb200_stream_prefetch).  With `fmt`: host or device frame, and the library also builds its
        detection image and pyramid behind the current frame's tracking chain (lvkb200_stream_prefetch_frame)."""
        ptr, pitch, h, w, ch, space = _buffer_info(frame)
        if ch != 3:
            raise ValueError("prefetch takes packed 8UC3 frames")
        if fmt is None:
            if space != _capi.MEM_HOST:
                raise ValueError("prefetch of a device frame needs its format")
            _capi.check(self._lib.lvkb200_stream_prefetch(self._h, ptr, pitch, w, h))
        else:
            _capi.check(self._lib.lvkb200_stream_prefetch_frame(self._h, ptr, pitch, w, h, int(fmt), space))

    def submit_async(self, frame, out, fmt: int = BGR, timestamp: int = 0):
        """-> (Result, ticket).  Like submit(); a host `out` is filled after the call (see wait_output)."""
        ptr, pitch, h, w, ch, space = _buffer_info(frame)
        optr, opitch, oh, ow, och, ospace = _buffer_info(out)
        if (oh, ow, och) != (h, w, ch):
            raise ValueError("output buffer must match the input frame")
        res = _capi.Result()
        ticket = C.c_uint64(0)
        _capi.check(self._lib.lvkb200_stream_submit_async(self._h, ptr, pitch, w, h, fmt, timestamp, space, optr,
                                                          opitch, ospace, C.byref(res), C.byref(ticket)))
        return res, int(ticket.value)

    def wait_output(self, ticket: int):
        _capi.check(self._lib.lvkb200_stream_wait_output(self._h, ticket))

    def event_record(self, index: int):
        _capi.check(self._lib.lvkb200_stream_event_record(self._h, index))

    def event_elapsed_ms(self, start: int, stop: int) -> float:
        ms = C.c_float(0)
        _capi.check(self._lib.lvkb200_stream_event_elapsed_ms(self._h, start, stop, C.byref(ms)))
        return float(ms.value)

    def stage_times_us(self) -> dict:
        t = (C.c_float * _capi.STAGE_COUNT)()
        _capi.check(self._lib.lvkb200_stream_stage_times_us(self._h, t))
        return dict(zip(STAGE_NAMES, [float(v) for v in t]))

    def stage_totals_us(self, reset: bool = False):
        """-> ({stage: total_us}, {stage: samples}) accumulated since the last reset (CUDA events, own stream)."""
        t = (C.c_double * _capi.STAGE_COUNT)()
        n = (C.c_uint64 * _capi.STAGE_COUNT)()
        _capi.check(self._lib.lvkb200_stream_stage_totals_us(self._h, t, n, int(reset)))
        return dict(zip(STAGE_NAMES, [float(v) for v in t])), dict(zip(STAGE_NAMES, [int(v) for v in n]))

    def set_profiling(self, enable: bool = True):
        _capi.check(self._lib.lvkb200_stream_set_profiling(self._h, int(enable)))

    def set_debug_capture(self, enable: bool = True):
        _capi.check(self._lib.lvkb200_stream_set_debug_capture(self._h, int(enable)))

    def debug_fetch(self, which: int, dtype, shape_tail=()):
        size = C.c_size_t(0)
        _capi.check(self._lib.lvkb200_stream_debug_fetch(self._h, which, None, 0, C.byref(size)))
        if size.value == 0:
            return None
        raw = np.empty(size.value, dtype=np.uint8)
        _capi.check(self._lib.lvkb200_stream_debug_fetch(self._h, which, raw.ctypes.data, raw.nbytes, C.byref(size)))
        arr = raw.view(dtype)
        return arr.reshape((-1,) + tuple(shape_tail)) if shape_tail else arr

    # ---- stage-level (parity tests: oracle inputs -> one GPU stage)
    def remap_homography(self, src, t_inv, background=(255, 0, 255), yuv: bool = False, out=None):
        ptr, pitch, h, w, ch, space = _buffer_info(src)
        if ch != 3:
            raise ValueError("remap needs 8UC3 frames")  # Image.cpp:96
        if out is None:
            out = np.empty((h, w, 3), dtype=np.uint8)
        optr, opitch, _, _, _, ospace = _buffer_info(out)
        t = np.ascontiguousarray(t_inv, dtype=np.float64).reshape(9)
        bg = (C.c_uint8 * 3)(*[int(v) & 255 for v in background[:3]])
        _capi.check(self._lib.lvkb200_remap_homography(self._h, ptr, pitch, w, h, space, optr, opitch, ospace,
                                                       t.ctypes.data_as(C.POINTER(C.c_double)), bg, int(yuv)))
        return out

    def remap_mesh(self, src, offsets, background=(255, 0, 255), yuv: bool = False, out=None):
        ptr, pitch, h, w, ch, space = _buffer_info(src)
        if out is None:
            out = np.empty((h, w, 3), dtype=np.uint8)
        optr, opitch, _, _, _, ospace = _buffer_info(out)
        m = _f32(offsets)
        rows, cols = m.shape[:2]
        bg = (C.c_uint8 * 3)(*[int(v) & 255 for v in background[:3]])
        _capi.check(self._lib.lvkb200_remap_mesh(self._h, ptr, pitch, w, h, space, optr, opitch, ospace,
                                                 m.ctypes.data_as(C.POINTER(C.c_float)), cols, rows, bg, int(yuv)))
        return out

    def warp_mesh_apply(self, src, offsets, background=(255, 0, 255), yuv: bool = False, out=None):
        """WarpMesh::apply.  Returns (out, t_inv) — t_inv is the dst->src transform of the 2x2 branch."""
        ptr, pitch, h, w, ch, space = _buffer_info(src)
        if out is None:
            out = np.empty((h, w, 3), dtype=np.uint8)
        optr, opitch, _, _, _, ospace = _buffer_info(out)
        m = _f32(offsets)
        rows, cols = m.shape[:2]
        t = np.zeros(9, dtype=np.float64)
        bg = (C.c_uint8 * 3)(*[int(v) & 255 for v in background[:3]])
        _capi.check(self._lib.lvkb200_warp_mesh_apply(self._h, ptr, pitch, w, h, space, optr, opitch, ospace,
                                                      m.ctypes.data_as(C.POINTER(C.c_float)), cols, rows, bg,
                                                      int(yuv), t.ctypes.data_as(C.POINTER(C.c_double))))
        return out, t.reshape(3, 3)

    def detection_image(self, frame, fmt: int, det_res):
        ptr, pitch, h, w, ch, space = _buffer_info(frame)
        det = np.empty((det_res[1], det_res[0]), dtype=np.uint8)
        _capi.check(self._lib.lvkb200_detection_image(self._h, ptr, pitch, w, h, fmt, space,
                                                      det.ctypes.data_as(C.POINTER(C.c_uint8)), det_res[0],
                                                      det_res[1]))
        return det

    def fast_detect(self, image: np.ndarray, roi, threshold: int, capacity: int = 1 << 16):
        img = np.ascontiguousarray(image, dtype=np.uint8)
        kps = (_capi.KeyPoint * capacity)()
        n = C.c_int(0)
        _capi.check(self._lib.lvkb200_fast_detect(self._h, img.ctypes.data_as(C.POINTER(C.c_uint8)), img.shape[1],
                                                  img.shape[0], roi[0], roi[1], roi[2], roi[3], int(threshold), kps,
                                                  capacity, C.byref(n)))
        arr = np.frombuffer(kps, dtype=np.dtype([("x", "f4"), ("y", "f4"), ("response", "f4"), ("class_id", "i4")]),
                            count=n.value)
        return arr.copy()

    def lk_track(self, prev: np.ndarray, nxt: np.ndarray, points, call_index: int = 0):
        p = np.ascontiguousarray(prev, dtype=np.uint8)
        q = np.ascontiguousarray(nxt, dtype=np.uint8)
        pts = _f32(points, (-1, 2))
        n = pts.shape[0]
        matched = np.empty((n, 2), dtype=np.float32)
        status = np.empty(n, dtype=np.uint8)
        _capi.check(self._lib.lvkb200_lk_track(self._h, p.ctypes.data_as(C.POINTER(C.c_uint8)),
                                               q.ctypes.data_as(C.POINTER(C.c_uint8)), p.shape[1], p.shape[0],
                                               pts.ctypes.data_as(C.POINTER(C.c_float)), n, int(call_index),
                                               matched.ctypes.data_as(C.POINTER(C.c_float)),
                                               status.ctypes.data_as(C.POINTER(C.c_uint8))))
        return matched, status

    def find_homography(self, src_points, dst_points, threshold: float):
        a, b = _f32(src_points, (-1, 2)), _f32(dst_points, (-1, 2))
        n = a.shape[0]
        H = np.zeros(9, dtype=np.float64)
        mask = np.zeros(n, dtype=np.uint8)
        _capi.check(self._lib.lvkb200_find_homography(self._h, a.ctypes.data_as(C.POINTER(C.c_float)),
                                                      b.ctypes.data_as(C.POINTER(C.c_float)), n, float(threshold),
                                                      H.ctypes.data_as(C.POINTER(C.c_double)),
                                                      mask.ctypes.data_as(C.POINTER(C.c_uint8))))
        return H.reshape(3, 3), mask

    def estimate_affine_partial(self, src_points, dst_points, threshold: float):
        a, b = _f32(src_points, (-1, 2)), _f32(dst_points, (-1, 2))
        n = a.shape[0]
        H = np.zeros(9, dtype=np.float64)
        mask = np.zeros(n, dtype=np.uint8)
        _capi.check(self._lib.lvkb200_estimate_affine_partial(self._h, a.ctypes.data_as(C.POINTER(C.c_float)),
                                                              b.ctypes.data_as(C.POINTER(C.c_float)), n,
                                                              float(threshold),
                                                              H.ctypes.data_as(C.POINTER(C.c_double)),
                                                              mask.ctypes.data_as(C.POINTER(C.c_uint8))))
        return H.reshape(3, 3), mask

    def estimate_local_motions(self, tracked, matched, mesh_state):
        a, b = _f32(tracked, (-1, 2)), _f32(matched, (-1, 2))
        n = a.shape[0]
        state = _f32(mesh_state).copy()
        offsets = np.zeros_like(state)
        mask = np.zeros(n, dtype=np.uint8)
        _capi.check(self._lib.lvkb200_estimate_local_motions(self._h, a.ctypes.data_as(C.POINTER(C.c_float)),
                                                             b.ctypes.data_as(C.POINTER(C.c_float)), n,
                                                             state.ctypes.data_as(C.POINTER(C.c_float)),
                                                             offsets.ctypes.data_as(C.POINTER(C.c_float)),
                                                             mask.ctypes.data_as(C.POINTER(C.c_uint8))))
        return state, offsets, mask

    # ---- VSFilter::filter for asynchronous OBS sources: ingest -> stabilize -> egress
    def submit_obs(self, frame: "ObsFrame", out: "ObsFrame") -> _capi.Result:
        cin, space = frame.to_c()
        cout, ospace = out.to_c()
        res = _capi.Result()
        _capi.check(self._lib.lvkb200_stream_submit_obs(self._h, C.byref(cin), space, C.byref(cout), ospace, C.byref(res)))
        if res.has_output:
            out.timestamp = cout.timestamp
        return res

    def prefetch_obs(self, frame: "ObsFrame"):
        """Announces the NEXT OBS-layout frame (host planes): upload + to_ocl conversion on the copy-in stream."""
        cin, space = frame.c_cached()
        if space != _capi.MEM_HOST:
            raise ValueError("prefetch_obs takes host planes")
        _capi.check(self._lib.lvkb200_stream_prefetch_obs(self._h, C.byref(cin)))

    def submit_obs_async(self, frame: "ObsFrame", out: "ObsFrame"):
        """-> (Result, ticket): lvkb200_stream_submit_obs_async; the planes of `out` are filled after the call
        (wait_output(ticket))."""
        cin, space = frame.c_cached()
        cout, ospace = out.c_cached()
        if space != _capi.MEM_HOST or ospace != _capi.MEM_HOST:
            raise ValueError("submit_obs_async takes host planes")
        res = _capi.Result()
        ticket = C.c_uint64(0)
        _capi.check(self._lib.lvkb200_stream_submit_obs_async(self._h, C.byref(cin), C.byref(cout), C.byref(res), C.byref(ticket)))
        if res.has_output:
            out.timestamp = int(res.out_timestamp)
        return res, int(ticket.value)

    def submit_obs_batch(self, frames, outs):
        """lvkb200_stream_submit_obs_batch: `frames[i]` -> `outs[i]` (host ObsFrames of one layout and size; `outs` may
        cycle through >= 4 buffers), pipelined inside ONE FFI call.  Returns the per-step Results; every output has landed."""
        n = len(frames)
        if len(outs) != n:
            raise ValueError("one output frame per input frame")
        cin, cout = (_capi.ObsFrame * n)(), (_capi.ObsFrame * n)()
        for i in range(n):
            a, sa = frames[i].c_cached()
            b, sb = outs[i].c_cached()
            if sa != _capi.MEM_HOST or sb != _capi.MEM_HOST:
                raise ValueError("submit_obs_batch takes host planes")
            cin[i], cout[i] = a, b
        results = (_capi.Result * n)()
        _capi.check(self._lib.lvkb200_stream_submit_obs_batch(self._h, cin, cout, n, results))
        for i in range(n):
            if results[i].has_output:
                outs[i].timestamp = int(results[i].out_timestamp)
        return results

    def stream_obs(self, frames, callback, outputs) -> int:
        """Pipelined VSFilter::filter over a sequence of host OBS-layout frames (the NV12 / I420 analogue of
        `Stream.__call__`): frame t+1 is uploaded and converted while frame t is tracked, output t-1 is converted back and
        downloaded meanwhile; two outputs stay in flight.  `outputs`: >= 3 reusable ObsFrame buffers.  Returns the number
        of outputs delivered to `callback(ObsFrame) -> bool` (a true return stops the stream)."""
        frames = list(frames)
        if len(outputs) < 3:
            raise ValueError("stream_obs() needs at least 3 output frames")
        pending, delivered = [], 0
        for i, f in enumerate(frames):
            if i + 1 < len(frames):
                self.prefetch_obs(frames[i + 1])
            out = outputs[i % len(outputs)]
            res, ticket = self.submit_obs_async(f, out)
            if res.has_output:
                pending.append((ticket, out))
            while len(pending) > 2:
                tk, o = pending.pop(0)
                self.wait_output(tk)
                delivered += 1
                if callback(o):
                    for tk2, _ in pending:
                        self.wait_output(tk2)
                    return delivered
        for tk, o in pending:
            self.wait_output(tk)
            delivered += 1
            if callback(o):
                break
        for tk, _ in pending:
            self.wait_output(tk)
        return delivered

    # ---- lvk::DeblockingFilter
    def deblock(self, frame, settings: "DeblockingFilterSettings | None" = None, fmt: int = BGR, out=None):
        """DeblockingFilter::filter on one frame (host or device); `out` defaults to a new buffer, may be `frame`."""
        ptr, pitch, h, w, ch, space = _buffer_info(frame)
        if ch != 3:
            raise ValueError("deblocking takes packed 8UC3 frames")
        if out is None:
            out = np.empty_like(frame) if isinstance(frame, np.ndarray) else frame.new_empty(frame.shape)
        optr, opitch, oh, ow, och, ospace = _buffer_info(out)
        if (oh, ow, och) != (h, w, ch):
            raise ValueError("output buffer must match the input frame")
        c = (settings or DeblockingFilterSettings()).to_c()
        _capi.check(self._lib.lvkb200_deblock(self._h, C.byref(c), ptr, pitch, w, h, fmt, space, optr, opitch, ospace))
        return out

    def set_deblocking(self, settings: "DeblockingFilterSettings | None"):
        """Chains a DeblockingFilter in front of the stabilizer, on the device (None switches it off)."""
        if settings is None:
            _capi.check(self._lib.lvkb200_stream_set_deblocking(self._h, None))
        else:
            c = settings.to_c()
            _capi.check(self._lib.lvkb200_stream_set_deblocking(self._h, C.byref(c)))

    # ---- lvk::ScalingFilter: lvk::upscale / lvk::sharpen (Functions/Image.cpp:155-233)
    @staticmethod
    def _new_like(frame, h, w):
        return np.empty((h, w, 3), dtype=np.uint8) if isinstance(frame, np.ndarray) else frame.new_empty((h, w, 3))

    def upscale(self, frame, size, yuv: bool = False, out=None):
        """lvk::upscale(src, dst, size, yuv): FSR-EASU to size = (width, height) >= the frame's."""
        ptr, pitch, h, w, ch, space = _buffer_info(frame)
        if ch != 3:
            raise ValueError("upscale takes packed 8UC3 frames")
        ow, oh = int(size[0]), int(size[1])
        if out is None:
            out = self._new_like(frame, oh, ow)
        optr, opitch, bh, bw, och, ospace = _buffer_info(out)
        if (bh, bw, och) != (oh, ow, 3):
            raise ValueError("output buffer must have the requested size")
        _capi.check(self._lib.lvkb200_upscale(self._h, ptr, pitch, w, h, space, optr, opitch, ow, oh, ospace, int(yuv)))
        return out

    def sharpen(self, frame, sharpness: float, out=None):
        """lvk::sharpen(src, dst, sharpness): FSR-RCAS; `out` may be `frame`."""
        ptr, pitch, h, w, ch, space = _buffer_info(frame)
        if ch != 3:
            raise ValueError("sharpen takes packed 8UC3 frames")
        if out is None:
            out = self._new_like(frame, h, w)
        optr, opitch, bh, bw, och, ospace = _buffer_info(out)
        if (bh, bw, och) != (h, w, 3):
            raise ValueError("output buffer must match the input frame")
        _capi.check(self._lib.lvkb200_sharpen(self._h, ptr, pitch, w, h, space, optr, opitch, ospace, float(sharpness)))
        return out

    def scaling_filter(self, frame, settings: "ScalingFilterSettings | None" = None, out=None):
        """ScalingFilter::filter: upscale then sharpen, the intermediate frame stays on the device."""
        st = settings or ScalingFilterSettings()
        ptr, pitch, h, w, ch, space = _buffer_info(frame)
        if ch != 3:
            raise ValueError("the scaling filter takes packed 8UC3 frames")
        ow, oh = int(st.output_size[0]), int(st.output_size[1])
        if out is None:
            out = self._new_like(frame, oh, ow)
        optr, opitch, bh, bw, och, ospace = _buffer_info(out)
        if (bh, bw, och) != (oh, ow, 3):
            raise ValueError("output buffer must have the configured output size")
        c = st.to_c()
        _capi.check(self._lib.lvkb200_scaling_filter(self._h, C.byref(c), ptr, pitch, w, h, space, optr, opitch, ospace))
        return out


def _plane_info(buf):
    """-> (pointer, pitch_bytes, memspace) of one 2-D uint8 plane (numpy array or torch tensor)."""
    if isinstance(buf, np.ndarray):
        if buf.dtype != np.uint8 or buf.ndim != 2 or buf.strides[1] != 1:
            raise ValueError("planes must be 2-D uint8 arrays with contiguous rows")
        return buf.ctypes.data, buf.strides[0], _capi.MEM_HOST
    if hasattr(buf, "data_ptr") and hasattr(buf, "is_cuda"):
        if str(buf.dtype) != "torch.uint8" or buf.dim() != 2 or buf.stride(1) != 1:
            raise ValueError("planes must be 2-D uint8 tensors with contiguous rows")
        return buf.data_ptr(), buf.stride(0), _capi.MEM_DEVICE if buf.is_cuda else _capi.MEM_HOST
    raise TypeError(f"unsupported plane buffer type {type(buf)}")


@dataclass
class ObsFrame:
    """The part of obs_source_frame the ingest reads: a video format name ("NV12", "I420", ...), the frame size and one
    2-D uint8 buffer (rows x row bytes) per plane — all in host memory or all in device memory."""
    format: str
    width: int
    height: int
    planes: list
    timestamp: int = 0

    def c_cached(self):
        """to_c() once per frame object (the plane buffers of a reused frame do not move)."""
        cached = self.__dict__.get("_c")
        if cached is None:
            cached = self.to_c()
            self.__dict__["_c"] = cached
        cached[0].timestamp = int(self.timestamp)
        return cached

    def to_c(self):
        c = _capi.ObsFrame()
        c.format = _capi.VIDEO_FORMATS[self.format]
        c.width, c.height, c.timestamp = int(self.width), int(self.height), int(self.timestamp)
        spaces = set()
        for i, p in enumerate(self.planes):
            ptr, pitch, space = _plane_info(p)
            c.data[i], c.linesize[i] = ptr, pitch
            spaces.add(space)
        if len(spaces) != 1:
            raise ValueError("all planes of a frame must live in the same memory space")
        return c, spaces.pop()


class FrameIngest:
    """lvk::FrameIngest (Modules/OBS-Plugin/Interop/FrameIngest.hpp:28-60): converts between OBS frame layouts and the
    packed frames the filters take.  `FrameIngest.Select(name)` returns None for a format LVK does not support."""

    def __init__(self, obs_format: str, stream: "Stream | None" = None, device: int = 0):
        self._obs_format = obs_format
        self._lib = _capi.load()
        self._ocl_format = self._lib.lvkb200_video_format_ocl(_capi.VIDEO_FORMATS[obs_format])
        self.stream = stream or Stream(None, device)

    @staticmethod
    def Select(obs_format: str, stream: "Stream | None" = None, device: int = 0):
        if obs_format not in _capi.VIDEO_FORMATS:
            return None
        return FrameIngest(obs_format, stream, device)

    def obs_format(self) -> str:
        return self._obs_format

    def ocl_format(self) -> int:
        return self._ocl_format

    def upload_obs_frame(self, src: ObsFrame, out=None) -> VideoFrame:
        """FrameIngest::upload_obs_frame: planes -> packed frame (8UC3, or 8UC1 for Y800)."""
        if src.format != self._obs_format:
            raise LvkB200Error(_capi.ERR_INVALID, "src->format == m_OBSFormat")
        c, space = src.to_c()
        ch = 1 if self._ocl_format == GRAY else 3
        if out is None:
            shape = (src.height, src.width) if ch == 1 else (src.height, src.width, 3)
            out = np.empty(shape, np.uint8) if space == _capi.MEM_HOST else src.planes[0].new_empty(shape)
        optr, opitch, oh, ow, och, ospace = _buffer_info(out)
        if (oh, ow, och) != (src.height, src.width, ch):
            raise ValueError("output buffer does not match the frame")
        _capi.check(self._lib.lvkb200_frame_upload(self.stream._h, C.byref(c), space, optr, opitch, ospace))
        return VideoFrame(out, src.timestamp, self._ocl_format)

    def download_ocl_frame(self, src: VideoFrame, dst: ObsFrame):
        """FrameIngest::download_ocl_frame: packed frame -> the planes of `dst` (written in place)."""
        if dst.format != self._obs_format:
            raise LvkB200Error(_capi.ERR_INVALID, "dst->format == m_OBSFormat")
        ptr, pitch, h, w, ch, space = _buffer_info(src.data)
        c, dspace = dst.to_c()
        _capi.check(self._lib.lvkb200_frame_download(self.stream._h, ptr, pitch, w, h, src.format, space, C.byref(c), dspace))
        dst.timestamp = src.timestamp


@dataclass
class DeblockingFilterSettings:
    """lvk::DeblockingFilterSettings (Filters/DeblockingFilter.hpp:26-32)."""
    detection_levels: int = 3
    block_size: int = 16
    filter_size: int = 5
    filter_scaling: float = 4.0

    def to_c(self) -> _capi.DeblockSettings:
        return _capi.DeblockSettings(int(self.detection_levels), int(self.block_size), int(self.filter_size),
                                     float(self.filter_scaling))


class DeblockingFilter:
    """lvk::DeblockingFilter behind lvk::VideoFilter::apply (Filters/DeblockingFilter.hpp:34-59)."""

    def __init__(self, settings: DeblockingFilterSettings | None = None, device: int = 0, stream: "Stream | None" = None):
        self._settings = settings or DeblockingFilterSettings()
        self.stream = stream or Stream(None, device)
        self.alias = "Deblocking Filter"
        self.configure(self._settings)

    def settings(self) -> DeblockingFilterSettings:
        return self._settings

    def configure(self, settings: DeblockingFilterSettings):
        # DeblockingFilter::configure preconditions (DeblockingFilter.cpp:38-42)
        if not (settings.block_size > 0 and settings.filter_size >= 3 and settings.filter_size % 2 == 1
                and settings.detection_levels > 0 and settings.filter_scaling > 1.0):
            raise LvkB200Error(_capi.ERR_INVALID, "DeblockingFilter::configure precondition")
        self._settings = settings

    def filter_region(self, width: int, height: int):
        """DeblockingFilter::filter_region (:136-139) for a frame of this size: (x, y, w, h)."""
        bs = int(self._settings.block_size)
        return 0, 0, (width // bs) * bs, (height // bs) * bs

    def apply(self, frame: VideoFrame, output=None) -> VideoFrame:
        out = self.stream.deblock(frame.data, self._settings, frame.format, output)
        return VideoFrame(out, frame.timestamp, frame.format)


@dataclass
class ScalingFilterSettings:
    """lvk::ScalingFilterSettings (Filters/ScalingFilter.hpp:27-32)."""
    output_size: tuple = (1920, 1080)
    sharpness: float = 0.8
    yuv_input: bool = True

    def to_c(self) -> _capi.ScalingSettings:
        return _capi.ScalingSettings(int(self.output_size[0]), int(self.output_size[1]), float(self.sharpness),
                                     int(bool(self.yuv_input)))


class ScalingFilter:
    """lvk::ScalingFilter behind lvk::VideoFilter::apply (Filters/ScalingFilter.hpp:34-52, .cpp:28-59)."""

    def __init__(self, settings: "ScalingFilterSettings | None" = None, device: int = 0, stream: "Stream | None" = None):
        self.stream = stream or Stream(None, device)
        self.alias = "Scaling Filter"
        self.configure(settings or ScalingFilterSettings())

    def settings(self) -> ScalingFilterSettings:
        return self._settings

    def configure(self, settings: ScalingFilterSettings):
        # ScalingFilter::configure preconditions (ScalingFilter.cpp:43-45)
        if not (0.0 <= settings.sharpness <= 1.0 and settings.output_size[0] > 0 and settings.output_size[1] > 0):
            raise LvkB200Error(_capi.ERR_INVALID, "ScalingFilter::configure precondition")
        self._settings = settings

    def apply(self, frame: VideoFrame, output=None) -> VideoFrame:
        out = self.stream.scaling_filter(frame.data, self._settings, output)
        return VideoFrame(out, frame.timestamp, frame.format)  # output.timestamp = input.timestamp (:58)


class CompositeFilter:
    """lvk::CompositeFilter (Filters/CompositeFilter.cpp:58-88): applies its filters in order, each one's output
    being the next one's input.  The chain BASELINE config 5 names — DeblockingFilter -> StabilizationFilter — runs
    as ONE device pipeline: the deblocking stage is attached to the stabilizer's stream and works on the frame inside
    its device ring (no extra transfer, no intermediate buffer).  Other chains run filter by filter."""

    def __init__(self, filters):
        self.filters = list(filters)
        self.alias = "Composite Filter"
        self._fused = (len(self.filters) == 2 and isinstance(self.filters[0], DeblockingFilter)
                       and isinstance(self.filters[1], StabilizationFilter))
        if self._fused:
            self.filters[1].stream.set_deblocking(self.filters[0].settings())

    def apply(self, frame: VideoFrame, output=None) -> VideoFrame:
        if self._fused:
            return self.filters[1].apply(frame, output)
        for i, f in enumerate(self.filters):
            frame = f.apply(frame, output if i == len(self.filters) - 1 else None)
            if frame.empty():
                return frame
        return frame

    def stream_frames(self, frames, callback, outputs=None) -> int:
        """lvk::VideoFilter::stream (Filters/VideoFilter.cpp:62-209) for a chain: the fused Deblocking -> Stabilization
        chain runs pipelined on the stabilizer's device pipeline; any other chain is applied frame by frame (what the
        reference's filter thread does, VideoFilter.cpp:130-139): empty outputs (a filter that is still buffering) are
        skipped, a true return of `callback` terminates the stream.  Returns the number of outputs delivered."""
        if self._fused:
            return self.filters[1].stream_frames(frames, callback, outputs)
        delivered = 0
        for frame in frames:
            out = self.apply(frame)
            if out.empty():
                continue
            delivered += 1
            if callback(out):
                break
        return delivered


class StabilizationFilter:
    """lvk::StabilizationFilter behind lvk::VideoFilter::apply (Filters/StabilizationFilter.hpp:42-63,
    Filters/VideoFilter.hpp:32-61).  apply() returns a VideoFrame whose .data is None while the look-ahead
    queue fills (the reference releases `output`), otherwise the stabilized, delayed frame."""

    def __init__(self, settings: StabilizationFilterSettings | None = None, device: int = 0):
        self._settings = settings or StabilizationFilterSettings()
        self.stream = Stream(self._settings, device)
        self.alias = "Stabilization Filter"
        self.last_result = None

    def settings(self) -> StabilizationFilterSettings:
        return self._settings

    def configure(self, settings: StabilizationFilterSettings):
        self.stream.configure(settings)
        self._settings = settings

    def restart(self):
        self.stream.restart()

    def reset_context(self):
        self.stream.reset_context()

    def ready(self) -> bool:
        return self.stream.ready()

    def frame_delay(self) -> int:
        return self.stream.frame_delay()

    def stable_region(self, width: int, height: int):
        return self.stream.stable_region(width, height)

    def stream_frames(self, frames, callback, outputs=None) -> int:
        """lvk::VideoFilter::stream(input, callback, profile) (Filters/VideoFilter.cpp:62-209); also reachable as
        `filter.stream(frames, callback)` because the `stream` attribute is callable.  Pushes every frame of
        `frames` (a sequence of VideoFrame with HOST data) through the filter and hands each non-empty output to
        `callback(VideoFrame) -> bool` (a true return terminates the stream, VideoFilter.cpp:180-206; None/False continue).  The reference overlaps input,
        filtering and output with three threads; here the upload of frame t+1 and the download of output t-1 overlap
        the processing of frame t on separate CUDA streams.  `outputs`: optional list of >= 3 reusable host buffers
        (pinned memory makes the copies truly asynchronous).  Returns the number of outputs delivered."""
        return self.stream(frames, callback, outputs)  # Stream.__call__ (the attribute doubles as the method)

    def apply(self, frame: VideoFrame, output=None) -> VideoFrame:
        """lvk::VideoFilter::apply.  `output`: optional preallocated buffer (numpy or CUDA tensor) receiving the result.
        The returned frame is COMPLETE, as in the reference (a UMat access synchronises implicitly): for a CUDA output
        the library may hold the remap back until the next submit (lvkb200.h), so this mirror synchronises the stream
        before handing the buffer out.  Callers that want the overlap use `stream.submit` / `stream_frames` and order
        against `stream.sync()` / `event_record` themselves."""
        data = frame.data
        if output is None:
            output = np.empty_like(data) if isinstance(data, np.ndarray) else data.new_empty(data.shape)
        res = self.stream.submit(data, output, frame.format, frame.timestamp)
        self.last_result = res
        if not res.has_output:
            return VideoFrame(None, 0, UNKNOWN)
        if not isinstance(output, np.ndarray):
            self.stream.sync()
        return VideoFrame(output, int(res.out_timestamp), int(res.out_format))
