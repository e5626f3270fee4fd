"""
Deterministic synthetic clips for tests and bench.  Test/bench infrastructure —
not part of the product path.  The reference ships no test clips; these stand in for BASELINE.json's
"synthetic pan clip" (config 0) and "synthetic hand-shake sequence" (configs 1-4).

Base texture (seed 1234): 3 octaves of Gaussian-filtered uniform noise + random rectangles and discs
(FAST-rich, LK-friendly), BGR 8UC3 with per-channel gain (1.0, 0.9, 0.8).  Camera path: 'pan'
(constant velocity) or 'shake' (sum of sinusoids + AR(1) jitter, small rotation and zoom).  Frames are
rendered with cv2.warpAffine(INTER_LINEAR, BORDER_REFLECT) from a canvas larger than the frame.
"""
from __future__ import annotations

import math

import cv2
import numpy as np

RESOLUTIONS = {"270p": (480, 270), "720p": (1280, 720), "1080p": (1920, 1080), "4k": (3840, 2160)}


def make_canvas(width: int, height: int, seed: int = 1234) -> np.ndarray:
    rng = np.random.default_rng(seed)
    m = int(round(0.08 * max(width, height)))
    cw, ch = width + 2 * m, height + 2 * m
    tex = np.zeros((ch, cw), dtype=np.float32)
    for sigma, weight in ((1.5, 0.5), (6.0, 0.3), (24.0, 0.2)):
        noise = rng.random((ch, cw), dtype=np.float32)
        blur = cv2.GaussianBlur(noise, (0, 0), sigma)
        blur = (blur - blur.mean()) / (blur.std() + 1e-6)
        tex += weight * blur
    tex = (tex - tex.min()) / (tex.max() - tex.min())
    tex = 16.0 + tex * (235.0 - 16.0)
    scale = (width * height) / 2.0e6
    for _ in range(int(400 * scale)):
        w, h = rng.integers(8, 65, size=2)
        x, y = rng.integers(0, cw - w), rng.integers(0, ch - h)
        tex[y:y + h, x:x + w] += rng.uniform(-60, 60)
    for _ in range(int(200 * scale)):
        r = int(rng.integers(4, 33))
        x, y = int(rng.integers(r, cw - r)), int(rng.integers(r, ch - r))
        delta = float(rng.uniform(-60, 60))
        yy, xx = np.ogrid[-r:r + 1, -r:r + 1]
        mask = (xx * xx + yy * yy) <= r * r
        tex[y - r:y + r + 1, x - r:x + r + 1][mask] += delta
    tex = np.clip(tex, 0, 255)
    canvas = np.stack([tex * 1.0, tex * 0.9, tex * 0.8], axis=-1)
    return np.clip(np.rint(canvas), 0, 255).astype(np.uint8)


def camera_path(kind: str, n: int, width: int, fps: float, seed: int = 42):
    """Returns arrays (tx, ty, rot_deg, scale) of length n."""
    t = np.arange(n, dtype=np.float64) / fps
    if kind == "pan":
        return 2.0 * np.arange(n), 0.5 * np.arange(n), np.zeros(n), np.ones(n)
    rng = np.random.default_rng(seed)
    freqs = (1.1, 2.3, 4.7, 7.9)
    amps = np.array((0.6, 0.4, 0.25, 0.1)) * 0.01 * width
    tx = np.zeros(n)
    ty = np.zeros(n)
    for f, a in zip(freqs, amps):
        tx += a * np.sin(2 * math.pi * f * t + rng.uniform(0, 2 * math.pi))
        ty += 0.7 * a * np.sin(2 * math.pi * f * t + rng.uniform(0, 2 * math.pi))
    jx = np.zeros(n)
    jy = np.zeros(n)
    sig = 0.0005 * width
    for i in range(1, n):
        jx[i] = 0.9 * jx[i - 1] + rng.normal(0, sig)
        jy[i] = 0.9 * jy[i - 1] + rng.normal(0, sig)
    rot = 0.4 * np.sin(2 * math.pi * 1.7 * t + rng.uniform(0, 2 * math.pi)) + rng.normal(0, 0.01, n)
    scale = 1.0 + 0.003 * np.sin(2 * math.pi * 0.9 * t + rng.uniform(0, 2 * math.pi))
    return tx + jx, ty + jy, rot, scale


class Clip:
    """Lazy frame renderer: clip[i] -> HxWx3 uint8 BGR."""

    def __init__(self, resolution="1080p", kind="shake", frames=330, fps=60.0, seed=42, canvas_seed=1234):
        self.width, self.height = RESOLUTIONS[resolution] if isinstance(resolution, str) else resolution
        self.kind, self.n, self.fps, self.seed = kind, frames, fps, seed
        self.canvas = make_canvas(self.width, self.height, canvas_seed)
        self.occluder = None
        if kind == "occluder":
            # SURVEY 8(d) "occluder variant, seed 7": the hand-shake clip with an independently moving textured block
            # covering 15 % of the frame (0.45 W x 1/3 H) - the features on it are RANSAC outliers
            self.seed = seed = 7
            bw, bh = int(round(0.45 * self.width)), int(round(self.height / 3.0))
            tex = make_canvas(bw, bh, 7)
            y0, x0 = (tex.shape[0] - bh) // 2, (tex.shape[1] - bw) // 2
            self.occluder = np.ascontiguousarray(tex[y0:y0 + bh, x0:x0 + bw, ::-1])  # other channel gains than the scene
            t = np.arange(frames, dtype=np.float64) / fps
            self.occ_x = 0.5 * (self.width - bw) + 0.25 * (self.width - bw) * np.sin(2 * math.pi * 0.35 * t)
            self.occ_y = 0.5 * (self.height - bh) + 0.30 * (self.height - bh) * np.sin(2 * math.pi * 0.23 * t + 1.0)
        self.tx, self.ty, self.rot, self.scale = camera_path("shake" if kind == "occluder" else kind, frames, self.width, fps, seed)
        if kind == "pan":  # keep the pan inside the canvas margin by wrapping the ramp into a triangle wave
            m = 0.08 * max(self.width, self.height) * 0.9
            self.tx = np.abs(((self.tx + m) % (4 * m)) - 2 * m) - m
            self.ty = np.abs(((self.ty + m) % (4 * m)) - 2 * m) - m

    def __len__(self):
        return self.n

    def matrix(self, i: int) -> np.ndarray:
        """2x3 affine mapping output pixel -> canvas pixel (used with WARP_INVERSE_MAP)."""
        ch, cw = self.canvas.shape[:2]
        cx, cy = self.width / 2.0, self.height / 2.0
        a = math.radians(self.rot[i])
        s = self.scale[i]
        c, sn = math.cos(a) * s, math.sin(a) * s
        ox = (cw - self.width) / 2.0 + self.tx[i]
        oy = (ch - self.height) / 2.0 + self.ty[i]
        # p_canvas = R*(p - c) + c + o
        return np.array([[c, -sn, cx - c * cx + sn * cy + ox],
                         [sn, c, cy - sn * cx - c * cy + oy]], dtype=np.float64)

    def __getitem__(self, i: int) -> np.ndarray:
        frame = cv2.warpAffine(self.canvas, self.matrix(i), (self.width, self.height),
                               flags=cv2.INTER_LINEAR | cv2.WARP_INVERSE_MAP, borderMode=cv2.BORDER_REFLECT)
        if self.occluder is not None:
            m = np.array([[1, 0, self.occ_x[i]], [0, 1, self.occ_y[i]]], dtype=np.float64)
            bh, bw = self.occluder.shape[:2]
            block = cv2.warpAffine(self.occluder, m, (self.width, self.height), flags=cv2.INTER_LINEAR)
            mask = cv2.warpAffine(np.full((bh, bw), 255, np.uint8), m, (self.width, self.height), flags=cv2.INTER_NEAREST)
            frame[mask > 0] = block[mask > 0]
        return frame

    def occluder_rect(self, i: int):
        """(x, y, w, h) of the moving block in frame i (occluder clips only)."""
        bh, bw = self.occluder.shape[:2]
        return float(self.occ_x[i]), float(self.occ_y[i]), bw, bh

    def frames(self, start=0, stop=None):
        for i in range(start, self.n if stop is None else stop):
            yield self[i]
