"""
ORACLE — TEST INFRASTRUCTURE ONLY.  A second, independent restatement of the reference's EASU remap, written directly
from the OpenCL source (LiveVisionKit/Functions/OpenCL/Sources/FSR.cl:55-68 APrxLo*, :98-126 easu_tap, :131-176
easu_accumulate, :181-318 easu, :407-452 easu_remap_homography) as vectorised NumPy float32 — no multiply-add is fused
(NumPy has no FMA), i.e. the STRICT arithmetic.  It shares no code with oracle/easu_ref.c (scalar C, explicit FMA rule)
and none with oracle/ref_build (the reference's kernel text compiled through a shim): tests/test_fsr_ref_cpu.py checks
that it reproduces the reference's strict build bit for bit and brackets easu_ref.c.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32


def _as_float(u):
    return u.astype(np.uint32).view(np.float32)


def _as_uint(f):
    return f.astype(np.float32).view(np.uint32)


def aprx_lo_rcp(a):  # FSR.cl:65
    return _as_float(np.uint32(0x7ef07ebb) - _as_uint(a))


def aprx_lo_rsq(a):  # FSR.cl:60
    return _as_float(np.uint32(0x5f347d74) - (_as_uint(a) >> np.uint32(1)))


def _cl_max(a, b):  # OpenCL max(x, y) = y if x < y else x
    return np.where(a < b, b, a)


def _cl_min(a, b):  # OpenCL min(x, y) = y if y < x else x
    return np.where(b < a, b, a)


def _saturate(x):  # FSR.cl:79: fmax(0, fmin(1, x))
    return np.fmax(f32(0), np.fmin(f32(1), x))


def _accumulate(dirx, diry, length, w, lA, lB, lC, lD, lE):  # FSR.cl:131-176 (w = the corner's bilinear weight)
    dc, cb = lD - lC, lC - lB
    lenx = aprx_lo_rcp(_cl_max(np.abs(dc), np.abs(cb)))
    dir_x = lD - lB
    dirx = dirx + dir_x * w
    lenx = _saturate(np.abs(dir_x) * lenx)
    lenx = lenx * lenx
    length = length + lenx * w
    ec, ca = lE - lC, lC - lA
    leny = aprx_lo_rcp(_cl_max(np.abs(ec), np.abs(ca)))
    dir_y = lE - lA
    diry = diry + dir_y * w
    leny = _saturate(np.abs(dir_y) * leny)
    leny = leny * leny
    length = length + leny * w
    return dirx, diry, length


def remap_homography(src: np.ndarray, t_inv: np.ndarray, background=(255, 0, 255), yuv: bool = False) -> np.ndarray:
    """lvk::remap(src, dst, homography, background, inverted=true): Image.cpp:85-151 + FSR.cl:407-452, strict float32."""
    with np.errstate(all="ignore"):
        src = np.ascontiguousarray(src, dtype=np.uint8)
        rows, cols = src.shape[:2]
        t = np.asarray(t_inv, dtype=np.float64).reshape(3, 3).astype(np.float32)  # cv::Vec4f(t.at<double>(..))
        fy, fx = np.meshgrid(np.arange(rows, dtype=np.float32), np.arange(cols, dtype=np.float32), indexing="ij")
        dz = f32(1) / (t[2, 0] * fx + t[2, 1] * fy + t[2, 2])
        offx = (t[0, 0] * fx + t[0, 1] * fy + t[0, 2]) * dz - fx
        offy = (t[1, 0] * fx + t[1, 1] * fy + t[1, 2]) * dz - fy
        subx, suby = fx + offx, fy + offy
        sx, sy = np.trunc(subx).astype(np.int64), np.trunc(suby).astype(np.int64)  # convert_int2_rtz
        sx = np.where(np.isfinite(subx), sx, -(1 << 30))
        sy = np.where(np.isfinite(suby), sy, -(1 << 30))
        ppx, ppy = subx - np.floor(subx), suby - np.floor(suby)
        border = (sx < 1) | (sy < 1) | (sx >= cols - 4) | (sy >= rows - 4)
        in_src = (sx >= 0) & (sx < cols) & (sy >= 0) & (sy < rows)
        out = np.empty_like(src)
        out[...] = np.array([int(c) & 255 for c in background[:3]], dtype=np.uint8)
        near = border & in_src
        out[near] = src[sy[near], sx[near]]
        e = ~border
        if not e.any():
            return out
        x, y, px, py = sx[e], sy[e], ppx[e].astype(np.float32), ppy[e].astype(np.float32)
        norm = f32(0.00392156862)

        def tex(dx, dy):
            return src[y + dy, x + dx].astype(np.float32) * norm  # (n, 3)

        b, c = tex(0, -1), tex(1, -1)
        ee, f, g, h = tex(-1, 0), tex(0, 0), tex(1, 0), tex(2, 0)
        i, j, k, l = tex(-1, 1), tex(0, 1), tex(1, 1), tex(2, 1)
        n, o = tex(0, 2), tex(1, 2)

        def luma(p):  # FSR.cl:229-241: WITHOUT YUV_INPUT luma = channel 0, WITH it B*0.5 + (R*0.5 + G)
            return (p[:, 2] * f32(0.5) + (p[:, 0] * f32(0.5) + p[:, 1])) if yuv else p[:, 0]

        bL, cL, eL, fL, gL, hL, iL, jL, kL, lL, nL, oL = map(luma, (b, c, ee, f, g, h, i, j, k, l, n, o))
        one = f32(1)
        zero = np.zeros_like(px)
        dirx, diry, length = zero, zero, zero
        dirx, diry, length = _accumulate(dirx, diry, length, (one - px) * (one - py), bL, eL, fL, gL, jL)
        dirx, diry, length = _accumulate(dirx, diry, length, px * (one - py), cL, fL, gL, hL, kL)
        dirx, diry, length = _accumulate(dirx, diry, length, (one - px) * py, fL, iL, jL, kL, nL)
        dirx, diry, length = _accumulate(dirx, diry, length, px * py, gL, jL, kL, lL, oL)
        dir2x, dir2y = dirx * dirx, diry * diry
        dir_r = dir2x + dir2y
        zro = dir_r < f32(1.0 / 32768.0)
        dir_r = np.where(zro, one, aprx_lo_rsq(dir_r))
        dirx = np.where(zro, one, dirx) * dir_r
        diry = diry * dir_r
        length = length * f32(0.5)
        length = length * length
        stretch = (dirx * dirx + diry * diry) * aprx_lo_rcp(_cl_max(np.abs(dirx), np.abs(diry)))
        len2x = one + (stretch - one) * length
        len2y = one + f32(-0.5) * length
        lob = f32(0.5) + f32((1.0 / 4.0 - 0.04) - 0.5) * length
        clp = aprx_lo_rcp(lob)
        mi4 = _cl_min(f, _cl_min(g, _cl_min(j, k)))
        ma4 = _cl_max(f, _cl_max(g, _cl_max(j, k)))
        aC = np.zeros_like(f)
        aW = np.zeros_like(px)

        def tap(offx_, offy_, colour):  # FSR.cl:98-126
            nonlocal aC, aW
            ox, oy = f32(offx_) - px, f32(offy_) - py
            vx = (ox * dirx + oy * diry) * len2x
            vy = (ox * (-diry) + oy * dirx) * len2y
            d2 = _cl_min(vx * vx + vy * vy, clp)
            wA = lob * d2 - one
            wB = f32(2.0 / 5.0) * d2 - one
            wA = wA * wA
            wB = f32(25.0 / 16.0) * (wB * wB) - f32(25.0 / 16.0 - 1.0)
            w = wB * wA
            aC = aC + colour * w[:, None]
            aW = aW + w

        for ox_, oy_, col in ((0, -1, b), (1, -1, c), (-1, 1, i), (0, 1, j), (0, 0, f), (-1, 0, ee), (1, 1, k), (2, 1, l),
                              (2, 0, h), (1, 0, g), (0, 2, n), (1, 2, o)):
            tap(ox_, oy_, col)
        fpx = _cl_min(ma4, _cl_max(mi4, aC * (one / aW)[:, None]))
        out[e] = np.trunc(fpx * f32(255.0)).astype(np.uint8)  # convert_uchar3
        return out
