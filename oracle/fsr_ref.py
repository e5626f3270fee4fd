"""
ORACLE — TEST INFRASTRUCTURE ONLY.  ctypes binding of oracle/_ref/libfsrcl_ref_{strict,contract}.so: the reference's
OWN OpenCL kernel source (LiveVisionKit/Functions/OpenCL/Sources/FSR.cl) compiled for the CPU by
oracle/ref_build/build_ref.sh (an OpenCL-C shim + the launch code of Functions/Image.cpp:28-233).  This is the
reference itself, not a restatement; oracle/easu_ref.c and the CUDA kernels are checked against it.

  strict   : g++ -ffp-contract=off          (no multiply-add is fused)
  contract : g++ -ffp-contract=fast -mfma   (fused wherever gcc can; OpenCL C allows contraction by default)

The libraries are prebuilt in the build container (the GPU box has no /root/reference) and travel with the repo
snapshot; `available()` says whether they are there.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS: dict = {}
FLAVOURS = ("strict", "contract")

_u8p, _f32p, _f64p = (ctypes.POINTER(ctypes.c_uint8), ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_double))
_i = ctypes.c_int


def build(reference_root: str = "/root/reference") -> None:
    """Runs the recipe (a no-op when the reference tree is absent or the libraries are up to date)."""
    subprocess.check_call(["bash", os.path.join(_HERE, "ref_build", "build_ref.sh"), reference_root])


def _path(flavour: str) -> str:
    return os.path.join(_HERE, "_ref", f"libfsrcl_ref_{flavour}.so")


def available(flavour: str = "strict") -> bool:
    return os.path.exists(_path(flavour))


def lib(flavour: str = "strict"):
    if flavour not in _LIBS:
        if not available(flavour):
            build()
        L = ctypes.CDLL(_path(flavour))
        L.ref_easu_remap_homography.argtypes = [_u8p, _i, _i, _i, _u8p, _i, _f64p, _u8p, _i, _i]
        L.ref_easu_remap_map.argtypes = [_u8p, _i, _i, _i, _u8p, _i, _f32p, _i, _i, _i, _u8p, _i, _i]
        L.ref_easu_scale.argtypes = [_u8p, _i, _i, _i, _u8p, _i, _i, _i, _i, _i]
        L.ref_rcas.argtypes = [_u8p, _i, _i, _i, _u8p, _i, ctypes.c_float, _i]
        L.ref_rcas_kernel_sharpness.argtypes = [ctypes.c_float]
        L.ref_rcas_kernel_sharpness.restype = ctypes.c_float
        for f in (L.ref_easu_remap_homography, L.ref_easu_remap_map, L.ref_easu_scale, L.ref_rcas):
            f.restype = None
        _LIBS[flavour] = L
    return _LIBS[flavour]


def _p(a, ty):
    return a.ctypes.data_as(ty)


def _bg(background):
    return np.array([int(background[0]) & 255, int(background[1]) & 255, int(background[2]) & 255], dtype=np.uint8)


def remap_homography(src, t_inv, background=(255, 0, 255), yuv=False, flavour="strict", threads=0):
    """lvk::remap(src, dst, homography, background, inverted=true) — Image.cpp:85-151 -> easu_remap_homography."""
    src = np.ascontiguousarray(src, dtype=np.uint8)
    dst = np.zeros_like(src)
    t = np.ascontiguousarray(t_inv, dtype=np.float64).reshape(9)
    bg = _bg(background)
    lib(flavour).ref_easu_remap_homography(_p(src, _u8p), src.strides[0], src.shape[0], src.shape[1], _p(dst, _u8p),
                                           dst.strides[0], _p(t, _f64p), _p(bg, _u8p), int(yuv), threads)
    return dst


def remap_map(src, offset_map, background=(255, 0, 255), yuv=False, flavour="strict", threads=0):
    """lvk::remap(src, dst, offset_map, background) — Image.cpp:28-81 -> easu_remap."""
    src = np.ascontiguousarray(src, dtype=np.uint8)
    m = np.ascontiguousarray(offset_map, dtype=np.float32)
    rows, cols = m.shape[:2]
    dst = np.zeros((rows, cols, 3), dtype=np.uint8)
    bg = _bg(background)
    lib(flavour).ref_easu_remap_map(_p(src, _u8p), src.strides[0], src.shape[0], src.shape[1], _p(dst, _u8p),
                                    dst.strides[0], _p(m, _f32p), m.strides[0], rows, cols, _p(bg, _u8p), int(yuv), threads)
    return dst


def upscale(src, size, yuv=False, flavour="strict", threads=0):
    """lvk::upscale — Image.cpp:155-201 -> easu_scale.  size = (width, height)."""
    src = np.ascontiguousarray(src, dtype=np.uint8)
    w, h = int(size[0]), int(size[1])
    dst = np.zeros((h, w, 3), dtype=np.uint8)
    lib(flavour).ref_easu_scale(_p(src, _u8p), src.strides[0], src.shape[0], src.shape[1], _p(dst, _u8p), dst.strides[0],
                                h, w, int(yuv), threads)
    return dst


def sharpen(src, sharpness, flavour="strict", threads=0):
    """lvk::sharpen — Image.cpp:205-233 -> rcas (out of place)."""
    src = np.ascontiguousarray(src, dtype=np.uint8)
    dst = np.zeros_like(src)
    L = lib(flavour)
    ks = L.ref_rcas_kernel_sharpness(float(np.float32(sharpness)))
    L.ref_rcas(_p(src, _u8p), src.strides[0], src.shape[0], src.shape[1], _p(dst, _u8p), dst.strides[0], ks, threads)
    return dst


def lsb_histogram(a, b, bins=4):
    """[#bytes with |a-b| = 0, 1, 2, >= 3] — the form every parity artifact under profiles/ uses."""
    d = np.abs(a.astype(np.int16) - b.astype(np.int16)).ravel()
    h = np.bincount(np.minimum(d, bins - 1), minlength=bins)
    return [int(v) for v in h]
