"""ScalingFilter oracle pinning (CPU): the C restatement of easu_scale / rcas (oracle/easu_ref.c, FSR.cl:326-358 and
:460-535) against an independent NumPy restatement of RCAS, against the already-pinned homography remap for EASU, and
against first-principles properties.  The reference holds no vectors for this filter and cannot be built here
(no OpenCL runtime): parity unpinned, as for the rest of the path."""
import numpy as np
import pytest


def _textured(h, w, seed=0):
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    img[h // 8: h // 4, w // 8: w // 3] = 0      # all-zero ring: 0 * inf in the RCAS limiter
    img[h // 2: h // 2 + h // 8, w // 2: w // 2 + w // 5] = 255  # all-one ring: 4*mn4 - 4 == 0
    yy, xx = np.mgrid[0:h, 0:w]
    smooth = (96 + 64 * np.sin(xx / 9.0) * np.cos(yy / 7.0)).astype(np.uint8)
    img[:, : w // 4] = smooth[:, : w // 4, None]
    return img


def _fma(a, b, c):
    """fmaf on float32 arrays through float64 (the product is exact; the sum is rounded twice, 53 then 24 bits — it
    can differ from a true FMA in rare double-rounding cases, hence the tolerance below)."""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def _rcas_numpy(src, sharp):
    f32 = np.float32
    norm = f32(0.00392156862)
    p = src.astype(np.float32) * norm
    b, h, d, f, e = p[:-2, 1:-1], p[2:, 1:-1], p[1:-1, :-2], p[1:-1, 2:], p[1:-1, 1:-1]
    with np.errstate(all="ignore"):
        mn4 = np.fmin(b, np.fmin(d, np.fmin(f, h)))
        mx4 = np.fmax(b, np.fmax(d, np.fmax(f, h)))
        hit_min = np.fmin(mn4, e) * (f32(1.0) / (f32(4.0) * mx4))
        hit_max = (f32(1.0) - np.fmax(mx4, e)) * (f32(1.0) / _fma(np.full_like(mn4, 4.0), mn4, np.full_like(mn4, -4.0)))
        lobe_c = np.fmax(-hit_min, hit_max)
        lobe = np.fmax(lobe_c[..., 2], np.fmax(lobe_c[..., 1], lobe_c[..., 0]))
        lobe = (np.fmin(np.fmax(lobe, f32(-0.1875)), f32(0.0)) * f32(sharp)).astype(np.float32)
        a = _fma(np.full_like(lobe, 4.0), lobe, np.full_like(lobe, 1.0))
        bb = (np.uint32(0x7ef19fff) - a.view(np.uint32)).view(np.float32)
        rcp = bb * _fma(-bb, a, np.full_like(a, 2.0))
        s = ((b + d) + h) + f
        v = _fma(s, lobe[..., None], e) * rcp[..., None]
        out = np.clip((v * f32(255.0)).astype(np.int64), 0, 255).astype(np.uint8)
    res = src.copy()
    res[1:-1, 1:-1] = out
    return res


@pytest.mark.parametrize("sharpness", [0.0, 0.8, 1.0])
def test_rcas_matches_numpy_restatement(oracle, sharpness):
    src = _textured(97, 131, 3)
    got = oracle.sharpen(src, sharpness)
    ref = _rcas_numpy(src, oracle.rcas_kernel_sharpness(sharpness))
    d = np.abs(got.astype(int) - ref.astype(int))
    assert d.max() <= 1 and (d > 0).mean() <= 1e-4
    # border pixels are copied (FSR.cl:478-484)
    assert (got[0] == src[0]).all() and (got[-1] == src[-1]).all() and (got[:, 0] == src[:, 0]).all() and (got[:, -1] == src[:, -1]).all()


def test_rcas_properties(oracle):
    assert float(oracle.rcas_kernel_sharpness(1.0)) == 1.0 and float(oracle.rcas_kernel_sharpness(0.0)) == 0.25
    flat = np.full((20, 30, 3), 137, np.uint8)
    assert np.abs(oracle.sharpen(flat, 1.0).astype(int) - 137).max() <= 1  # truncating conversion of 137 * (1/255) * 255
    src = _textured(64, 80, 5)
    weak, strong = oracle.sharpen(src, 0.0).astype(int), oracle.sharpen(src, 1.0).astype(int)
    # a stronger setting moves pixels further from the input (on average), never beyond [0, 255]
    assert np.abs(strong - src).mean() > np.abs(weak - src).mean() > 0
    # thread-count independence of the row-parallel driver
    assert (oracle.sharpen(src, 0.8, threads=1) == oracle.sharpen(src, 0.8, threads=7)).all()
    # tiny images are all border
    for shape in ((1, 1), (2, 5), (3, 2)):
        t = _textured(8, 8)[: shape[0], : shape[1]]
        assert (oracle.sharpen(t, 0.8) == t).all()


def test_upscale_agrees_with_the_remap_kernel(oracle):
    """easu_scale is easu_remap_homography with a pure scaling transform, up to the float rounding of the source
    position (x * r versus x + (x * r - x)): nearly every pixel must agree exactly."""
    src = _textured(90, 120, 7)
    for (dw, dh) in ((240, 180), (180, 135), (200, 173)):
        up = oracle.upscale(src, (dw, dh), yuv=False)
        rx, ry = np.float32(120) / np.float32(dw), np.float32(90) / np.float32(dh)
        big = np.zeros((dh, dw, 3), np.uint8)
        big[:90, :120] = src
        # the remap works on equal-size images: embed the source, restrict the comparison to where its border logic
        # (taps inside the 120x90 source) coincides with easu_scale's
        ref = oracle.remap_homography(big, np.diag([float(rx), float(ry), 1.0]), (0, 0, 0), False)
        sx = (np.arange(dw, dtype=np.float32) * rx).astype(int)
        sy = (np.arange(dh, dtype=np.float32) * ry).astype(int)
        inner = ((sy >= 1) & (sy < 90 - 4))[:, None] & ((sx >= 1) & (sx < 120 - 4))[None, :]
        same = (up == ref).all(axis=2)
        assert same[inner].mean() > 0.995
        assert np.abs(up.astype(int) - ref.astype(int))[inner].max() <= 2
        # border band: nearest-neighbour copy of src[sy, sx] (FSR.cl:341-351)
        outer = ~inner
        nn = src[sy][:, sx]
        assert (up[outer] == nn[outer]).all()


def test_upscale_properties(oracle):
    src = _textured(50, 70, 9)
    assert (oracle.upscale(src, (70, 50)) == src).all()  # size == src.size(): copy (Image.cpp:162-166)
    flat = np.full((40, 40, 3), 91, np.uint8)
    assert np.abs(oracle.upscale(flat, (100, 90)).astype(int) - 91).max() <= 1
    up = oracle.upscale(src, (140, 100), yuv=True)
    assert up.shape == (100, 140, 3)
    # EASU clamps to the min/max of the 2x2 neighbourhood: no new extremes
    assert up.min() >= src.min() and up.max() <= src.max()
    assert (oracle.upscale(src, (140, 100), yuv=True, threads=1) == up).all()
    # the yuv flag changes the luma the edge direction is estimated from
    assert (oracle.upscale(src, (140, 100), yuv=False) != up).any()
    f = oracle.ScalingFilter(oracle.ScalingFilterSettings((140, 100), 0.8, True))
    assert (f.apply(src) == oracle.sharpen(up, 0.8)).all()
    with pytest.raises(AssertionError):
        oracle.upscale(src, (60, 50))  # LVK_ASSERT(size >= src) — Image.cpp:157


def test_oracle_reproduces_the_committed_golden(oracle):
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "scaling_golden.npz"), allow_pickle=False)
    size = (int(g["size"][0]), int(g["size"][1]))
    assert (oracle.upscale(g["src"], size, False) == g["up_bgr"]).all()
    assert (oracle.upscale(g["src"], size, True) == g["up_yuv"]).all()
    assert (oracle.sharpen(g["src"], 0.8) == g["sharp_08"]).all()
    assert (oracle.sharpen(g["src"], 0.0) == g["sharp_00"]).all()
    assert (oracle.sharpen(g["src"], 1.0) == g["sharp_10"]).all()
    assert (oracle.ScalingFilter(oracle.ScalingFilterSettings(size, 0.8, True)).apply(g["src"]) == g["filter_yuv_08"]).all()
