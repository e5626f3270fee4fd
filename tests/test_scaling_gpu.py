"""ScalingFilter parity on the GPU: lvk::upscale (EASU kernel, MODE 2 of k_easu_remap) and
lvk::sharpen (k_rcas) through the C-ABI against the CPU restatement of FSR.cl (oracle/easu_ref.c).  Contract for
warped / filtered pixels: <= 1 LSB; the kernels share the oracle's explicit FMA placement, so the tests demand
bit-exact bytes."""
import numpy as np
import pytest

# bit-exact comparisons with the oracle restatement: these modules run the EXACT arithmetic build of the EASU kernels
# (tests that exercise the default contract build say so and switch it back on)
pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("exact_build")]


def _textured(h, w, seed=0):
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    img[h // 8: h // 4, w // 8: w // 3] = 0
    img[h // 2: h // 2 + h // 8, w // 2: w // 2 + w // 5] = 255
    yy, xx = np.mgrid[0:h, 0:w]
    smooth = (96 + 64 * np.sin(xx / 9.0) * np.cos(yy / 7.0)).astype(np.uint8)
    img[:, : w // 4] = smooth[:, : w // 4, None]
    return img


def _clip_frame(size):
    from tools.synth import Clip
    return Clip(size, "shake", frames=2, seed=21)[1]


@pytest.mark.parametrize("src_wh,dst_wh", [((480, 270), (960, 540)), ((1280, 720), (1920, 1080)), ((640, 360), (1920, 1080)),
                                           ((333, 217), (500, 400)), ((1920, 1080), (3840, 2160)), ((37, 19), (38, 19)),
                                           ((8, 8), (64, 48))])
def test_upscale_bit_exact(gpu_stream, oracle, src_wh, dst_wh):
    src = _textured(src_wh[1], src_wh[0], 1) if src_wh[0] < 1000 else _clip_frame("1080p" if src_wh[0] == 1920 else "720p")
    assert src.shape[:2] == (src_wh[1], src_wh[0])
    for yuv in (False, True):
        got = gpu_stream.upscale(src, dst_wh, yuv)
        ref = oracle.upscale(src, dst_wh, yuv)
        d = np.abs(got.astype(int) - ref.astype(int))
        print(f"upscale {src_wh}->{dst_wh} yuv={yuv}: {int((d > 0).sum())} differing bytes, max {int(d.max())}")
        assert (got == ref).all()


def test_upscale_same_size_is_a_copy_and_preconditions(gpu_stream):
    import livevisionkit_b200 as L
    src = _textured(60, 80, 2)
    assert (gpu_stream.upscale(src, (80, 60)) == src).all()
    with pytest.raises(L.LvkB200Error):
        gpu_stream.upscale(src, (79, 60))  # LVK_ASSERT(size >= src) — Image.cpp:157
    with pytest.raises(L.LvkB200Error):
        gpu_stream.sharpen(src, 1.5)       # LVK_ASSERT_01 — Image.cpp:209
    with pytest.raises(L.LvkB200Error):
        L.ScalingFilter(L.ScalingFilterSettings((0, 10), 0.5))


@pytest.mark.parametrize("wh", [(1920, 1080), (963, 541), (128, 16), (129, 17), (127, 15), (5, 4), (3, 3), (2, 7), (1, 1),
                                (3840, 2160)])
@pytest.mark.parametrize("sharpness", [0.8, 0.0, 1.0])
def test_sharpen_bit_exact(gpu_stream, oracle, wh, sharpness):
    if wh[0] >= 3840 and sharpness != 0.8:
        pytest.skip("one 4K case is enough")
    src = _clip_frame("1080p") if wh == (1920, 1080) else (_clip_frame("4k") if wh[0] == 3840 else _textured(wh[1], wh[0], 4))
    got = gpu_stream.sharpen(src, sharpness)
    ref = oracle.sharpen(src, sharpness)
    d = np.abs(got.astype(int) - ref.astype(int))
    print(f"sharpen {wh} s={sharpness}: {int((d > 0).sum())} differing bytes, max {int(d.max())}; "
          f"filter changed {100 * float((ref != src).mean()):.1f}% of the bytes")
    assert (got == ref).all()


def test_sharpen_device_buffers_unaligned_and_in_place(gpu_stream, oracle):
    torch = pytest.importorskip("torch")
    src = _textured(141, 203, 6)
    ref = oracle.sharpen(src, 0.8)
    dev = torch.from_numpy(src).cuda()
    out = torch.empty_like(dev)
    gpu_stream.sharpen(dev, 0.8, out)
    gpu_stream.sync()
    assert (out.cpu().numpy() == ref).all()
    # rows that are neither 16-byte aligned nor 16-byte pitched: a window of a larger device image
    big = torch.zeros((150, 260, 3), dtype=torch.uint8, device="cuda")
    big[5:146, 7:210] = dev
    win = big[5:146, 7:210]
    out2 = torch.zeros((150, 260, 3), dtype=torch.uint8, device="cuda")
    gpu_stream.sharpen(win, 0.8, out2[3:144, 11:214])
    gpu_stream.sync()
    assert (out2[3:144, 11:214].cpu().numpy() == ref).all()
    assert int(out2.sum()) == int(ref.astype(np.int64).sum())  # nothing written outside the window
    # in place (what ScalingFilter.cpp:57 asks for): every tap still reads the unsharpened frame
    gpu_stream.sharpen(dev, 0.8, dev)
    gpu_stream.sync()
    assert (dev.cpu().numpy() == ref).all()


@pytest.mark.parametrize("src_size,out_size,yuv", [("720p", (1920, 1080), True), ((960, 540), (1920, 1080), False),
                                                   ("1080p", (3840, 2160), True)])
def test_scaling_filter_vs_oracle(gpu_stream, oracle, src_size, out_size, yuv):
    import livevisionkit_b200 as L
    src = _clip_frame(src_size)
    ref = oracle.ScalingFilter(oracle.ScalingFilterSettings(out_size, 0.8, yuv)).apply(src)
    flt = L.ScalingFilter(L.ScalingFilterSettings(out_size, 0.8, yuv), stream=gpu_stream)
    got = flt.apply(L.VideoFrame(src, 123, L.YUV if yuv else L.BGR))
    assert got.timestamp == 123 and got.data.shape == ref.shape
    assert (got.data == ref).all()
    # device in, device out
    torch = pytest.importorskip("torch")
    dev = torch.from_numpy(src).cuda()
    out = torch.empty((out_size[1], out_size[0], 3), dtype=torch.uint8, device="cuda")
    gpu_stream.scaling_filter(dev, flt.settings(), out)
    gpu_stream.sync()
    assert (out.cpu().numpy() == ref).all()


def test_sharpen_every_ring_level(gpu_stream, oracle):
    """k_rcas replaces the limiter's two IEEE divisions by the bare MUFU.RCP + Newton-Raphson sequence, valid on the
    denominators' known domain (fsr.cu: rcp_limiter): exercise every denominator value — rings whose max and whose
    min are each of the 256 levels, including the all-0 and all-255 rings (0 * inf) — with arbitrary centres."""
    rng = np.random.default_rng(12)
    img = np.zeros((16 * 6, 16 * 6 * 3, 3), np.uint8)
    for k in range(256):
        y, x = (k // 16) * 6, (k % 16) * 6
        img[y:y + 6, x:x + 6] = k                                           # ring max == ring min == k
        img[y:y + 6, 96 + x:96 + x + 6] = rng.integers(0, k + 1, (6, 6, 3))  # ring max <= k, often == k
        img[y:y + 6, 192 + x:192 + x + 6] = rng.integers(k, 256, (6, 6, 3))  # ring min >= k
        img[y + 2, x + 2] = rng.integers(0, 256, 3)                          # arbitrary centres
        img[y + 3, x + 4] = rng.integers(0, 256, 3)
    for sharpness in (1.0, 0.37):
        got, ref = gpu_stream.sharpen(img, sharpness), oracle.sharpen(img, sharpness)
        assert (got == ref).all(), f"{int((got != ref).sum())} bytes differ"


@pytest.mark.parametrize("src_wh,dst_wh", [((480, 270), (960, 540)), ((1280, 720), (1920, 1080)), ((333, 217), (500, 400)),
                                           ((1920, 1080), (3840, 2160))])
def test_upscale_contract_build_vs_reference(gpu_stream, oracle, src_wh, dst_wh):
    """The DEFAULT (contract) build of lvk::upscale against the reference's own easu_scale compiled for the CPU
    (oracle/_ref) and against the restatement: max 1 LSB vs the contract reference, texel flips excepted vs strict."""
    import livevisionkit_b200 as L
    from oracle import fsr_ref as R
    src = _textured(src_wh[1], src_wh[0], 1) if src_wh[0] < 1000 else _clip_frame("1080p" if src_wh[0] == 1920 else "720p")
    L.set_remap_exact(False)
    try:
        got = gpu_stream.upscale(src, dst_wh, False)
    finally:
        L.set_remap_exact(True)  # the module-wide fixture restores the default afterwards
    n = got.size
    ho = R.lsb_histogram(got, oracle.upscale(src, dst_wh, False))
    hc = R.lsb_histogram(got, R.upscale(src, dst_wh, False, "contract"))
    hs = R.lsb_histogram(got, R.upscale(src, dst_wh, False, "strict"))
    print(f"upscale {src_wh}->{dst_wh} contract build: vs restatement {ho}, vs reference contract {hc}, strict {hs}")
    assert ho[2] == 0 and ho[3] == 0 and ho[0] >= 0.999 * n
    assert hc[2] == 0 and hc[3] == 0 and hc[0] >= 0.999 * n
    assert hs[2] + hs[3] < 1e-4 * n and hs[0] >= 0.999 * n


@pytest.mark.parametrize("sharpness", [0.0, 0.8, 1.0])
def test_sharpen_vs_reference_rcas(gpu_stream, sharpness):
    """k_rcas against the reference's own rcas kernel compiled for the CPU (oracle/_ref): max 1 LSB, >= 99.99 % equal."""
    from oracle import fsr_ref as R
    src = _clip_frame("720p")
    got = gpu_stream.sharpen(src, sharpness)
    for flavour in R.FLAVOURS:
        h = R.lsb_histogram(got, R.sharpen(src, sharpness, flavour))
        print(f"rcas s={sharpness} vs reference/{flavour}: {h}")
        assert h[2] == 0 and h[3] == 0 and h[0] >= 0.9999 * got.size
