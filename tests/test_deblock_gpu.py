"""DeblockingFilter parity on the GPU: the CUDA path through the C-ABI against
  * the restated OpenCV arithmetic (bit-exact: all of it is integer / round-once float32 work), and
  * the cv2 calls the reference itself makes, Filters/DeblockingFilter.cpp:48-118 (<= 1 LSB: cv2's IPP float resize
    differs from OpenCV's own kernels by <= 1 ulp in the blend weights),
stand-alone and chained in front of the stabilizer (CompositeFilter, BASELINE config 5)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def D():
    from oracle import deblock_oracle
    return deblock_oracle


def _frame(D, size, seed=0):
    from tools.synth import Clip
    return D.blocky_frame(Clip(size, "shake", frames=2, seed=10 + seed)[1], 16, 0.8, seed=seed)


def test_golden(gpu_stream):
    import livevisionkit_b200 as L
    g = np.load(os.path.join(G, "deblock_golden.npz"), allow_pickle=False)
    assert (gpu_stream.deblock(g["frame"], None, L.BGR) == g["out_bgr"]).all()
    assert (gpu_stream.deblock(g["frame"], None, L.YUV) == g["out_yuv"]).all()


@pytest.mark.parametrize("size", ["1080p", "720p", "4k", (963, 541), (1366, 768), (170, 120)])
def test_matches_reference_calls(gpu_stream, D, oracle, size):
    import livevisionkit_b200 as L
    frame = _frame(D, size, 1)
    for fmt, lfmt in ((oracle.BGR, L.BGR), (oracle.YUV, L.YUV), (oracle.RGB, L.RGB)):
        got = gpu_stream.deblock(frame, None, lfmt)
        exact = D.deblock_restated(frame, fmt)
        ref = D.DeblockingFilter().apply(frame, fmt)
        d = np.abs(got.astype(int) - ref.astype(int))
        print(f"{size} fmt={fmt}: vs restated {int((got != exact).sum())} bytes, vs cv2 {int((d > 0).sum())} bytes "
              f"(max {int(d.max())}); filter changed {100 * float((ref != frame).mean()):.1f}% of the bytes")
        assert (got == exact).all()
        assert d.max() <= 1 and (d > 0).sum() <= 1e-5 * d.size
        assert (ref != frame).mean() > 0.2


@pytest.mark.parametrize("levels,block,ksize,scale", [(1, 16, 5, 4.0), (2, 16, 3, 4.0), (6, 16, 7, 2.0), (3, 8, 5, 4.0),
                                                      (3, 32, 5, 8.0), (255, 16, 5, 4.0)])
def test_settings(gpu_stream, D, oracle, levels, block, ksize, scale):
    import livevisionkit_b200 as L
    frame = D.blocky_frame(_frame(D, (640, 360), 2), block, 0.8, seed=block)
    so = D.DeblockingSettings(detection_levels=levels, block_size=block, filter_size=ksize, filter_scaling=scale)
    sg = L.DeblockingFilterSettings(detection_levels=levels, block_size=block, filter_size=ksize, filter_scaling=scale)
    got = gpu_stream.deblock(frame, sg, L.BGR)
    assert (got == D.deblock_restated(frame, oracle.BGR, so)).all()
    d = np.abs(got.astype(int) - D.DeblockingFilter(so).apply(frame, oracle.BGR).astype(int))
    assert d.max() <= 1 and (d > 0).sum() <= 1e-5 * d.size


def test_preconditions_and_unsupported(gpu_stream):
    import livevisionkit_b200 as L
    f = np.zeros((64, 64, 3), np.uint8)
    for bad in (dict(block_size=0), dict(filter_size=4), dict(filter_size=1), dict(detection_levels=0),
                dict(filter_scaling=1.0)):  # DeblockingFilter::configure, DeblockingFilter.cpp:38-42
        with pytest.raises(L.LvkB200Error):
            gpu_stream.deblock(f, L.DeblockingFilterSettings(**bad))
        with pytest.raises(L.LvkB200Error):
            L.DeblockingFilter(L.DeblockingFilterSettings(**bad), stream=gpu_stream)
    for unsupported in (dict(filter_scaling=2.5), dict(block_size=12), dict(filter_size=9)):  # stated in the header
        with pytest.raises(L.LvkB200Error) as e:
            gpu_stream.deblock(f, L.DeblockingFilterSettings(**unsupported))
        assert e.value.status == 1
    # no whole macroblock: the frame passes through
    tiny = np.arange(10 * 12 * 3, dtype=np.uint8).reshape(10, 12, 3)
    assert (gpu_stream.deblock(tiny) == tiny).all()


def test_device_memory_in_place_and_filter_class(gpu_stream, D, oracle):
    torch = pytest.importorskip("torch")
    import livevisionkit_b200 as L
    frame = _frame(D, "1080p", 3)
    exact = D.deblock_restated(frame, oracle.BGR)
    d = torch.from_numpy(frame).cuda()
    out = gpu_stream.deblock(d, None, L.BGR, out=d)  # in place, like the reference (output = std::move(input))
    gpu_stream.sync()
    assert (out.cpu().numpy() == exact).all()
    flt = L.DeblockingFilter(stream=gpu_stream)
    res = flt.apply(L.VideoFrame(frame, 42, L.BGR))
    assert res.timestamp == 42 and res.format == L.BGR and (res.data == exact).all()
    assert flt.filter_region(1920, 1080) == (0, 0, 1920, 1072)
    # every textured macroblock stays bit-identical; only flat ones change
    noisy = np.random.default_rng(0).integers(0, 256, (256, 384, 3), dtype=np.uint8)
    assert (gpu_stream.deblock(noisy) == noisy).all()


def test_chained_in_front_of_stabilizer(D, oracle):
    """CompositeFilter{Deblocking, Stabilization}: the fused device chain == deblock every frame, then stabilize."""
    import livevisionkit_b200 as L
    from tools.synth import Clip
    clip = Clip((640, 360), "shake", frames=16, seed=5)
    frames = [D.blocky_frame(clip[i], 16, 0.8, seed=i) for i in range(len(clip))]
    settings = L.StabilizationFilterSettings.obs_homography_preset()
    chain = L.CompositeFilter([L.DeblockingFilter(), L.StabilizationFilter(settings)])
    plain = L.StabilizationFilter(settings)
    outs = 0
    for i, f in enumerate(frames):
        a = chain.apply(L.VideoFrame(f, i, L.BGR))
        b = plain.apply(L.VideoFrame(D.deblock_restated(f, oracle.BGR), i, L.BGR))
        assert a.empty() == b.empty()
        if not a.empty():
            outs += 1
            assert a.timestamp == b.timestamp and (a.data == b.data).all()
    assert outs == len(frames) - 10
    # switching the stage off restores the plain filter
    chain.filters[1].stream.set_deblocking(None)
    chain.filters[1].restart()
    plain.restart()
    for i, f in enumerate(frames[:12]):
        a = chain.filters[1].apply(L.VideoFrame(f, i, L.BGR))
        b = plain.apply(L.VideoFrame(f, i, L.BGR))
        assert a.empty() == b.empty() and (a.empty() or (a.data == b.data).all())


def test_composite_stream_frames_generic_chain():
    """CompositeFilter.stream_frames for a chain that is not the fused one (Stabilization -> Scaling): outputs equal the
    filters applied one after the other, empty frames of the buffering stabilizer are skipped, a true callback return
    stops the stream (VideoFilter.cpp:130-139, 180-206)."""
    import livevisionkit_b200 as L
    from tools.synth import Clip
    clip = Clip((640, 360), "shake", frames=16)
    settings = L.StabilizationFilterSettings.obs_homography_preset()
    scale = L.ScalingFilterSettings((960, 540), 0.8, False)
    chain = L.CompositeFilter([L.StabilizationFilter(settings), L.ScalingFilter(scale)])
    stab, scaler = L.StabilizationFilter(settings), L.ScalingFilter(scale)
    frames = [L.VideoFrame(clip[i], i, L.BGR) for i in range(16)]
    got = []
    n = chain.stream_frames(frames, lambda vf: got.append((vf.timestamp, vf.data.copy())) and False)
    want = []
    for f in frames:
        v = stab.apply(f)
        if not v.empty():
            w = scaler.apply(v)
            want.append((w.timestamp, w.data.copy()))
    assert n == len(want) == 6 and len(got) == 6
    for (ta, a), (tb, b) in zip(want, got):
        assert ta == tb and a.shape == (540, 960, 3) and (a == b).all()
    chain2 = L.CompositeFilter([L.StabilizationFilter(settings), L.ScalingFilter(scale)])
    seen = []
    assert chain2.stream_frames(frames, lambda vf: seen.append(vf.timestamp) or len(seen) >= 2) == 2
