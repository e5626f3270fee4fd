"""FrameIngest parity: CUDA ingest / egress kernels (through the C-ABI) vs the reference's own OpenCV calls restated in
oracle/ingest_oracle.py.  Contract: byte work -> bit-exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

FORMATS = ["I420", "I422", "I444", "I40A", "I42A", "YUVA", "NV12", "YUY2", "YVYU", "UYVY", "AYUV", "BGR3", "Y800"]
SIZES = [(64, 36), (1920, 1080), (482, 270), (3840, 2160)]


@pytest.fixture(scope="module")
def I():
    from oracle import ingest_oracle
    return ingest_oracle


def _obs(L, frame):
    return L.ObsFrame(frame["format"], frame["width"], frame["height"], [p.copy() for p in frame["planes"]], timestamp=77)


@pytest.mark.parametrize("fmt", FORMATS)
@pytest.mark.parametrize("size", SIZES[:3])
def test_upload_parity(gpu_stream, I, fmt, size):
    import livevisionkit_b200 as L
    w, h = size
    frame = I.random_frame(fmt, w, h, seed=w + 3 * h)
    ingest = L.FrameIngest.Select(fmt, stream=gpu_stream)
    got = ingest.upload_obs_frame(_obs(L, frame))
    ref = I.upload_obs_frame(frame)
    assert got.timestamp == 77 and got.format == ingest.ocl_format()
    assert got.data.shape == ref.shape
    assert (got.data == ref).all(), f"{fmt} {size}: {(got.data != ref).sum()} bytes differ"


@pytest.mark.parametrize("fmt", FORMATS)
@pytest.mark.parametrize("size", SIZES[:3])
def test_download_parity(gpu_stream, I, fmt, size):
    import livevisionkit_b200 as L
    w, h = size
    rng = np.random.default_rng(w * h + len(fmt))
    ch = 1 if fmt == "Y800" else 3
    img = rng.integers(0, 256, size=(h, w, ch) if ch == 3 else (h, w), dtype=np.uint8)
    ref = I.download_ocl_frame(img, fmt)
    ingest = L.FrameIngest.Select(fmt, stream=gpu_stream)
    dst = L.ObsFrame(fmt, w, h, [np.full((rows, rb), 13, np.uint8) for rows, rb in I.plane_shapes(fmt, w, h)])
    ingest.download_ocl_frame(L.VideoFrame(img, 5, ingest.ocl_format()), dst)
    assert dst.timestamp == 5
    for a, b in zip(ref, dst.planes):
        assert (a.reshape(b.shape) == b).all(), f"{fmt} {size}: {(a.reshape(b.shape) != b).sum()} bytes differ"


def test_unsupported_and_invalid(gpu_stream, I):
    import livevisionkit_b200 as L
    assert L.FrameIngest.Select("P010") is None  # FrameIngest::Select -> nullptr
    frame = I.random_frame("I420", 64, 36)
    odd = L.ObsFrame("I420", 63, 36, [np.zeros((36, 63), np.uint8), np.zeros((18, 31), np.uint8), np.zeros((18, 31), np.uint8)])
    with pytest.raises(L.LvkB200Error):
        L.FrameIngest.Select("I420", stream=gpu_stream).upload_obs_frame(odd)
    with pytest.raises(L.LvkB200Error):  # download into a layout whose ocl format differs from the frame's
        L.FrameIngest.Select("I420", stream=gpu_stream).download_ocl_frame(
            L.VideoFrame(np.zeros((36, 64, 3), np.uint8), 0, L.BGR), _obs(L, frame))


def test_device_planes_with_pitch_4k(gpu_stream, I):
    """Device-resident NV12 at the largest configured size, planes with a padded pitch."""
    torch = pytest.importorskip("torch")
    import livevisionkit_b200 as L
    w, h = SIZES[3]
    frame = I.random_frame("NV12", w, h, seed=1)
    ref = I.upload_obs_frame(frame)
    pad = 64
    planes = []
    for p in frame["planes"]:
        t = torch.zeros((p.shape[0], p.shape[1] + pad), dtype=torch.uint8, device="cuda")
        t[:, :p.shape[1]] = torch.from_numpy(p).cuda()
        planes.append(t[:, :p.shape[1]])
    ingest = L.FrameIngest.Select("NV12", stream=gpu_stream)
    out = torch.empty((h, w, 3), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    ingest.upload_obs_frame(L.ObsFrame("NV12", w, h, planes), out=out)
    gpu_stream.sync()
    assert (out.cpu().numpy() == ref).all()
    # and back: idempotence of download(upload(x)) on luma, exact parity on chroma
    back = [torch.empty_like(p) for p in planes]
    ingest.download_ocl_frame(L.VideoFrame(out, 0, L.YUV), L.ObsFrame("NV12", w, h, back))
    gpu_stream.sync()
    refb = I.download_ocl_frame(ref, "NV12")
    for a, b in zip(refb, back):
        assert (a.reshape(b.shape) == b.cpu().numpy()).all()


def test_submit_obs_matches_packed_submit(I):
    """VSFilter async path: NV12 in -> stabilize -> NV12 out equals ingest + packed submit + egress done step by step."""
    import livevisionkit_b200 as L
    from tools.synth import Clip
    w, h = 640, 360
    clip = Clip((w, h), "shake", frames=14)
    settings = L.StabilizationFilterSettings.obs_homography_preset()
    a, b = L.Stream(settings, 0), L.Stream(settings, 0)
    ingest = L.FrameIngest.Select("NV12", stream=L.Stream(None, 0))
    outputs = 0
    for i in range(14):
        planes = I.download_ocl_frame(clip[i], "NV12")  # synthetic NV12 source (BGR bytes reinterpreted as YUV)
        src = L.ObsFrame("NV12", w, h, [p.copy() for p in planes], timestamp=i)
        dst = L.ObsFrame("NV12", w, h, [np.zeros_like(p) for p in planes])
        res = a.submit_obs(src, dst)
        packed = ingest.upload_obs_frame(src).data
        out = np.empty_like(packed)
        res2 = b.submit(packed, out, L.YUV, i)
        assert res.has_output == res2.has_output
        if res.has_output:
            outputs += 1
            want = L.ObsFrame("NV12", w, h, [np.zeros_like(p) for p in planes])
            ingest.download_ocl_frame(L.VideoFrame(out, res2.out_timestamp, L.YUV), want)
            assert dst.timestamp == res2.out_timestamp == i - 10
            for p, q in zip(dst.planes, want.planes):
                assert (p == q).all()
    assert outputs == 4
    a.close(); b.close()


@pytest.mark.parametrize("fmt", ["NV12", "I420", "UYVY"])
def test_pipelined_obs_stream_equals_submit_obs(I, fmt):
    """lvkb200_stream_prefetch_obs / _submit_obs_async (upload + to_ocl on the copy-in stream, to_obs + download on the
    copy-out stream, two outputs in flight) deliver exactly the planes and timestamps of the synchronous
    lvkb200_stream_submit_obs, frame for frame."""
    import livevisionkit_b200 as L
    from tools.synth import Clip
    w, h, n = 1280, 720, 24
    clip = Clip((w, h), "shake", frames=n)
    settings = L.StabilizationFilterSettings.obs_homography_preset()
    a, b = L.Stream(settings, 0), L.Stream(settings, 0)
    sources = []
    for i in range(n):
        planes = I.download_ocl_frame(clip[i], fmt)  # synthetic source in the layout (BGR bytes reinterpreted as YUV)
        sources.append(L.ObsFrame(fmt, w, h, [np.ascontiguousarray(p).reshape(s) for p, s in zip(planes, I.plane_shapes(fmt, w, h))],
                                  timestamp=500 + i))
    want = []
    for src in sources:
        dst = L.ObsFrame(fmt, w, h, [np.zeros_like(p) for p in src.planes])
        if a.submit_obs(src, dst).has_output:
            want.append((dst.timestamp, [p.copy() for p in dst.planes]))
    outs = [L.ObsFrame(fmt, w, h, [np.zeros_like(p) for p in sources[0].planes]) for _ in range(3)]
    got = []
    delivered = b.stream_obs(sources, lambda o: got.append((o.timestamp, [p.copy() for p in o.planes])) and False, outs)
    assert delivered == len(want) == n - 10 and len(got) == len(want)
    # the whole sequence through ONE call (lvkb200_stream_submit_obs_batch), one output frame per input
    c = L.Stream(settings, 0)
    bouts = [L.ObsFrame(fmt, w, h, [np.zeros_like(p) for p in sources[0].planes]) for _ in range(n)]
    res = c.submit_obs_batch(sources, bouts)
    batch = [(bouts[i].timestamp, bouts[i].planes) for i in range(n) if res[i].has_output]
    assert len(batch) == len(want)
    for (ts_a, pa), (ts_b, pb) in zip(want, batch):
        assert ts_a == ts_b and all((p == q).all() for p, q in zip(pa, pb)), f"{fmt}: batch planes differ at {ts_a}"
    c.close()
    for (ts_a, pa), (ts_b, pb) in zip(want, got):
        assert ts_a == ts_b
        for p, q in zip(pa, pb):
            assert (p == q).all(), f"{fmt}: pipelined planes differ at timestamp {ts_a}"
    a.close(); b.close()
