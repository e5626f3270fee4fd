"""Free-running parity: lvkb200 StabilizationFilter (C-ABI, CUDA) vs the oracle on the same synthetic clip.

What is asserted, per frame:
  * output cadence (empty frames while the look-ahead queue fills), timestamps, QA state;
  * detection image bit-exact; FAST/grid feature list bit-exact while the two states are identical;
  * LK status identical, matched points <= 0.01 px; inlier masks identical (clean clip);
  * estimated homography within the stated corner-displacement bound of cv2's USAC model;
  * the remap inside the pipeline is bit-exact GIVEN the pipeline's own transform (oracle remap of the queued frame
    with the GPU's dst->src transform), and the final pixels are compared to the oracle's final pixels as a report.
"""
import numpy as np
import pytest

# bit-exact comparisons with the oracle restatement: these modules run the EXACT arithmetic build of the EASU kernels
# (tests that exercise the default contract build say so and switch it back on)
pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("exact_build")]

cv2 = pytest.importorskip("cv2")


def _corner_disp(A, B, w, h):
    c = np.array([[0, 0, 1], [w, 0, 1], [0, h, 1], [w, h, 1]], dtype=np.float64).T
    a, b = A @ c, B @ c
    return float(np.abs(a[:2] / a[2] - b[:2] / b[2]).max())


def _kp_array(tr):
    return np.array(tr, dtype=np.float64).reshape(-1, 4)


@pytest.mark.parametrize("preset,res,frames", [("H", "1080p", 36), ("D", "720p", 30), ("F", "1080p", 30)])
def test_free_running_vs_oracle(oracle, preset, res, frames):
    import livevisionkit_b200 as L
    from livevisionkit_b200 import _capi as K
    from tools.synth import Clip

    clip = Clip(res, "pan" if preset == "D" else "shake", frames=frames, fps=30 if preset == "D" else 60)
    # H = OBS "Homography" preset, D = library defaults (2x2 mesh, LSCG), F = OBS "Vector Field" preset (16x16 mesh)
    so = {"H": oracle.StabilizationSettings.obs_homography_preset, "D": oracle.StabilizationSettings,
          "F": oracle.StabilizationSettings.obs_field_preset}[preset]()
    sg = {"H": L.StabilizationFilterSettings.obs_homography_preset, "D": L.StabilizationFilterSettings,
          "F": L.StabilizationFilterSettings.obs_field_preset}[preset]()
    ref = oracle.StabilizationFilter(so)
    flt = L.StabilizationFilter(sg, device=0)
    ref.restart()  # scene quality 1.0 -> the trust factor ramps up immediately, corrections become non-trivial
    flt.restart()
    flt.stream.set_debug_capture(True)
    w, h = clip.width, clip.height
    det_w, det_h = so.detection_resolution
    queue = []
    stats = {"max_H_disp": 0.0, "max_lk": 0.0, "max_T_disp": 0.0, "pix_exact_given_T": [], "pix_vs_oracle": []}
    in_sync = True
    for i in range(frames):
        frame = clip[i]
        queue.append(frame)
        out_ref, ts_ref = ref.apply(frame, oracle.BGR, 1000 + i)
        tr = ref.trace
        vf = flt.apply(L.VideoFrame(frame, 1000 + i, L.BGR))
        res_g = flt.last_result
        s = flt.stream

        # cadence / metadata
        assert (out_ref is None) == vf.empty(), f"frame {i}: output cadence differs"
        assert bool(res_g.has_motion) == bool(tr["has_motion"]), f"frame {i}: has_motion differs"
        assert abs(res_g.trust_factor - float(tr["trust"])) < 1e-6
        assert abs(res_g.scene_quality - float(tr["scene_quality"])) < 1e-5

        # K0
        det = s.debug_fetch(K.DBG_DETECTION_IMAGE, np.uint8).reshape(det_h, det_w)
        assert (det == tr["det"]).all(), f"frame {i}: detection image differs"

        if "detected" in tr:
            kd = s.debug_fetch(K.DBG_DETECTED, np.dtype([("x", "f4"), ("y", "f4"), ("r", "f4"), ("c", "i4")]))
            a = _kp_array(tr["detected"])
            assert len(kd) == len(a), f"frame {i}: feature count {len(kd)} vs {len(a)}"
            g = np.stack([kd["x"], kd["y"], kd["r"], kd["c"]], axis=1).astype(np.float64)
            same = bool((g == a).all())
            if in_sync and not same:
                # float LK differences (<= 1e-4 px) in propagated positions are the only legitimate source
                assert np.abs(g[:, :2] - a[:, :2]).max() <= 0.01 and (g[:, 2:] == a[:, 2:]).all(), \
                    f"frame {i}: feature lists diverged"
            fc = s.debug_fetch(K.DBG_FAST_COUNTS, np.int32)
            assert list(fc) == list(tr["fast_counts"]), f"frame {i}: FAST counts {list(fc)} vs {tr['fast_counts']}"
        if "lk_status" in tr:
            st = s.debug_fetch(K.DBG_LK_STATUS, np.uint8)
            mt = s.debug_fetch(K.DBG_LK_MATCHED, np.float32).reshape(-1, 2)
            assert (st == tr["lk_status"]).all(), f"frame {i}: LK status differs"
            ok = st == 1
            stats["max_lk"] = max(stats["max_lk"], float(np.abs(mt[ok] - tr["lk_out"][ok]).max()))
        if "inliers" in tr:
            inl = s.debug_fetch(K.DBG_INLIERS, np.uint8)
            assert (inl == tr["inliers"]).all(), f"frame {i}: inlier mask differs"
            assert abs(res_g.tracking_stability - float(tr["stability"])) < 1e-6
        if "H" in tr:
            Hg = s.debug_fetch(K.DBG_HOMOGRAPHY, np.float64).reshape(3, 3)
            stats["max_H_disp"] = max(stats["max_H_disp"], _corner_disp(Hg, tr["H"], det_w, det_h))

        if not vf.empty():
            src = queue[i - ref.frame_delay()]
            assert vf.timestamp == ts_ref == 1000 + i - ref.frame_delay()
            if sg.motion_resolution == (2, 2):
                T = s.debug_fetch(K.DBG_WARP_TRANSFORM, np.float64).reshape(3, 3)
                stats["max_T_disp"] = max(stats["max_T_disp"], _corner_disp(T, tr["warp_T"], w, h))
                given = oracle.remap_homography(src, T, so.background_colour, False)
                stats["pix_exact_given_T"].append(float((given == vf.data).mean()))
                assert np.abs(given.astype(np.int16) - vf.data.astype(np.int16)).max() <= 1
            d = np.abs(out_ref.astype(np.int16) - vf.data.astype(np.int16))
            stats["pix_vs_oracle"].append((int(d.max()), float((d <= 1).mean())))
    print(f"[{preset} {res}] max LK |dpos| {stats['max_lk']:.2e} px; max H corner disp vs cv2 {stats['max_H_disp']:.4f} px "
          f"(detection res); max warp-transform corner disp {stats['max_T_disp']:.4f} px (frame res); "
          f"pixels bit-exact given own transform: min {min(stats['pix_exact_given_T'] or [1.0]):.6f}; "
          f"final pixels vs oracle (max |d|, frac<=1LSB): {stats['pix_vs_oracle'][-3:]}")
    assert stats["max_lk"] <= 0.01
    assert stats["max_H_disp"] <= 0.25
    assert min(stats["pix_exact_given_T"] or [1.0]) == 1.0
    assert flt.frame_delay() == ref.frame_delay() == 10
    if preset == "F":  # mesh path end to end: the per-pixel offsets come from a 512-unknown LSCG solve on both sides
        assert all(frac >= 0.97 for _, frac in stats["pix_vs_oracle"]), stats["pix_vs_oracle"]


def test_host_and_device_frames_agree(oracle):
    torch = pytest.importorskip("torch")
    import livevisionkit_b200 as L
    from tools.synth import Clip
    clip = Clip("720p", "shake", frames=14)
    s = L.StabilizationFilterSettings.obs_homography_preset()
    a, b = L.StabilizationFilter(s, 0), L.StabilizationFilter(s, 0)
    for i in range(14):
        f = clip[i]
        va = a.apply(L.VideoFrame(f, i, L.BGR))
        d = torch.from_numpy(f).cuda()
        out = torch.empty_like(d)
        torch.cuda.synchronize()
        vb = b.apply(L.VideoFrame(d, i, L.BGR), output=out)
        b.stream.sync()
        assert va.empty() == vb.empty()
        if not va.empty():
            assert (va.data == vb.data.cpu().numpy()).all() and va.timestamp == vb.timestamp == i - 10


def test_pipelined_stream_equals_apply(oracle):
    """VideoFilter::stream analogue (prefetch + submit_async + wait_output on extra CUDA streams) must deliver exactly
    what the synchronous apply() delivers, frame for frame."""
    torch = pytest.importorskip("torch")
    import livevisionkit_b200 as L
    from tools.synth import Clip
    clip = Clip("720p", "shake", frames=26)
    s = L.StabilizationFilterSettings.obs_homography_preset()
    ref = L.StabilizationFilter(s, 0)
    expected = []
    for i in range(26):
        v = ref.apply(L.VideoFrame(clip[i], 100 + i, L.BGR))
        if not v.empty():
            expected.append((v.timestamp, v.data.copy()))
    pipe = L.StabilizationFilter(s, 0)
    frames = [L.VideoFrame(torch.from_numpy(clip[i]).pin_memory(), 100 + i, L.BGR) for i in range(26)]
    outs = [torch.empty_like(frames[0].data).pin_memory() for _ in range(3)]
    got = []
    n = pipe.stream(frames, lambda vf: got.append((vf.timestamp, vf.data.numpy().copy())), outs)
    assert n == len(expected) == 16
    for (te, de), (tg, dg) in zip(expected, got):
        assert te == tg and (de == dg).all()
    # early stop: a true return from the callback terminates the stream (VideoFilter.cpp:180-206)
    pipe2 = L.StabilizationFilter(s, 0)
    seen = []
    n2 = pipe2.stream(frames, lambda vf: (seen.append(vf.timestamp), len(seen) >= 3)[1], outs)
    assert n2 == 3 and seen == [100, 101, 102]


@pytest.mark.parametrize("preset,deblock", [("H", False), ("D", False), ("H", True), ("F", False)])
def test_lookahead_equals_plain_submit(oracle, preset, deblock):
    """lvkb200_stream_prefetch_frame (the next frame's copy, detection image and pyramid queued behind the current
    frame's tracking chain) only moves work earlier: every output byte and every tracker observable must equal the
    plain submit sequence — device-resident frames, with and without the chained deblocking stage, including a
    frame that was announced but replaced, a restart and a reconfiguration in mid-stream."""
    torch = pytest.importorskip("torch")
    import livevisionkit_b200 as L
    from livevisionkit_b200 import _capi as K
    from tools.synth import Clip
    n = 30
    clip = Clip("720p", "shake", frames=n)
    mk = {"H": L.StabilizationFilterSettings.obs_homography_preset, "D": L.StabilizationFilterSettings,
          "F": L.StabilizationFilterSettings.obs_field_preset}[preset]
    dev = [torch.from_numpy(clip[i]).cuda() for i in range(n)]
    decoy = torch.from_numpy(clip[3]).cuda()

    def run(lookahead):
        flt = L.StabilizationFilter(mk(), 0)
        if deblock:
            flt.stream.set_deblocking(L.DeblockingFilterSettings())
        outs, results = [], []
        for i in range(n):
            if i == 17:
                flt.restart()
            if i == 22:
                s2 = mk()
                s2.crop_to_stable_region = not s2.crop_to_stable_region
                flt.configure(s2)
            if lookahead and i + 1 < n:
                # frame 12 is announced with a buffer that is then NOT the one submitted: its look-ahead must be dropped
                flt.stream.prefetch(decoy if i + 1 == 12 else dev[i + 1], L.BGR)
            out = torch.empty_like(dev[0])
            r = flt.stream.submit(dev[i], out, L.BGR, i)
            flt.stream.sync()
            results.append((r.has_output, r.out_timestamp, r.feature_count, round(r.tracking_stability, 6), r.has_motion))
            outs.append(out.cpu().numpy() if r.has_output else None)
        flt.stream.close()
        return outs, results

    plain_o, plain_r = run(False)
    ahead_o, ahead_r = run(True)
    assert plain_r == ahead_r
    assert sum(o is not None for o in plain_o) >= 8  # the restart empties the 10-frame queue once
    for a, b in zip(plain_o, ahead_o):
        assert (a is None) == (b is None)
        if a is not None:
            assert (a == b).all()


def test_concurrent_streams_from_two_host_threads(oracle):
    """SURVEY 8(b) threading contract: a handle is externally synchronised, DIFFERENT handles are fully concurrent (own
    CUDA streams, no shared mutable state).  Two host threads drive two stabilizers (different presets, different
    clips) on the same GPU at the same time (ctypes releases the GIL inside every call); each must produce exactly the
    bytes of its own sequential run."""
    import threading
    import livevisionkit_b200 as L
    from tools.synth import Clip
    n = 24
    jobs = [(L.StabilizationFilterSettings.obs_homography_preset, Clip("720p", "shake", frames=n, seed=5)),
            (L.StabilizationFilterSettings, Clip((960, 540), "pan", frames=n, seed=6))]
    frames = [[c[i] for i in range(n)] for _, c in jobs]

    def run(k, sink):
        flt = L.StabilizationFilter(jobs[k][0](), 0)
        for i in range(n):
            v = flt.apply(L.VideoFrame(frames[k][i], i, L.BGR))
            sink.append(None if v.empty() else (v.timestamp, v.data.copy()))
        flt.stream.close()

    sequential = [[], []]
    for k in range(2):
        run(k, sequential[k])
    concurrent = [[], []]
    errors = []

    def guarded(k):
        try:
            run(k, concurrent[k])
        except Exception as e:  # surfaced below: an exception in a thread must fail the test
            errors.append(e)

    threads = [threading.Thread(target=guarded, args=(k,)) for k in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for k in range(2):
        assert len(sequential[k]) == len(concurrent[k]) == n
        assert sum(o is not None for o in sequential[k]) == n - 10
        for a, b in zip(sequential[k], concurrent[k]):
            assert (a is None) == (b is None)
            if a is not None:
                assert a[0] == b[0] and (a[1] == b[1]).all()


def test_stabilize_output_off_is_a_pure_delay(oracle):
    import livevisionkit_b200 as L
    from tools.synth import Clip
    clip = Clip((640, 360), "shake", frames=13)
    s = L.StabilizationFilterSettings.obs_homography_preset()
    s.stabilize_output = False
    f = L.StabilizationFilter(s, 0)
    for i in range(13):
        v = f.apply(L.VideoFrame(clip[i], i, L.BGR))
        if i < 10:
            assert v.empty()
        else:
            assert (v.data == clip[i - 10]).all() and v.timestamp == i - 10


def test_configure_preconditions(oracle):
    import livevisionkit_b200 as L
    s = L.StabilizationFilterSettings()
    s.min_tracking_quality = 1.5  # LVK_ASSERT_01 (StabilizationFilter.cpp:44)
    with pytest.raises(L.LvkB200Error) as e:
        L.StabilizationFilter(s, 0)
    assert e.value.status == 1


def test_submit_batch_equals_per_frame_submit():
    """lvkb200_stream_submit_batch (one FFI call for a whole sequence, frame i+1 announced before frame i) produces exactly
    the outputs, cadence and timestamps of per-frame lvkb200_stream_submit."""
    torch = pytest.importorskip("torch")
    import livevisionkit_b200 as L
    from tools.synth import Clip
    clip = Clip((1280, 720), "shake", frames=26)
    frames = [torch.from_numpy(clip[i]).cuda() for i in range(26)]
    settings = L.StabilizationFilterSettings.obs_homography_preset()
    a, b = L.Stream(settings, 0), L.Stream(settings, 0)
    outs_a = [torch.zeros_like(frames[0]) for _ in range(26)]
    outs_b = [torch.zeros_like(frames[0]) for _ in range(26)]
    res_a = [a.submit(frames[i], outs_a[i], L.BGR, 100 + i) for i in range(26)]
    a.sync()
    res_b = list(b.submit_batch(frames[:9], outs_b[:9], L.BGR, [100 + i for i in range(9)]))
    res_b += list(b.submit_batch(frames[9:], outs_b[9:], L.BGR, [100 + i for i in range(9, 26)]))
    b.sync()
    assert sum(r.has_output for r in res_a) == 16
    for i in range(26):
        assert res_a[i].has_output == res_b[i].has_output and res_a[i].out_timestamp == res_b[i].out_timestamp
        assert abs(res_a[i].trust_factor - res_b[i].trust_factor) == 0.0
        assert torch.equal(outs_a[i], outs_b[i]), f"frame {i}: batch output differs"
    a.close()
    b.close()


def test_submit_batch_host_pipelined_equals_apply():
    """lvkb200_stream_submit_batch with HOST frames and HOST outputs runs the pipelined path inside one call (outputs
    cycling through four buffers): every output equals the synchronous per-frame apply()."""
    import livevisionkit_b200 as L
    from tools.synth import Clip
    n = 24
    clip = Clip((960, 540), "shake", frames=n)
    frames = [clip[i] for i in range(n)]
    settings = L.StabilizationFilterSettings.obs_homography_preset()
    a, b = L.StabilizationFilter(settings, 0), L.Stream(settings, 0)
    want = [a.apply(L.VideoFrame(frames[i], 7 + i, L.BGR)) for i in range(n)]
    outs = [np.zeros_like(frames[0]) for _ in range(n)]
    res = b.submit_batch(frames, outs, L.BGR, [7 + i for i in range(n)])
    for i in range(n):
        assert bool(res[i].has_output) == (not want[i].empty())
        if res[i].has_output:
            assert res[i].out_timestamp == want[i].timestamp and (outs[i] == want[i].data).all(), f"frame {i}"
    # cycling outputs: the last four outputs are intact when the call returns
    ring = [np.zeros_like(frames[0]) for _ in range(4)]
    c = L.Stream(settings, 0)
    c.submit_batch(frames, [ring[i % 4] for i in range(n)], L.BGR, [7 + i for i in range(n)])
    for i in range(n - 4, n):
        assert (ring[i % 4] == want[i].data).all()
    b.close(); c.close()
