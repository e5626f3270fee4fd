"""The C++ mirror of the reference interface (livevisionkit_b200/compat/lvk/lvk.hpp: lvk::StabilizationFilter,
lvk::VideoFrame, settings structs, assert_handler) compiled like a reference caller and run on the GPU; its output
is compared with the Python mirror on the same frames."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _frame(i, w=640, h=360):
    x = np.arange(w)[None, :] + 2 * i
    y = np.arange(h)[:, None] + i
    v = (((x // 16 + y // 16) % 2) * 120 + ((x * 7 + y * 13) % 61) + 40).astype(np.uint8)
    return np.stack([v, (v.astype(np.int32) * 9 // 10).astype(np.uint8), (v.astype(np.int32) * 8 // 10).astype(np.uint8)],
                    axis=-1)


def test_cpp_compat_layer_matches_python_mirror(tmp_path):
    import livevisionkit_b200 as L
    exe = str(tmp_path / "test_compat")
    libdir = os.path.join(ROOT, "livevisionkit_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", os.path.join(ROOT, "tests", "cpp", "test_compat.cpp"),
                           "-o", exe, f"-L{libdir}", "-l:liblvkb200.so", f"-Wl,-rpath,{libdir}"])
    out = subprocess.run([exe, "14"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    lines = [l.split() for l in out.stdout.strip().splitlines()]
    assert "min_tracking_quality" in out.stderr  # the failed precondition reached lvk::context::assert_handler

    flt = L.StabilizationFilter(L.StabilizationFilterSettings.obs_homography_preset(), 0)
    for i in range(14):
        v = flt.apply(L.VideoFrame(_frame(i), 1000 + i, L.BGR))
        idx, ts, total = lines[i]
        assert int(idx) == i
        if v.empty():
            assert ts == "empty"
        else:
            assert int(ts) == v.timestamp == 1000 + i - 10
            assert int(total) == int(v.data.astype(np.uint64).sum())
    assert lines[14][0] == "timing_ms" and float(lines[14][1]) > 0.0

    # lvk::ScalingFilter / lvk::DeblockingFilter / lvk::CompositeFilter of the compat header on the last frame
    def checksum(a):
        hh, ww = a.shape[:2]
        wgt = 1 + (np.arange(ww * 3)[None, :] + np.arange(hh)[:, None]) % 7
        return int((a.reshape(hh, ww * 3).astype(np.uint64) * wgt.astype(np.uint64)).sum())

    assert lines[15] == ["stream", "4"]  # VideoFilter::stream delivered the 4 outputs apply() produced (checked in C++)
    named = {l[0]: int(l[1]) for l in lines[16:]}
    frame = _frame(13)
    scaler = L.ScalingFilter(L.ScalingFilterSettings((960, 540), 0.8, False), 0)
    assert named["scaling"] == checksum(scaler.apply(L.VideoFrame(frame, 77, L.BGR)).data)
    deblocked = L.DeblockingFilter(device=0).apply(L.VideoFrame(frame.copy(), 77, L.BGR)).data
    assert named["deblocking"] == checksum(deblocked)
    # the C++ deblocker worked in place on the caller's pixels (shallow copy, VideoFilter.cpp:55-58): the chain saw them
    twice = L.DeblockingFilter(device=0).apply(L.VideoFrame(deblocked.copy(), 77, L.BGR)).data
    assert named["composite"] == checksum(scaler.apply(L.VideoFrame(twice, 77, L.BGR)).data)
    assert "sharpness" in out.stderr  # ScalingFilter::configure precondition reached the assert handler


def test_cpp_compat_layer_with_umat_videoframe(tmp_path):
    """lvk-compat built with the reference's own `struct VideoFrame : cv::UMat` (mock opencv2/), driven like the OBS
    plugin (apply(std::move(frame), frame), VSFilter.cpp:352-364) and through VideoFilter::stream(cv::VideoCapture&):
    same outputs as the Python mirror, Stopwatch statistics available through timings()."""
    import livevisionkit_b200 as L
    exe = str(tmp_path / "test_compat_umat")
    libdir = os.path.join(ROOT, "livevisionkit_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror",
                           "-I" + os.path.join(ROOT, "tests", "cpp", "mock_opencv"),
                           os.path.join(ROOT, "tests", "cpp", "test_compat_umat.cpp"), "-o", exe, f"-L{libdir}",
                           "-l:liblvkb200.so", f"-Wl,-rpath,{libdir}"])
    out = subprocess.run([exe, "14"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    lines = [l.split() for l in out.stdout.strip().splitlines()]
    flt = L.StabilizationFilter(L.StabilizationFilterSettings.obs_homography_preset(), 0)
    expected = []
    for i in range(14):
        v = flt.apply(L.VideoFrame(_frame(i), 1000 + i, L.BGR))
        idx, ts, total = lines[i]
        assert int(idx) == i
        if v.empty():
            assert ts == "empty"
        else:
            assert int(ts) == v.timestamp and int(total) == int(v.data.astype(np.uint64).sum())
            expected.append(int(total))
    assert lines[14][0] == "timings" and int(lines[14][1]) == 14 and float(lines[14][2]) > 0.0 and float(lines[14][3]) >= 0.0
    streamed = [l for l in lines if l[0].startswith("s") and l[0] != "stream"]
    # the capture path stamps frames 0, 1, ... (VideoFilter.cpp:96-100): outputs are frames 0..3, same pixels as apply()
    assert [int(l[2]) for l in streamed] == expected and [int(l[1]) for l in streamed] == [0, 1, 2, 3]
    assert lines[-1] == ["stream", "4"]


EDITOR = os.path.join(ROOT, "oracle", "_ref", "lvk-editor")


def _read_raw_clip(path):
    with open(path, "rb") as f:
        magic, w, h, _fps, n = f.readline().split()
        assert magic == b"LVKRAW1"
        w, h, n = int(w), int(h), int(n)
        return [np.frombuffer(f.read(w * h * 3), np.uint8).reshape(h, w, 3) for _ in range(n)]


@pytest.mark.skipif(not os.path.isfile(EDITOR), reason="oracle/_ref/lvk-editor is built where /root/reference exists "
                                                       "(oracle/ref_build/build_lvk_editor.sh); it travels prebuilt")
def test_reference_video_editor_unchanged_runs_config5_chain(tmp_path):
    """The reference's OWN VideoEditor CLI - its sources compiled unchanged against lvk-compat and linked with
    liblvkb200.so (tests/test_compat_cpu.py builds it; oracle/ref_build/build_lvk_editor.sh) - run the way BASELINE
    config 5 is spelled on its command line: `lvk-editor in out -f adb .l 2 -f vs .s 6 .cp 0.08`.  Frames go
    cv::VideoCapture -> CompositeFilter::stream -> {DeblockingFilter, StabilizationFilter} -> cv::VideoWriter (the
    mock OpenCV's raw-clip container); the written clip must equal, byte for byte, what the Python mirror produces for
    the same chain and settings through the same C-ABI."""
    import livevisionkit_b200 as L
    n = 30
    src, dst = str(tmp_path / "in.raw"), str(tmp_path / "out.raw")
    with open(src, "wb") as f:
        f.write(b"LVKRAW1 %d %d %.6f %010d\n" % (640, 360, 30.0, n))
        for i in range(n):
            f.write(_frame(i).tobytes())
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(ROOT, "livevisionkit_b200") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    if not os.access(EDITOR, os.X_OK):  # a snapshot that dropped the mode bits of the prebuilt binary
        try:
            os.chmod(EDITOR, 0o755)
        except OSError:
            pytest.skip("oracle/_ref/lvk-editor is not executable here")
    run = subprocess.run([EDITOR, src, dst, "-f", "adb", ".l", "2", "-f", "vs", ".s", "6", ".cp", "0.08"],
                         capture_output=True, text=True, timeout=300, env=env)
    assert run.returncode == 0, run.stderr[-2000:]
    written = _read_raw_clip(dst)

    deblocker = L.DeblockingFilter(L.DeblockingFilterSettings(detection_levels=2), device=0)
    settings = L.StabilizationFilterSettings(predictive_samples=6, corrective_limits=(0.08, 0.08))
    stabilizer = L.StabilizationFilter(settings, 0)
    expected = []
    for i in range(n):
        v = stabilizer.apply(deblocker.apply(L.VideoFrame(_frame(i).copy(), i, L.BGR)))
        if not v.empty():
            expected.append(np.array(v.data, copy=True))
    assert len(expected) == n - stabilizer.frame_delay() and len(written) == len(expected), (len(written), len(expected))
    for k, (a, b) in enumerate(zip(written, expected)):
        assert np.array_equal(a, b), f"output frame {k} differs from the Python mirror's"
