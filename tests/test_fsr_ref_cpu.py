"""Pins the oracle's FSR restatement (oracle/easu_ref.c) to the REFERENCE ITSELF (CPU, no GPU).

tests/golden/fsr_ref_golden.npz was produced by oracle/_ref/libfsrcl_ref_{strict,contract}.so = the reference's own
FSR.cl compiled for the CPU where it lies (oracle/ref_build/build_ref.sh + an OpenCL-C shim; generating script
tests/golden/make_fsr_ref_golden.py).  OpenCL C lets a compiler fuse multiply-adds, so the reference has two legal
arithmetics: strict (nothing fused) and contract (gcc fuses what it can).  They differ from each other by up to 2 LSB;
the restatement must lie within that spread.  Tolerances (stated, asserted):
  EASU vs contract : max 1 LSB, >= 99.99 % of bytes identical
  EASU vs strict   : max 2 LSB, >= 99.9 % identical, < 1e-4 of bytes beyond 1 LSB   (= strict vs contract itself)
  RCAS vs either   : max 1 LSB, >= 99.99 % identical
When the libraries themselves are present (build container, and the GPU box: they travel) the same is checked live
on a 720p frame and the three-way histogram is written to profiles/r02_parity_fsr_ref_cpu.json."""
import json
import os

import numpy as np
import pytest

from oracle import fsr_ref as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = np.load(os.path.join(ROOT, "tests", "golden", "fsr_ref_golden.npz"), allow_pickle=False)


def _check(name, got, want_strict, want_contract, exact=False):
    hs, hc = R.lsb_histogram(got, want_strict), R.lsb_histogram(got, want_contract)
    n = got.size
    if exact:  # RCAS: no position arithmetic, so no 2-LSB texel flips: max 1 LSB against either build
        assert hs[2:] == [0, 0] and hc[2:] == [0, 0] and min(hs[0], hc[0]) >= 0.9999 * n, f"{name}: strict {hs} contract {hc}"
        return hs, hc
    assert hc[2] == 0 and hc[3] == 0 and hc[0] >= 0.9999 * n, f"{name}: vs contract build {hc}"
    assert hs[3] == 0 and hs[2] < 1e-4 * n and hs[0] >= 0.999 * n, f"{name}: vs strict build {hs}"
    return hs, hc


def test_reference_builds_bracket_each_other():
    """The honest tolerance of an unpinned OpenCL device: strict vs contract of the SAME reference source."""
    for key in ("homography", "map", "scale0", "scale1", "scale2", "scale3"):
        h = R.lsb_histogram(G[f"{key}_strict"], G[f"{key}_contract"])
        assert h[3] == 0 and h[0] >= 0.999 * G[f"{key}_strict"].size, (key, h)
    assert (G["rcas_strict"] == G["rcas_contract"]).mean() > 0.9999


def test_restatement_homography_vs_reference(oracle):
    k = 0
    for t in G["transforms"]:
        for yuv in (False, True):
            out = oracle.remap_homography(G["src"], t, (255, 0, 255), yuv)
            _check(f"homography[{k}]", out, G["homography_strict"][k], G["homography_contract"][k])
            k += 1


def test_restatement_offset_map_vs_reference(oracle):
    for k, yuv in enumerate((False, True)):
        out = oracle.remap_map(G["src"], G["offset_map"], (0, 0, 0), yuv)
        _check(f"map[{k}]", out, G["map_strict"][k], G["map_contract"][k])


def test_restatement_upscale_vs_reference(oracle):
    for i, sz in enumerate(G["scale_sizes"]):
        for k, yuv in enumerate((False, True)):
            out = oracle.upscale(G["scale_src"], (int(sz[0]), int(sz[1])), yuv)
            _check(f"scale{i}[{k}]", out, G[f"scale{i}_strict"][k], G[f"scale{i}_contract"][k])


def test_restatement_rcas_vs_reference(oracle):
    for k, s in enumerate(G["rcas_sharpness"]):
        out = oracle.sharpen(G["rcas_src"], float(s))
        _check(f"rcas[{k}]", out, G["rcas_strict"][k], G["rcas_contract"][k], exact=True)


@pytest.mark.skipif(not (R.available("strict") and R.available("contract")), reason="oracle/_ref not built")
def test_live_reference_720p_and_artifact(oracle):
    """Same comparison on a 1280x720 hand-shake frame with the live libraries; writes the parity artifact."""
    from tools.synth import Clip
    frame = Clip((1280, 720), "shake", frames=3)[1]
    a = np.radians(0.7)
    t = np.array([[np.cos(a) * 1.01, -np.sin(a), 7.3], [np.sin(a), np.cos(a) * 1.01, -4.6], [1e-6, -2e-6, 1.0]])
    rec = {"what": "oracle/easu_ref.c (restatement) vs the reference's FSR.cl compiled for the CPU, 1280x720 frame, "
                   "histograms [#bytes |d|=0, 1, 2, >=3]", "bytes": int(frame.size)}
    for yuv in (False, True):
        o = oracle.remap_homography(frame, t, (255, 0, 255), yuv)
        s = R.remap_homography(frame, t, (255, 0, 255), yuv, "strict")
        c = R.remap_homography(frame, t, (255, 0, 255), yuv, "contract")
        hs, hc = _check(f"720p yuv={yuv}", o, s, c)
        rec[f"homography_yuv{int(yuv)}"] = {"restatement_vs_strict": hs, "restatement_vs_contract": hc,
                                            "strict_vs_contract": R.lsb_histogram(s, c)}
    up_o = oracle.upscale(frame[:360, :640], (1280, 720))
    rec["upscale_2x"] = {"restatement_vs_strict": R.lsb_histogram(up_o, R.upscale(frame[:360, :640], (1280, 720), False, "strict")),
                         "restatement_vs_contract": R.lsb_histogram(up_o, R.upscale(frame[:360, :640], (1280, 720), False, "contract"))}
    sh_o = oracle.sharpen(frame, 0.8)
    rec["rcas_0.8"] = {"restatement_vs_strict": R.lsb_histogram(sh_o, R.sharpen(frame, 0.8, "strict")),
                       "restatement_vs_contract": R.lsb_histogram(sh_o, R.sharpen(frame, 0.8, "contract"))}
    _check("rcas 720p", sh_o, R.sharpen(frame, 0.8, "strict"), R.sharpen(frame, 0.8, "contract"), exact=True)
    if os.path.isdir("/root/reference"):  # only the build container refreshes the tracked artifact
        with open(os.path.join(ROOT, "profiles", "r02_parity_fsr_ref_cpu.json"), "w") as f:
            json.dump(rec, f, indent=1)


def test_independent_numpy_restatement_equals_reference_strict_build(oracle):
    """oracle/easu_numpy.py — written from the FSR.cl text alone, vectorised NumPy float32 (no FMA), sharing no code with
    easu_ref.c or with the shim build — reproduces the reference's STRICT build bit for bit on every fixture, and
    brackets easu_ref.c (which follows the contracted arithmetic) within 1 LSB on < 1e-3 of the bytes."""
    from oracle import easu_numpy as N
    k = 0
    for t in G["transforms"]:
        for yuv in (False, True):
            out = N.remap_homography(G["src"], t, (255, 0, 255), yuv)
            assert (out == G["homography_strict"][k]).all(), f"transform {k}: {R.lsb_histogram(out, G['homography_strict'][k])}"
            h = R.lsb_histogram(out, oracle.remap_homography(G["src"], t, (255, 0, 255), yuv))
            assert h[2] == 0 and h[3] == 0 and h[1] < 1e-3 * out.size, h
            k += 1
