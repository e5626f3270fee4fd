"""Per-stage device time of the pipelined host path (uploads/downloads in flight) next to the device-resident path.

Answers "what does the tracking chain cost while the copy engines saturate PCIe?".  Profiling mode = eager launches
with CUDA events around every stage.  Diagnostic only.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

import torch  # noqa: E402

import livevisionkit_b200 as L  # noqa: E402
from tools.synth import Clip  # noqa: E402

N = 160
clip = Clip("1080p", "shake", frames=N, seed=3)
host = [torch.from_numpy(clip[i]).pin_memory() for i in range(N)]
outs = [torch.empty_like(host[0]).pin_memory() for _ in range(3)]
settings = L.StabilizationFilterSettings.obs_homography_preset()


def show(tag, stream):
    totals, counts = stream.stage_totals_us(reset=True)
    print(tag, " ".join(f"{k}={totals[k] / max(counts[k], 1):.1f}" for k in totals))


# device resident
dev = [f.cuda() for f in host]
dout = [torch.empty_like(dev[0]) for _ in range(4)]
flt = L.StabilizationFilter(settings, device=0)
flt.stream.set_profiling(True)
for i in range(N):
    flt.stream.submit(dev[i], dout[i % 4], L.BGR, i)
    if i == 40:
        flt.stream.stage_totals_us(reset=True)
show("device-resident:", flt.stream)
flt.stream.close()

# pipelined, host buffers
flt = L.StabilizationFilter(settings, device=0)
flt.stream.set_profiling(True)
frames = [L.VideoFrame(host[i], i, L.BGR) for i in range(N)]
flt.stream(frames[:40], lambda vf: False, outs)
flt.stream.stage_totals_us(reset=True)
flt.stream(frames[40:], lambda vf: False, outs)
show("pipelined host :", flt.stream)
flt.stream.close()

# synchronous apply, host buffers
flt = L.StabilizationFilter(settings, device=0)
flt.stream.set_profiling(True)
for i in range(N):
    flt.apply(frames[i], output=outs[i % 3])
    if i == 40:
        flt.stream.stage_totals_us(reset=True)
show("sync apply host:", flt.stream)
flt.stream.close()
