"""Times the pipelined NV12 path (Stream.stream_obs) against the packed-BGR pipelined path on the same clip."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch, cv2
import livevisionkit_b200 as L
from tools.synth import Clip
w, h, n = 1920, 1080, 140
clip = Clip("1080p", "shake", frames=n)
bgr = torch.empty((n + 3, h, w, 3), dtype=torch.uint8).pin_memory()
nv = torch.empty((n + 3, h * 3 // 2, w), dtype=torch.uint8).pin_memory()
for i in range(n):
    f = clip[i]; bgr[i].copy_(torch.from_numpy(f))
    i420 = cv2.cvtColor(f, cv2.COLOR_BGR2YUV_I420); d = nv[i].numpy(); d[:h] = i420[:h]
    uv = d[h:].reshape(h // 2, w // 2, 2); uv[:, :, 0] = i420[h:h + h // 4].reshape(h // 2, w // 2); uv[:, :, 1] = i420[h + h // 4:].reshape(h // 2, w // 2)
def obs(t, ts):
    a = t.numpy(); return L.ObsFrame("NV12", w, h, [a[:h], a[h:]], timestamp=ts)
S = L.StabilizationFilterSettings.obs_homography_preset()
s1 = L.Stream(S, 0)
srcs = [obs(nv[i], i) for i in range(n)]; outs = [obs(nv[n + i], 0) for i in range(3)]
s1.stream_obs(srcs[:20], lambda o: False, outs)
t0 = time.perf_counter(); k = s1.stream_obs(srcs[20:], lambda o: False, outs); s1.sync(); t1 = time.perf_counter()
print("nv12 pipelined: %.1f us/frame (%d outputs)" % (1e6 * (t1 - t0) / (n - 20), k))
s1.close()
s2 = L.Stream(S, 0)
fr = [L.VideoFrame(L.FrameRef(bgr[i]), i, L.BGR) for i in range(n)]; po = [L.FrameRef(bgr[n + i]) for i in range(3)]
s2(fr[:20], lambda v: False, po)
t0 = time.perf_counter(); k = s2(fr[20:], lambda v: False, po); s2.sync(); t1 = time.perf_counter()
print("bgr  pipelined: %.1f us/frame (%d outputs)" % (1e6 * (t1 - t0) / (n - 20), k))
s2.close()
# variant A: no look-ahead announcement (submit_obs_async uploads inside the call)
s3 = L.Stream(S, 0)
pend = []
for i in range(20):
    r, t = s3.submit_obs_async(srcs[i], outs[i % 3])
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(20, n):
    r, t = s3.submit_obs_async(srcs[i], outs[i % 3])
    if t: pend.append(t)
    if len(pend) > 2: s3.wait_output(pend.pop(0))
s3.sync(); t1 = time.perf_counter()
print("nv12 no-announce: %.1f us/frame" % (1e6 * (t1 - t0) / (n - 20)))
s3.close()
# variant B: synchronous submit_obs
s4 = L.Stream(S, 0)
for i in range(20): s4.submit_obs(srcs[i], outs[0])
t0 = time.perf_counter()
for i in range(20, n): s4.submit_obs(srcs[i], outs[0])
s4.sync(); t1 = time.perf_counter()
print("nv12 synchronous submit_obs: %.1f us/frame" % (1e6 * (t1 - t0) / (n - 20)))
s4.close()
