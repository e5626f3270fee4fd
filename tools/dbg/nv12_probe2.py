import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch, cv2
import livevisionkit_b200 as L
from livevisionkit_b200 import _capi as K
from tools.synth import Clip
w, h, n = 1920, 1080, 40
clip = Clip("1080p", "shake", frames=n)
nv = torch.empty((n + 3, h * 3 // 2, w), dtype=torch.uint8).pin_memory()
for i in range(n):
    f = clip[i]
    i420 = cv2.cvtColor(f, cv2.COLOR_BGR2YUV_I420); d = nv[i].numpy(); d[:h] = i420[:h]
    uv = d[h:].reshape(h // 2, w // 2, 2); uv[:, :, 0] = i420[h:h + h // 4].reshape(h // 2, w // 2); uv[:, :, 1] = i420[h + h // 4:].reshape(h // 2, w // 2)
def obs(t, ts):
    a = t.numpy(); return L.ObsFrame("NV12", w, h, [a[:h], a[h:]], timestamp=ts)
S = L.StabilizationFilterSettings.obs_homography_preset()
s = L.Stream(S, 0)
srcs = [obs(nv[i], i) for i in range(n)]; outs = [obs(nv[n + i], 0) for i in range(3)]
for i in range(n):
    if i + 1 < n: s.prefetch_obs(srcs[i + 1])
    r, t = s.submit_obs_async(srcs[i], outs[i % 3])
    if t: s.wait_output(t)
    st = s.stage_times_us()
    print(i, "motion", r.has_motion, "stab %.3f q %.3f trust %.3f" % (r.tracking_stability, r.scene_quality, r.trust_factor), {k: round(v, 1) for k, v in st.items()})
