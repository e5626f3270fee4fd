import sys; sys.path.insert(0,'.')
import numpy as np
import livevisionkit_b200 as L
from livevisionkit_b200 import _capi as K
from oracle import lvk_oracle as O
from tools.synth import Clip
clip = Clip("1080p","shake",frames=8)
so = O.StabilizationSettings.obs_homography_preset(); sg = L.StabilizationFilterSettings.obs_homography_preset()
ref = O.StabilizationFilter(so); flt = L.StabilizationFilter(sg,0)
ref.restart(); flt.restart(); flt.stream.set_debug_capture(True)
s=flt.stream
for i in range(6):
    f=clip[i]
    ref.apply(f,O.BGR,i); tr=ref.trace
    flt.apply(L.VideoFrame(f,i,L.BGR))
    if "lk_in" in tr:
        kd = s.debug_fetch(K.DBG_DETECTED, np.dtype([("x","f4"),("y","f4"),("r","f4"),("c","i4")]))
        gin = np.stack([kd["x"],kd["y"]],axis=1)
        din = np.abs(gin-tr["lk_in"]).max(axis=1)
        mt = s.debug_fetch(K.DBG_LK_MATCHED, np.float32).reshape(-1,2)
        dout = np.abs(mt-tr["lk_out"]).max(axis=1)
        # direct stage-level LK on the oracle's exact inputs
        prev_det = prev; cur_det = tr["det"]
        m2, st2 = s.lk_track(prev_det, cur_det, tr["lk_in"])
        d2 = np.abs(m2-tr["lk_out"]).max(axis=1)
        ages = kd["c"]
        print(f"frame {i}: n={len(gin)} in: max {din.max():.2e} nonzero {int((din>0).sum())} | out: max {dout.max():.2e} nonzero {int((dout>0).sum())} | stage-LK on oracle inputs: max {d2.max():.2e} nonzero {int((d2>0).sum())} | aged {int((ages>0).sum())}; nonzero-out among aged {int(((dout>0)&(ages>0)).sum())} among new {int(((dout>0)&(ages==0)).sum())}")
        bad=np.argsort(-d2)[:3]
        for b in bad: print("   worst stage:", tr["lk_in"][b], tr["lk_out"][b], m2[b], "status", tr["lk_status"][b], st2[b])
    prev = tr["det"]
