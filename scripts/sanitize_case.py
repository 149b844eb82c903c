"""Small end-to-end case for compute-sanitizer (memcheck): tvl1 fused (lanes, TMA, odd sizes) + farneback + preproc."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import denseflow_b200 as d
from denseflow_b200 import synth
for (w, h) in [(131, 97), (340, 256)]:
    fr = synth.stream(h, w, 5, seed=3)
    e = d.OpticalFlowDual_TVL1.create(0, w, h)
    e.set("iterations", 30)
    out = e.calc_batch(list(fr), step=1)
    q = e.calc_batch(list(fr), step=-2, bound=20)
    e.set("fused", 0); o2 = e.calc(fr[0], fr[1])
    print("tvl1", w, h, float(np.abs(out).max()), np.isfinite(out).all())
    f = d.FarnebackOpticalFlow.create(0, w, h)
    of = f.calc_batch(list(fr), step=1)
    print("farn", w, h, float(np.abs(of).max()), np.isfinite(of).all())
    bgr = torch.from_numpy(np.stack([fr[0]] * 3, -1).copy()).cuda()
    g = e.bgr_to_gray_device(bgr); r = e.resize_gray_device(g, 77, 55); torch.cuda.synchronize()
print("ok")
