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
    png, bounds = e.flow_to_png_image_device(torch.from_numpy(out[0]).cuda())
    chain = e.process_bgr_batch([np.stack([x] * 3, -1) for x in fr], step=1, bound=20, new_size=(64, 48))
    print("png bounds", bounds, "chain jpegs", len(chain), len(chain[0][0]))
from denseflow_b200 import listrun
clips = [list(synth.stream(96, 128, n, seed=9 + n)) for n in (5, 1, 7, 3)]
print("list", listrun.run_list("tvl1", [0, 0], clips, step=1, bound=20, chunk_flows=3)["flows"])
print("ok")
