"""First GPU probe: 1080p timing of the unfused engine + per-kernel event timing."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import denseflow_b200 as d
from denseflow_b200 import synth

H, W = 1080, 1920
fr = synth.stream(H, W, 4, seed=1)
e = d.OpticalFlowDual_TVL1.create(0, W, H)
e.set("fused", 0)
dev = torch.from_numpy(fr).cuda()
out = torch.empty((3, H, W, 2), dtype=torch.float32, device="cuda")
for _ in range(2):
    e.calc_batch_device(dev, 1, out); torch.cuda.synchronize()
e.reset_counters()
t = time.time(); e.calc_batch_device(dev, 1, out); torch.cuda.synchronize(); dt = time.time() - t
c = e.counters(); it, sizes = e.tvl1_stats()
print("unfused 1080p: %.2f ms/pair, launches/pair %d, px-iters/pair %.1fM, iters %s" % (dt / 3 * 1e3, c["kernel_launches"] / 3, c["pixel_iters"] / 3 / 1e6, it[::-1].tolist()))
print("eff GB/s (88 B/px-iter): %.0f" % (c["pixel_iters"] * 88 / dt / 1e9))
# host path
flows = e.calc_batch(list(fr), 1)
t = time.time(); flows = e.calc_batch(list(fr), 1); dt = time.time() - t
print("host batch path (pageable): %.2f ms/pair" % (dt / 3 * 1e3))
print("mean flow", flows[0][..., 0].mean(), flows[0][..., 1].mean())
