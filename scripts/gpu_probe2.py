"""1080p timing of the fused vs unfused engine (device-resident frames)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import denseflow_b200 as d
from denseflow_b200 import synth

H, W = 1080, 1920
N = 5
fr = synth.stream(H, W, N, seed=1)
dev = torch.from_numpy(fr).cuda()
out = torch.empty((N - 1, H, W, 2), dtype=torch.float32, device="cuda")
ref = None
for fused, k in [(0, 0), (1, 8), (1, 4), (1, 6), (1, 2), (1, 1)]:
    e = d.OpticalFlowDual_TVL1.create(0, W, H)
    e.set("fused", fused)
    if fused: e.set("fused_k", k)
    for _ in range(2):
        e.calc_batch_device(dev, 1, out); torch.cuda.synchronize()
    e.reset_counters()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(); e.calc_batch_device(dev, 1, out); ev1.record(); torch.cuda.synchronize()
    dt = ev0.elapsed_time(ev1) / 1e3
    c = e.counters(); it, sizes = e.tvl1_stats()
    res = out.cpu().numpy()
    if ref is None: ref = res.copy()
    print("fused=%d k=%d: %.3f ms/pair (%.1f pairs/s), launches/pair %.0f, px-iters/pair %.1fM, total iters %d, AEE vs unfused %.2e, GB/s@64B %.0f" % (
        fused, k, dt / (N - 1) * 1e3, (N - 1) / dt, c["kernel_launches"] / (N - 1), c["pixel_iters"] / (N - 1) / 1e6, it.sum(), synth.aee(res, ref), c["pixel_iters"] * 64 / dt / 1e9))
    e.release()
