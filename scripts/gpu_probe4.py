"""Throughput at the other BASELINE configs: tvl1 340x256 (lanes sweep), farn 1280x720."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import denseflow_b200 as d
from denseflow_b200 import synth

def run(alg, W, H, N, sets, seed):
    fr = synth.stream(H, W, N, seed=seed)
    dev = torch.from_numpy(fr).cuda()
    out = torch.empty((N - 1, H, W, 2), dtype=torch.float32, device="cuda")
    for kv in sets:
        e = d.create(alg, 0, W, H)
        for k, v in kv.items(): e.set(k, v)
        e.calc_batch_device(dev, 1, out); torch.cuda.synchronize()
        e.reset_counters()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(); e.calc_batch_device(dev, 1, out); ev1.record(); torch.cuda.synchronize()
        dt = ev0.elapsed_time(ev1) / 1e3
        c = e.counters()
        print("%s %dx%d %s: %.3f ms/pair (%.1f pairs/s) launches/pair %.1f px-iters/pair %.1fM" % (alg, W, H, kv, dt / (N - 1) * 1e3, (N - 1) / dt, c["kernel_launches"] / (N - 1), c["pixel_iters"] / (N - 1) / 1e6))
        e.release()

run("tvl1", 340, 256, 64, [{"lanes": 1}, {"lanes": 4}, {"lanes": 8}, {"lanes": 0}, {"fused": 0}], 100)
run("tvl1", 256, 256, 33, [{"lanes": 0}], 0)
run("farn", 1280, 720, 17, [{}], 2)
run("farn", 1920, 1080, 9, [{}], 1)
