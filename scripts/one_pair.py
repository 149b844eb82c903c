"""Runs a few pairs through the fused engine (target for ncu). args: W H nframes lanes variant"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import denseflow_b200 as d
from denseflow_b200 import synth
W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1920, 1080)
n = int(sys.argv[3]) if len(sys.argv) > 3 else 4
lanes = int(sys.argv[4]) if len(sys.argv) > 4 else 1
variant = sys.argv[5] if len(sys.argv) > 5 else "default"
fr = synth.stream(H, W, n, seed=1)
dev = torch.from_numpy(fr).cuda()
e = d.OpticalFlowDual_TVL1.create(0, W, H, variant)
e.set("lanes", lanes)
out = e.calc_batch_device(dev, 1)
torch.cuda.synchronize()
print("done", out.shape, e.counters())
