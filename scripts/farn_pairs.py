import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import denseflow_b200 as d
from denseflow_b200 import synth
W, H, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
fr = synth.stream(H, W, n, seed=2)
dev = torch.from_numpy(fr).cuda()
e = d.FarnebackOpticalFlow.create(0, W, H)
out = e.calc_batch_device(dev, 1); torch.cuda.synchronize()
out = e.calc_batch_device(dev, 1); torch.cuda.synchronize()
print("done", e.counters())
