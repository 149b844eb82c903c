"""One Farneback pair with the TMA-staged iteration kernel (target for compute-sanitizer). args: variant"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import denseflow_b200 as d
from denseflow_b200 import synth
variant = sys.argv[1] if len(sys.argv) > 1 else "default"
a, b, _ = synth.pair(256, 256, 0)
e0 = d.FarnebackOpticalFlow.create(0, 256, 256, variant); e0.set("use_tma", 0)
ref = e0.calc(a, b)
e1 = d.FarnebackOpticalFlow.create(0, 256, 256, variant)
out = e1.calc(a, b)
print(variant, "tma == ldg:", np.array_equal(ref, out), "max diff", float(np.abs(ref - out).max()))
