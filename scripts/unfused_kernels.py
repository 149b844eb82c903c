"""Target for ncu: the two stand-alone inner-loop kernels of the unfused schedule at W x H on an L2-busting rotation of
operand sets (dfb_debug_time_kernel).  args: W H"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import denseflow_b200 as d
W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1920, 1080)
e = d.OpticalFlowDual_TVL1.create(0, W, H)
for name in ("estimate_u", "estimate_dual"):
    print(name, e.debug_time_kernel(name, W, H, sets=6, reps=6), "ms")
