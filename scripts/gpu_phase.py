"""Phase breakdown (in-kernel %globaltimer profiler) of the fused engine at a given size."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import denseflow_b200 as d
from denseflow_b200 import synth
W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1920, 1080)
ks = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [8]
lanes = int(sys.argv[4]) if len(sys.argv) > 4 else 1
variant = sys.argv[5] if len(sys.argv) > 5 else "default"
fr = synth.stream(H, W, 4, seed=1)
dev = torch.from_numpy(fr).cuda()
for k in ks:
    e = d.OpticalFlowDual_TVL1.create(0, W, H, variant)
    e.set("fused_k", k); e.set("lanes", lanes)
    out = e.calc_batch_device(dev, 1); torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(); out = e.calc_batch_device(dev, 1); ev1.record(); torch.cuda.synchronize()
    p = e.phase_ns(); it, sizes = e.tvl1_stats()
    tot = sum(p[k_] for k_ in ("level_start", "warp", "tiles", "barrier", "upsample_merge"))
    print("k=%d  %.3f ms/pair (events, 3 pairs)  | last pair CTA0 total %.3f ms: level_start %.3f warp %.3f tiles %.3f barrier %.3f up/merge %.3f" % (
        k, ev0.elapsed_time(ev1) / 3, tot / 1e6, p["level_start"] / 1e6, p["warp"] / 1e6, p["tiles"] / 1e6, p["barrier"] / 1e6, p["upsample_merge"] / 1e6))
    print("   tile phases (CTA0 thread0): load %.3f iterate %.3f store %.3f ms" % (p["tile_load"] / 1e6, p["tile_iter"] / 1e6, p["tile_store"] / 1e6))
    for s in range(len(sizes)):
        ch = p["chunks_per_scale"][s]
        print("   scale %d %4dx%-4d iters %-22s tile time %.3f ms, %3d chunks, %.1f us/chunk" % (s, sizes[s][0], sizes[s][1], it[s].tolist(), p["tiles_per_scale"][s] / 1e6, ch, p["tiles_per_scale"][s] / 1e3 / max(ch, 1)))
    e.release()
