"""Generates the committed golden fixtures under tests/golden/ (run from the repo root).

  quantise_cases.npz   (v, bound, q): the CAST macro of /root/reference/src/common.cpp:6 evaluated by an
                       INDEPENDENT restatement in Python floats (IEEE double) with round-half-to-even
                       (Python's round()), i.e. not by oracle/quantise_oracle.c.
  farneback_cv2_*.npz  cv2.calcOpticalFlowFarneback(a,b,None,0.5,5,13,10,5,1.1,0) — real OpenCV CPU code
                       with the reference's Farneback defaults — on the seeded synthetic pairs, stored at
                       stride 4 (+ full-field means) so the oracle stays pinned even where cv2 is absent.
  tvl1_oracle_256.npz  frozen output of the CPU oracle on the config-2 pair (regression pin only:
                       no external TV-L1 implementation exists in this image — parity unpinned).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))


def cast_py(v, bound):
    L, H = float(-bound), float(bound)
    v = float(np.float32(v))
    if v > H:
        return 255
    if v < L:
        return 0
    return int(round(255 * (v - L) / (H - L)))  # Python round() is round-half-to-even == cvRound


def main():
    import cv2
    from denseflow_b200 import synth
    from oracle import pyoracle as O

    vs, bs, qs = [], [], []
    rng = np.random.default_rng(123)
    for bound in (1, 7, 20, 32, 64):
        k = np.arange(255)
        ties = (-bound + (k + 0.5) * (2 * bound) / 255).astype(np.float32)
        special = np.array([-bound, bound, np.nextafter(np.float32(-bound), np.float32(-1e9)),
                            np.nextafter(np.float32(bound), np.float32(1e9)), 0.0, -0.0, 1e30, -1e30, 0.5, 1.5, 2.5],
                           np.float32)
        rnd = (rng.standard_normal(2000) * bound * 0.7).astype(np.float32)
        v = np.concatenate([ties, special, rnd])
        vs.append(v)
        bs.append(np.full(v.shape, bound, np.int32))
        qs.append(np.array([cast_py(x, bound) for x in v], np.uint8))
    np.savez_compressed(os.path.join(HERE, "quantise_cases.npz"), v=np.concatenate(vs), bound=np.concatenate(bs),
                        q=np.concatenate(qs))

    for (h, w, seed) in [(256, 256, 0), (256, 340, 100)]:
        a, b, _ = synth.pair(h, w, seed)
        f = cv2.calcOpticalFlowFarneback(a, b, None, 0.5, 5, 13, 10, 5, 1.1, 0)
        np.savez_compressed(os.path.join(HERE, "farneback_cv2_%dx%d.npz" % (w, h)), flow_s4=f[::4, ::4].copy(),
                            mean=f.reshape(-1, 2).mean(0), sha_a=synth.sha1(a), sha_b=synth.sha1(b),
                            cv2_version=cv2.__version__)

    a, b, _ = synth.pair(256, 256, 0)
    flow, log = O.tvl1_calc(a, b, return_iters=True)
    np.savez_compressed(os.path.join(HERE, "tvl1_oracle_256.npz"), flow_s4=flow[::4, ::4].copy(),
                        mean=flow.reshape(-1, 2).mean(0), iters=log, sha_a=synth.sha1(a), sha_b=synth.sha1(b))
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
