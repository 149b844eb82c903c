"""CPU suite: the oracle against golden vectors / real OpenCV CPU code / analytic properties.

The reference has no tests and no golden vectors, so the pins are:
  quantiser  -> independent Python restatement of the CAST macro (tests/golden/quantise_cases.npz)
  Farneback  -> cv2.calcOpticalFlowFarneback, live when cv2 is importable, else the committed samples
  TV-L1      -> PARITY UNPINNED: frozen regression output + analytic ground-truth sanity only
"""
import os

import numpy as np
import pytest

from denseflow_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_synthetic_pair_is_the_surveyed_one():
    a, b, _ = synth.pair(256, 256, 0)
    assert synth.sha1(a) == "4ecb584297a6322225a5503fb4ef091037ad3223"
    assert synth.sha1(b) == "08cf74f14c60747ecd56a7a94c5cc5919c63c039"
    assert a[0, :4].tolist() == [101, 120, 125, 125]


def test_quantiser_matches_cast_macro_golden(oracle):
    g = np.load(os.path.join(GOLD, "quantise_cases.npz"))
    for bound in np.unique(g["bound"]):
        m = g["bound"] == bound
        v = g["v"][m]
        flow = np.stack([v, v[::-1]], -1).reshape(1, -1, 2).astype(np.float32)
        qx, qy = oracle.quantise(flow, int(bound))
        assert np.array_equal(qx.ravel(), g["q"][m])
        assert np.array_equal(qy.ravel(), g["q"][m][::-1])


def test_quantiser_known_answers(oracle):
    # cvRound ties to even: 0.5 -> 0, 1.5 -> 2, 2.5 -> 2, 127.5 -> 128; zero flow -> 128 at any bound
    b = 20
    for k, want in [(0, 0), (1, 2), (2, 2), (127, 128)]:
        v = np.float64(-b + (k + 0.5) * (2 * b) / 255)
        # only exactly representable ties are ties in fp32; build them in double and check the fp32 neighbour rounds consistently
        f = np.float32(v)
        q = int(round(255 * (float(f) + b) / (2 * b)))
        flow = np.array([[[f, 0.0]]], np.float32)
        qx, qy = oracle.quantise(flow, b)
        assert qx[0, 0] == q and qy[0, 0] == 128
        del want
    flow = np.array([[[25.0, -25.0]]], np.float32)
    qx, qy = oracle.quantise(flow, b)
    assert (qx[0, 0], qy[0, 0]) == (255, 0)


def test_farneback_constants_match_survey(oracle):
    g = np.zeros(6, np.float32); xg = np.zeros(6, np.float32); xxg = np.zeros(6, np.float32); ig = np.zeros(4, np.float32)
    oracle.lib().orc_farn_poly_constants(5, 1.1, g, xg, xxg, ig)
    np.testing.assert_allclose(g, [0.36267489, 0.23991476, 0.069450498, 0.0087977722, 0.00048769583, 1.1830532e-05], rtol=2e-6)
    np.testing.assert_allclose(ig, [0.826452292, -0.413263275, 0.341542384, 0.683023397], rtol=2e-6)
    assert [l[2] for l in oracle.farn_levels(1920, 1080)][::-1][:5] == [3, 3, 9, 19, 39]
    assert [l[:2] for l in oracle.farn_levels(1280, 720)] == [(80, 45), (160, 90), (320, 180), (640, 360), (1280, 720)]
    assert [l[:2] for l in oracle.farn_levels(256, 256)] == [(32, 32), (64, 64), (128, 128), (256, 256)]


@pytest.mark.parametrize("hw", [(256, 256, 0), (256, 340, 100)])
def test_farneback_oracle_pinned_to_opencv_cpu(oracle, hw):
    h, w, seed = hw
    a, b, _ = synth.pair(h, w, seed)
    gold = np.load(os.path.join(GOLD, "farneback_cv2_%dx%d.npz" % (w, h)))
    assert synth.sha1(a) == str(gold["sha_a"]) and synth.sha1(b) == str(gold["sha_b"])
    mine = oracle.farn_calc(a, b, oracle.farn_params(resize_convention=oracle.RESIZE_HALF_PIXEL))
    # committed samples of real OpenCV output
    assert synth.aee(mine[::4, ::4], gold["flow_s4"]) < 5e-6
    assert np.abs(mine[::4, ::4] - gold["flow_s4"]).max() < 1e-4
    np.testing.assert_allclose(mine.reshape(-1, 2).mean(0), gold["mean"], atol=1e-5)
    try:
        import cv2
    except ImportError:
        return
    live = cv2.calcOpticalFlowFarneback(a, b, None, 0.5, 5, 13, 10, 5, 1.1, 0)
    assert synth.aee(mine, live) < 5e-6
    # the CUDA resize convention (what GPU parity uses) moves the result only marginally on small motion
    cuda_conv = oracle.farn_calc(a, b)
    assert synth.aee(cuda_conv, live) < 1e-3


def test_tvl1_level_sizes(oracle):
    assert oracle.tvl1_level_sizes(1920, 1080) == [(1920, 1080), (1536, 864), (1229, 691), (983, 553), (786, 442)]
    assert oracle.tvl1_level_sizes(256, 256) == [(256, 256), (205, 205), (164, 164), (131, 131), (105, 105)]
    assert oracle.tvl1_level_sizes(340, 256) == [(340, 256), (272, 205), (218, 164), (174, 131), (139, 105)]
    assert len(oracle.tvl1_level_sizes(24, 40)) == 2  # 19x32 kept, 15x26 dropped (cols < 16)


def test_tvl1_oracle_regression_and_ground_truth(oracle, pair256):
    a, b, gt = pair256
    flow, log = oracle.tvl1_calc(a, b, return_iters=True)
    gold = np.load(os.path.join(GOLD, "tvl1_oracle_256.npz"))
    assert np.array_equal(log, gold["iters"])
    assert log.sum() == 1244  # SURVEY A.6 self-consistency figure
    assert np.abs(flow[::4, ::4] - gold["flow_s4"]).max() < 1e-4
    assert 0.05 < synth.aee(flow, gt) < 0.1  # analytic flow: 0.083 px in the survey's prototype
    assert synth.aee(flow[16:-16, 16:-16], gt[16:-16, 16:-16]) < 0.07


def test_tvl1_oracle_properties(oracle):
    a, _, _ = synth.pair(96, 128, 4)
    z, log = oracle.tvl1_calc(a, a, return_iters=True)
    assert np.abs(z).max() == 0.0 and (log == 2).all()  # identical frames: zero flow, minimum 2 iterations per warp
    # pure translation of a periodic-free texture: mean flow recovers the shift
    a2, b2, gt = synth.pair(128, 128, 9)
    f = oracle.tvl1_calc(a2, b2)
    assert abs(f[..., 0].mean() - gt[..., 0].mean()) < 0.05 and abs(f[..., 1].mean() - gt[..., 1].mean()) < 0.05


def test_tvl1_blocks_match_reference_formulas(oracle):
    """Known-answer checks of the building blocks on hand-computable inputs."""
    L = oracle.lib()
    # centred gradient of a ramp: 1 inside, 0.5 on the clamped borders
    w, h = 8, 5
    ramp = np.tile(np.arange(w, dtype=np.float32), (h, 1))
    dx = np.zeros_like(ramp); dy = np.zeros_like(ramp)
    L.orc_centered_gradient(ramp, w, h, dx, dy)
    assert np.all(dx[:, 1:-1] == 1.0) and np.all(dx[:, 0] == 0.5) and np.all(dx[:, -1] == 0.5) and np.all(dy == 0)
    # CUDA-convention resize of a ramp by 0.8: dst(x) = 1.25 x exactly (no half-pixel offset)
    dst = np.zeros((4, 6), np.float32)
    L.orc_resize_linear(ramp, w, h, dst, 6, 4, 1.25, 1.25, 0)
    np.testing.assert_allclose(dst[0], 1.25 * np.arange(6), rtol=0, atol=1e-6)
    # warp with zero flow reproduces I1 and its gradients exactly (bicubic weights collapse to the centre tap)
    rng = np.random.default_rng(0)
    I0 = rng.random((h, w)).astype(np.float32); I1 = rng.random((h, w)).astype(np.float32)
    I1x = np.zeros_like(I1); I1y = np.zeros_like(I1)
    L.orc_centered_gradient(I1, w, h, I1x, I1y)
    z = np.zeros_like(I1)
    outs = [np.zeros_like(I1) for _ in range(5)]
    L.orc_tvl1_warp_backward(I0, I1, I1x, I1y, z, z, w, h, *outs)
    np.testing.assert_allclose(outs[0], I1, atol=1e-6)
    np.testing.assert_allclose(outs[1], I1x, atol=1e-6)
    np.testing.assert_allclose(outs[4], I1 - I0, atol=1e-6)
    # dual step: p stays inside the unit ball scaled by ... (|p| <= 1 after projection-like update from p = 0)
    u1 = rng.standard_normal((h, w)).astype(np.float32); u2 = rng.standard_normal((h, w)).astype(np.float32)
    p = [np.zeros((h, w), np.float32) for _ in range(4)]
    L.orc_tvl1_estimate_dual(u1, u2, *p, w, h, np.float32(0.25 / 0.3))
    assert np.all(np.hypot(p[0], p[1]) < 1.0) and np.all(p[0][:, -1] == 0) and np.all(p[1][-1, :] == 0)


def test_tvl1_c_oracle_agrees_with_independent_numpy_restatement(oracle):
    """Two restatements of SURVEY Appendix A written separately (C loops vs vectorised numpy) must agree: same
    iteration schedule, flows equal to fp32 rounding.  (Self-consistency only — TV-L1 parity stays unpinned.)"""
    from oracle import tvl1_numpy as N
    a, b, _ = synth.pair(80, 96, 6)
    prm = oracle.tvl1_params(nscales=3, warps=3, iterations=60)
    fc, lc = oracle.tvl1_calc(a, b, prm, return_iters=True)
    fn, ln = N.calc(a, b, nscales=3, warps=3, iterations=60)
    assert np.array_equal(lc[:3, :3], ln[:3, :3])
    assert np.abs(fc - fn).max() < 5e-4 and synth.aee(fc, fn) < 2e-5
    # a size whose pyramid drops a level and a non-converging pair that runs into the cap
    a2, b2 = synth.noise_pair(40, 56, 3)
    prm = oracle.tvl1_params(nscales=5, warps=2, iterations=25)
    fc, lc = oracle.tvl1_calc(a2, b2, prm, return_iters=True)
    fn, ln = N.calc(a2, b2, nscales=5, warps=2, iterations=25)
    assert np.array_equal(lc[:, :2], ln[:, :2]) and lc.max() == 25
    assert synth.aee(fc, fn) < 1e-3
