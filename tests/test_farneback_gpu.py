"""GPU parity: CUDA Farneback (through the C ABI) vs the CPU oracle (CUDA resize convention), which itself
is pinned to OpenCV's CPU Farneback (tests/test_oracle_cpu.py)."""
import numpy as np
import pytest

from denseflow_b200 import synth

pytestmark = pytest.mark.gpu
AEE_TOL = 0.01


def _engine(w, h, variant="default"):
    import denseflow_b200 as d
    return d.FarnebackOpticalFlow.create(0, w, h, variant)


@pytest.mark.parametrize("variant", ["strict", "default"])
@pytest.mark.parametrize("shape", [(256, 256), (256, 340), (97, 131), (64, 64), (360, 640)])
def test_farneback_matches_oracle(oracle, shape, variant):
    h, w = shape
    a, b, gt = synth.pair(h, w, 0)
    ref = oracle.farn_calc(a, b)
    flow = _engine(w, h, variant).calc(a, b)
    aee = synth.aee(flow, ref)
    print(shape, variant, "AEE vs oracle", aee, "max", np.abs(flow - ref).max())
    assert np.isfinite(flow).all()
    assert aee <= (1e-4 if variant == "strict" else AEE_TOL)
    if min(h, w) >= 128:
        assert synth.aee(flow, gt) < 0.15  # sanity vs analytic flow (CPU Farneback gets 0.04-0.07 px here)


def test_farneback_small_frame_level_cropping(oracle):
    """Fewer than 5 levels when W*0.5^k < 32."""
    a, b, _ = synth.pair(96, 160, 2)
    assert len(oracle.farn_levels(160, 96)) == 2  # 160x96 -> 80x48 kept, 40x24 dropped (24 < 32)
    assert len(oracle.farn_levels(80, 48)) == 1
    assert synth.aee(_engine(160, 96).calc(a, b), oracle.farn_calc(a, b)) <= AEE_TOL
    a, b, _ = synth.pair(48, 80, 2)
    assert synth.aee(_engine(80, 48).calc(a, b), oracle.farn_calc(a, b)) <= AEE_TOL


def test_farneback_batch_and_quantise(oracle):
    fr = synth.stream(120, 160, 5, seed=21)
    e = _engine(160, 120)
    flows = e.calc_batch(list(fr), step=1)
    assert flows.shape == (4, 120, 160, 2)
    for i in range(4):
        assert np.array_equal(flows[i], e.calc(fr[i], fr[i + 1]))
    assert synth.aee(flows[1], oracle.farn_calc(fr[1], fr[2])) <= AEE_TOL
    qx, qy = e.calc_batch(list(fr), step=2, bound=20)
    f2 = e.calc_batch(list(fr), step=2)
    for i in range(3):
        ox, oy = oracle.quantise(f2[i], 20)
        assert np.array_equal(qx[i], ox) and np.array_equal(qy[i], oy)


def test_farneback_720p_config(oracle):
    """BASELINE.json configs[3] size."""
    fr = synth.stream(720, 1280, 2, seed=2)
    ref = oracle.farn_calc(fr[0], fr[1])
    flow = _engine(1280, 720).calc(fr[0], fr[1])
    assert synth.aee(flow, ref) <= AEE_TOL


def test_farneback_1080p_six_levels(oracle):
    """1920x1080 keeps the k = 5 level (60 x 34): smoothSize 79, the widest Gaussian of the path."""
    assert oracle.farn_levels(1920, 1080)[0][:3] == (60, 34, 79)
    fr = synth.stream(1080, 1920, 2, seed=1)
    ref = oracle.farn_calc(fr[0], fr[1])
    flow = _engine(1920, 1080).calc(fr[0], fr[1])
    assert synth.aee(flow, ref) <= AEE_TOL


@pytest.mark.parametrize("shape", [(97, 131), (360, 640), (720, 1280)])
def test_tma_staged_iteration_is_bit_identical(shape):
    """The persistent TMA-staged, double-buffered iteration kernel (default) and the LDG-staged one (use_tma = 0) read the same
    windows: zero-filled out-of-image entries are replaced by the index-clamped value on border tiles.  Batches of several
    pairs (the tile list walks pairs x rows x columns), odd sizes, partial tiles."""
    h, w = shape
    fr = synth.stream(h, w, 7, seed=40)
    e0 = _engine(w, h)
    e0.set("use_tma", 0)
    e1 = _engine(w, h)
    assert e1.get("use_tma") == 1
    a = e0.calc_batch(list(fr), step=1)
    b = e1.calc_batch(list(fr), step=1)
    assert np.array_equal(a, b)
    assert np.array_equal(b, e1.calc_batch(list(fr), step=1))  # and run to run (stage reuse, mbarrier phases)
