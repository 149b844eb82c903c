"""BGR frames -> gray -> resize -> flow -> bound+quantise -> JPEG on the GPU (dfb_process_bgr_batch_host) vs the
reference's CPU stages done with OpenCV around the same engine (src/denseflow_gpu.cpp:163-170, src/common.cpp:48-64)."""
import numpy as np
import pytest

from denseflow_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("alg", ["tvl1", "farn"])
@pytest.mark.parametrize("resize", [None, (170, 128)])
def test_bgr_pipeline_matches_cpu_stages(alg, resize):
    import cv2
    import denseflow_b200 as d
    gray = synth.stream(192, 256, 5, seed=41)
    rng = np.random.default_rng(0)
    # colour frames whose gray conversion is non-trivial
    bgr = [np.stack([np.roll(g, 3, 1), g, 255 - np.roll(g, 5, 0)], -1).astype(np.uint8) ^ rng.integers(0, 4, g.shape + (3,), dtype=np.uint8)
           for g in gray]
    w, h = resize if resize else (256, 192)
    e = d.create(alg, 0, 256, 192)
    out = e.process_bgr_batch(bgr, step=1, bound=20, new_size=resize)
    assert len(out) == 4
    # CPU stages with real OpenCV, flow + quantise with the same engine
    frames = [cv2.cvtColor(f, cv2.COLOR_BGR2GRAY) for f in bgr]
    if resize:
        frames = [cv2.resize(f, resize) for f in frames]
    qx, qy = e.calc_batch(frames, step=1, bound=20)
    for i, (jx, jy) in enumerate(out):
        dx = cv2.imdecode(np.frombuffer(jx, np.uint8), cv2.IMREAD_UNCHANGED)
        dy = cv2.imdecode(np.frombuffer(jy, np.uint8), cv2.IMREAD_UNCHANGED)
        assert dx.shape == (h, w) and dy.shape == (h, w)
        # identical quantised planes go into the encoder (gray/resize/flow/quantise are bit-exact); only JPEG loss remains
        assert np.abs(dx.astype(int) - qx[i].astype(int)).max() <= 8 and np.abs(dx.astype(int) - qx[i].astype(int)).mean() < 0.6
        assert np.abs(dy.astype(int) - qy[i].astype(int)).max() <= 8 and np.abs(dy.astype(int) - qy[i].astype(int)).mean() < 0.6
