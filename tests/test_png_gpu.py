"""GPU convertFlowToPngImage (/root/reference/src/common.cpp:18-46) bit-exact against the numpy restatement
(oracle/cv_restate.py, itself pinned to cv2 in tests/test_preproc_cpu.py)."""
import numpy as np
import pytest

from denseflow_b200 import synth
from oracle import cv_restate as R

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", [(64, 96, 0.7), (255, 341, 6.0), (120, 33, 20.0), (77, 201, 2.5), (1080, 1920, 3.0), (40, 50, 900.0)])
def test_png_image_is_bit_exact(case):
    import torch
    import denseflow_b200 as d
    h, w, scale = case
    rng = np.random.default_rng(h * 1000 + w)
    flow = (rng.standard_normal((h, w, 2)) * scale).astype(np.float32)
    if scale > 100:
        flow[3, 4] = (5000.0, -4000.0)  # larger than the frame: clipped to the extent
    e = d.create("tvl1", 0, 64, 64)
    for _ in range(2):  # twice: the reduction's ticket must re-arm itself
        bgr, bounds = e.flow_to_png_image_device(torch.from_numpy(flow).cuda())
        torch.cuda.synchronize()
        want, bx, by = R.flow_to_png_image(flow)
        assert bounds == (bx, by)
        assert np.array_equal(bgr.cpu().numpy(), want)


def test_png_image_of_an_engine_flow():
    import torch
    import denseflow_b200 as d
    a, b, _ = synth.pair(96, 128, 4)
    e = d.create("farn", 0, 128, 96)
    flow = e.calc(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda())
    bgr, bounds = e.flow_to_png_image_device(flow)
    torch.cuda.synchronize()
    want, bx, by = R.flow_to_png_image(flow.cpu().numpy())
    assert bounds == (bx, by) and np.array_equal(bgr.cpu().numpy(), want)
    import cv2
    ok, png = cv2.imencode(".png", bgr.cpu().numpy())  # encodeFlowMapPng's last step stays on the host (lossless)
    assert ok and np.array_equal(cv2.imdecode(png, cv2.IMREAD_COLOR), want)
