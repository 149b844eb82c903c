"""GPU BGR->gray and INTER_LINEAR uint8 resize — bit-exact vs the OpenCV restatements."""
import numpy as np
import pytest

from oracle import cv_restate as R

pytestmark = pytest.mark.gpu


def _engine():
    import denseflow_b200 as d
    return d.OpticalFlowDual_TVL1.create(0, 64, 64)


def test_bgr_to_gray_bit_exact():
    import torch
    rng = np.random.default_rng(0)
    e = _engine()
    for shape in [(97, 131), (1080, 1920), (1, 1)]:
        bgr = rng.integers(0, 256, shape + (3,), dtype=np.uint8)
        g = e.bgr_to_gray_device(torch.from_numpy(bgr).cuda()).cpu().numpy()
        assert np.array_equal(g, R.bgr2gray(bgr))


@pytest.mark.parametrize("case", [((240, 320), (160, 120)), ((240, 320), (341, 256)), ((240, 320), (400, 300)), ((1080, 1920), (455, 256)),
                                  ((480, 640), (340, 256)), ((77, 100), (33, 500)), ((240, 320), (320, 240))])
def test_resize_bit_exact(case):
    import torch
    (sh, sw), (dw, dh) = case
    rng = np.random.default_rng(2)
    src = rng.integers(0, 256, (sh, sw), dtype=np.uint8)
    e = _engine()
    out = e.resize_gray_device(torch.from_numpy(src).cuda(), dw, dh).cpu().numpy()
    assert np.array_equal(out, R.resize_linear_u8(src, dw, dh))
