"""The gray / resize restatements (oracle/cv_restate.py) against real OpenCV CPU code (cv2 is in the image)."""
import numpy as np
import pytest

from oracle import cv_restate as R

cv2 = pytest.importorskip("cv2")


def test_bgr2gray_matches_opencv():
    rng = np.random.default_rng(0)
    bgr = rng.integers(0, 256, (97, 131, 3), dtype=np.uint8)
    assert np.array_equal(R.bgr2gray(bgr), cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY))
    ramp = np.stack(np.meshgrid(np.arange(256), np.arange(256)), -1).astype(np.uint8)
    full = np.concatenate([ramp, np.full((256, 256, 1), 77, np.uint8)], -1)
    assert np.array_equal(R.bgr2gray(full), cv2.cvtColor(full, cv2.COLOR_BGR2GRAY))


@pytest.mark.parametrize("dst", [(160, 120), (341, 256), (224, 224), (400, 300), (455, 256), (256, 341), (100, 77), (639, 479),
                                 (340, 256), (33, 500), (320, 240)])
def test_resize_linear_u8_matches_opencv(dst):
    rng = np.random.default_rng(1)
    src = rng.integers(0, 256, (240, 320), dtype=np.uint8)
    dw, dh = dst
    assert np.array_equal(R.resize_linear_u8(src, dw, dh), cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR))


def test_new_size_rules():
    # DenseFlow::get_new_size, /root/reference/src/denseflow_gpu.cpp:57-78
    assert R.new_size(1920, 1080, new_short=256) == (True, 455, 256)
    assert R.new_size(320, 240, new_short=256) == (False, 320, 240)  # ns only shrinks
    assert R.new_size(1080, 1920, new_short=256) == (True, 256, 455)
    assert R.new_size(640, 480, new_width=340) == (True, 340, 255)
    assert R.new_size(640, 480, new_height=256) == (True, 341, 256)
    assert R.new_size(640, 480, 340, 256) == (True, 340, 256)
