"""nvJPEG encode of quantised flow planes vs OpenCV's imencode defaults (src/common.cpp:56-57)."""
import numpy as np
import pytest

from denseflow_b200 import synth

pytestmark = pytest.mark.gpu


def test_jpeg_encode_decodes_like_opencv(oracle):
    import cv2
    import torch
    import denseflow_b200 as d
    a, b, _ = synth.pair(256, 340, 5)
    e = d.OpticalFlowDual_TVL1.create(0, 340, 256)
    flow = e.calc(a, b)
    qx, qy = oracle.quantise(flow, 20)
    for plane in (qx, qy):
        jpg = e.encode_jpeg_gray_device(torch.from_numpy(plane).cuda(), 95)
        assert jpg[:2] == b"\xff\xd8" and jpg[-2:] == b"\xff\xd9"
        dec = cv2.imdecode(np.frombuffer(jpg, np.uint8), cv2.IMREAD_UNCHANGED)
        assert dec.shape == plane.shape and dec.dtype == np.uint8  # single gray component
        ok, ref = cv2.imencode(".jpg", plane)  # OpenCV defaults: quality 95
        ref_dec = cv2.imdecode(ref, cv2.IMREAD_UNCHANGED)
        err = np.abs(dec.astype(int) - plane.astype(int))
        ref_err = np.abs(ref_dec.astype(int) - plane.astype(int))
        print("nvjpeg bytes", len(jpg), "opencv bytes", len(ref), "max err", err.max(), ref_err.max(), "mean", err.mean(), ref_err.mean())
        assert err.max() <= max(ref_err.max() + 2, 4) and err.mean() <= ref_err.mean() + 0.25
        assert 0.5 < len(jpg) / len(ref) < 2.0
