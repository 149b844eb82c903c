"""nvJPEG decode of `-if` frame folders (SURVEY §8 f3 remainder) against OpenCV's imread path (libjpeg-turbo):
/root/reference/src/denseflow_gpu.cpp:154-163.  Not bit-identical by construction (different IDCT / chroma upsampling);
the test bounds the gray-level difference and the flow it causes."""
import numpy as np
import pytest

from denseflow_b200 import synth

pytestmark = pytest.mark.gpu


def test_jpeg_decode_close_to_imread_and_flow_impact_is_small():
    import cv2
    import torch
    import denseflow_b200 as d
    gray = synth.stream(256, 340, 2, seed=61)
    e = d.OpticalFlowDual_TVL1.create(0, 340, 256)
    ours, ref = [], []
    for g in gray:
        bgr = np.stack([np.roll(g, 2, 1), g, 255 - np.roll(g, 3, 0)], -1).astype(np.uint8)  # a colour frame: chroma planes matter
        ok, jpg = cv2.imencode(".jpg", bgr)                                                  # what a frame folder holds (4:2:0, q95)
        assert ok
        want = cv2.cvtColor(cv2.imdecode(jpg, cv2.IMREAD_COLOR), cv2.COLOR_BGR2GRAY)
        got = e.decode_jpeg_gray_device(jpg.tobytes(), 340, 256)
        torch.cuda.synchronize()
        got = got.cpu().numpy()
        assert got.shape == want.shape
        diff = np.abs(got.astype(int) - want.astype(int))
        print("decode: max level diff", diff.max(), "mean", diff.mean(), "differing pixels", (diff > 0).mean())
        assert diff.max() <= 4 and diff.mean() < 0.5
        ours.append(got)
        ref.append(want)
    aee = synth.aee(e.calc(ours[0], ours[1]), e.calc(ref[0], ref[1]))
    print("flow AEE nvJPEG-decoded vs imread-decoded frames: %.4f px" % aee)
    assert aee < 0.05
    # a gray (single component) JPEG decodes too
    ok, jg = cv2.imencode(".jpg", gray[0])
    got = e.decode_jpeg_gray_device(jg.tobytes(), 340, 256).cpu().numpy()
    want = cv2.cvtColor(cv2.imdecode(jg, cv2.IMREAD_COLOR), cv2.COLOR_BGR2GRAY)
    assert np.abs(got.astype(int) - want.astype(int)).max() <= 2
    with pytest.raises(RuntimeError):
        e.decode_jpeg_gray_device(jg.tobytes(), 64, 64)  # larger than the output buffer
