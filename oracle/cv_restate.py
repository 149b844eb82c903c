"""Numpy restatements of the two OpenCV CPU operations the reference's decode stage applies to every frame before
the hot path (/root/reference/src/denseflow_gpu.cpp:163-170): cvtColor(BGR2GRAY) and resize(INTER_LINEAR) on uint8.
TEST INFRASTRUCTURE (see oracle/oracle.h): pinned live against cv2 in tests/test_preproc_cpu.py; the CUDA kernels are
checked against these."""
import numpy as np


def bgr2gray(bgr):
    """OpenCV 4.x fixed-point BGR2GRAY: (B*3735 + G*19235 + R*9798 + 2^14) >> 15."""
    b, g, r = (bgr[..., i].astype(np.int64) for i in range(3))
    return ((b * 3735 + g * 19235 + r * 9798 + 16384) >> 15).astype(np.uint8)


def _coeffs(dn, sn, clamp):
    scale = 1.0 / (np.float64(dn) / sn)
    idx = np.zeros(dn, np.int64)
    a0 = np.zeros(dn, np.int64)
    a1 = np.zeros(dn, np.int64)
    for d in range(dn):
        f = np.float32((d + 0.5) * scale - 0.5)
        s = int(np.floor(f))
        f = np.float32(f - np.float32(s))
        if clamp:  # x direction: coefficients are reset at the borders (imgproc/resize.cpp)
            if s < 0:
                f, s = np.float32(0), 0
            if s >= sn - 1:
                f, s = np.float32(0), sn - 1
        idx[d] = s
        a0[d] = int(np.rint(np.float32(np.float32(1.0) - f) * np.float32(2048)))
        a1[d] = int(np.rint(f * np.float32(2048)))
    return idx, a0, a1


def resize_linear_u8(src, dw, dh):
    """cv::resize(src, (dw, dh), INTER_LINEAR) for CV_8UC1: 11-bit fixed-point coefficients, horizontal pass in int,
    vertical pass (((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2) >> 2; rows are index-clamped WITHOUT resetting beta."""
    sh, sw = src.shape
    if (dw, dh) == (sw, sh):
        return src.copy()
    xi, xa0, xa1 = _coeffs(dw, sw, True)
    yi, ya0, ya1 = _coeffs(dh, sh, False)
    s = src.astype(np.int64)
    rows = s[:, xi] * xa0 + s[:, np.minimum(xi + 1, sw - 1)] * xa1
    S0 = rows[np.clip(yi, 0, sh - 1)]
    S1 = rows[np.clip(yi + 1, 0, sh - 1)]
    out = ((((ya0[:, None] * (S0 >> 4)) >> 16) + ((ya1[:, None] * (S1 >> 4)) >> 16) + 2) >> 2)
    return np.clip(out, 0, 255).astype(np.uint8)


def new_size(w, h, new_width=0, new_height=0, new_short=0):
    """DenseFlow::get_new_size (/root/reference/src/denseflow_gpu.cpp:44-80): returns (do_resize, w, h)."""
    import math
    cround = lambda v: int(math.floor(v + 0.5))  # C round() of a positive value (half away from zero)
    if new_width > 0 and new_height > 0:
        return True, new_width, new_height
    if new_width > 0:
        return True, new_width, cround(h * 1.0 / w * new_width)
    if new_height > 0:
        return True, cround(w * 1.0 / h * new_height), new_height
    if new_short > 0 and min(w, h) > new_short:
        if w < h:
            return True, new_short, cround(h * 1.0 / w * new_short)
        return True, cround(w * 1.0 / h * new_short), new_short
    return False, w, h
