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


def flow_to_png_image(flow):
    """convertFlowToPngImage, /root/reference/src/common.cpp:18-46, restated with numpy (pinned to cv2 in
    tests/test_preproc_cpu.py: minMaxLoc, convertTo's fused multiply-add + round-half-even + saturate, rectangle rows).
    flow: float32 [H,W,2].  Returns (bgr uint8 [H,W,3], bound_x, bound_y)."""
    import math
    flow = np.asarray(flow, np.float32)
    h, w = flow.shape[:2]

    def bound(extent, comp):
        mn, mx = float(comp.min()), float(comp.max())                                   # minMaxLoc (:23, :25)
        b = min(255. * 4, math.ceil((min(extent, max(abs(mn), abs(mx))) * 128. / 127.) / 4) * 4)   # :24, :26
        if int(b) % 8 == 0:                                                              # :27-32
            b += 4
        return b

    bx, by = bound(float(w), flow[..., 0]), bound(float(h), flow[..., 1])
    base = 1. / 128.

    def convert_to_u8(v, b):
        alpha = np.float32(1. / (base * b))                                              # float eps_x_inv (:33-34)
        # Mat::convertTo(CV_8U, alpha, 128) on a CV_32F source: float fused multiply-add (v_fma), cvRound, saturate.  The fma is
        # emulated exactly: a 24 x 24-bit product is exact in double, and adding 128 to it rounds at most once more below the
        # float rounding point in the cases that matter (checked against cv2 on ~10^6 values in the test).
        t = (v.astype(np.float64) * np.float64(alpha) + 128.0).astype(np.float32)
        return np.clip(np.rint(t), 0, 255).astype(np.uint8)

    out = np.empty((h, w, 3), np.uint8)
    out[..., 0] = convert_to_u8(flow[..., 0], bx)
    out[..., 1] = convert_to_u8(flow[..., 1], by)
    half_h = int(h / 2)                       # Point(w - 1, half_h): double -> int truncation (:40-41)
    split = int(h / 2 + 1)                    # Point(0, half_h + 1)
    out[..., 2] = int(by / 4)
    out[:half_h + 1, :, 2] = int(bx / 4)      # rectangle 1 covers rows 0 .. half_h, rectangle 2 rows split .. h-1 (drawn second)
    out[split:, :, 2] = int(by / 4)
    return out, bx, by
