"""Deterministic synthetic frame generator (SURVEY.md §8(d) recipe; seeds fixed).

Used by tests and bench.py to make gray frame pairs / streams with a known smooth motion field.
Needs numpy + cv2 (GaussianBlur / remap) — host-side data prep only, not on the flow path.
"""
import hashlib

import numpy as np


def _texture(seed, H, W):
    import cv2
    r = np.random.default_rng(seed).random((H + 64, W + 64)).astype(np.float32)
    A = cv2.GaussianBlur(r, (0, 0), 3.0)
    B = cv2.GaussianBlur(r, (0, 0), 1.0)
    return 0.6 * (A - A.min()) / (A.max() - A.min()) + 0.4 * (B - B.min()) / (B.max() - B.min())


def _quant(f):
    return np.clip(np.rint(255.0 * f), 0, 255).astype(np.uint8)


def _shifted(T, H, W, X, Y):
    import cv2
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    return cv2.remap(T, (xx + 32 - X).astype(np.float32), (yy + 32 - Y).astype(np.float32), cv2.INTER_CUBIC)


def pair(H=256, W=256, seed=0):
    """Config-1/2 pair.  Returns (frame0 u8, frame1 u8, gt_flow float32 [H,W,2])."""
    T = _texture(seed, H, W)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    X = 2.0 + 1.5 * np.sin(2 * np.pi * yy / H)
    Y = -1.0 + 1.0 * np.cos(2 * np.pi * xx / W)
    f0 = T[32:32 + H, 32:32 + W]
    f1 = _shifted(T, H, W, X.astype(np.float32), Y.astype(np.float32))
    return _quant(f0), _quant(f1), np.stack([X, Y], -1).astype(np.float32)


def stream(H, W, n_frames, seed, phase=0.0):
    """Config-3/4/5 stream: frame t = texture displaced by (X_t, Y_t).  Returns uint8 [n,H,W]."""
    T = _texture(seed, H, W)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    out = np.empty((n_frames, H, W), np.uint8)
    for t in range(n_frames):
        tt = t + phase
        X = 12 * np.sin(2 * np.pi * tt / 50) + 1.5 * np.sin(2 * np.pi * yy / H + tt / 20)
        Y = 8 * np.cos(2 * np.pi * tt / 70) + 1.0 * np.cos(2 * np.pi * xx / W + tt / 25)
        out[t] = _quant(_shifted(T, H, W, X.astype(np.float32), Y.astype(np.float32)))
    return out


def noise_pair(H=256, W=256, seed=7):
    """Adversarial pair for the 300-iteration cap: frame b is an independent texture."""
    a = _quant(_texture(seed, H, W)[32:32 + H, 32:32 + W])
    b = _quant(_texture(seed + 1000, H, W)[32:32 + H, 32:32 + W])
    return a, b


def sha1(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


def aee(f, g):
    d = np.asarray(f, np.float64) - np.asarray(g, np.float64)
    return float(np.mean(np.hypot(d[..., 0], d[..., 1])))
