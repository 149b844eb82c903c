"""Second, independent restatement of SURVEY.md Appendix A (TV-L1 as cv::cuda::OpticalFlowDual_TVL1 executes it) in
vectorised numpy fp32 — TEST INFRASTRUCTURE (see oracle/oracle.h).  It shares no code with oracle/tvl1_oracle.c; the
CPU suite requires the two to agree, which guards the C oracle against indexing / control-flow slips.  (It cannot pin
the restatement to OpenCV: no TV-L1 implementation exists in this image — parity stays unpinned.)"""
import numpy as np

F = np.float32


def resize_cuda(src, dw, dh, fx, fy):
    """A.1: src = dst * f, floor, x2/y2 reads clamped, four-term accumulation in the upstream order."""
    sh, sw = src.shape
    dx = np.arange(dw, dtype=F)
    dy = np.arange(dh, dtype=F)
    sx = dx * F(fx)
    sy = dy * F(fy)
    x1 = np.floor(sx).astype(np.int64)
    y1 = np.floor(sy).astype(np.int64)
    x2, y2 = x1 + 1, y1 + 1
    x1r, y1r = np.minimum(x1, sw - 1), np.minimum(y1, sh - 1)
    x2r, y2r = np.minimum(x2, sw - 1), np.minimum(y2, sh - 1)
    wx2 = (x2.astype(F) - sx)[None, :]
    wx1 = (sx - x1.astype(F))[None, :]
    wy2 = (y2.astype(F) - sy)[:, None]
    wy1 = (sy - y1.astype(F))[:, None]
    out = np.zeros((dh, dw), F)
    out = out + src[np.ix_(y1r, x1r)] * (wx2 * wy2)
    out = out + src[np.ix_(y1r, x2r)] * (wx1 * wy2)
    out = out + src[np.ix_(y2r, x1r)] * (wx2 * wy1)
    out = out + src[np.ix_(y2r, x2r)] * (wx1 * wy1)
    return out.astype(F)


def gradient(I):
    """A.2 step 1: half central differences with index clamp."""
    Ixp = np.concatenate([I[:, 1:], I[:, -1:]], 1)
    Ixm = np.concatenate([I[:, :1], I[:, :-1]], 1)
    Iyp = np.concatenate([I[1:], I[-1:]], 0)
    Iym = np.concatenate([I[:1], I[:-1]], 0)
    return (F(0.5) * (Ixp - Ixm)).astype(F), (F(0.5) * (Iyp - Iym)).astype(F)


def keys(t):
    t = np.abs(t).astype(F)
    a = t * t * (F(1.5) * t - F(2.5)) + F(1.0)
    b = t * (t * (F(-0.5) * t + F(2.5)) - F(4.0)) + F(2.0)
    return np.where(t <= 1, a, np.where(t < 2, b, F(0))).astype(F)


def warp(I0, I1, I1x, I1y, u1, u2):
    """A.2 warp: taps cx = ceil(wx-2) .. floor(wx+2) (4 or 5 per axis), weight-normalised, clamp addressing."""
    h, w = I0.shape
    yy, xx = np.mgrid[0:h, 0:w]
    wx = (xx.astype(F) + u1).astype(F)
    wy = (yy.astype(F) + u2).astype(F)
    xmin = np.ceil(wx - F(2)).astype(np.int64)
    xmax = np.floor(wx + F(2)).astype(np.int64)
    ymin = np.ceil(wy - F(2)).astype(np.int64)
    ymax = np.floor(wy + F(2)).astype(np.int64)
    s = np.zeros((h, w), F); sx = np.zeros((h, w), F); sy = np.zeros((h, w), F); ws = np.zeros((h, w), F)
    for a in range(5):          # rows outer, columns inner: the upstream loop order
        cy = ymin + a
        for b in range(5):
            cx = xmin + b
            wgt = np.where((cy <= ymax) & (cx <= xmax), keys(wx - cx.astype(F)) * keys(wy - cy.astype(F)), F(0)).astype(F)
            yc = np.clip(cy, 0, h - 1)
            xc = np.clip(cx, 0, w - 1)
            s = (s + wgt * I1[yc, xc]).astype(F)
            sx = (sx + wgt * I1x[yc, xc]).astype(F)
            sy = (sy + wgt * I1y[yc, xc]).astype(F)
            ws = (ws + wgt).astype(F)
    c = (F(1) / ws).astype(F)
    I1w, ix, iy = (s * c).astype(F), (sx * c).astype(F), (sy * c).astype(F)
    grad = (ix * ix + iy * iy).astype(F)
    rho_c = (((I1w - ix * u1).astype(F) - iy * u2).astype(F) - I0).astype(F)
    return ix, iy, grad, rho_c


def divergence(pa, pb):
    """A.3 div: backward differences, p outside the image = 0."""
    h, w = pa.shape
    left = np.concatenate([np.zeros((h, 1), F), pa[:, :-1]], 1)
    up = np.concatenate([np.zeros((1, w), F), pb[:-1]], 0)
    interior = ((pa - left).astype(F) + (pb - up).astype(F)).astype(F)
    # the reference's border forms: (v1 + v2) - v2up on x == 0, (v1 - v1left) + v2 on y == 0, v1 + v2 in the corner
    out = interior.copy()
    out[1:, 0] = ((pa[1:, 0] + pb[1:, 0]).astype(F) - pb[:-1, 0]).astype(F)
    out[0, 1:] = ((pa[0, 1:] - pa[0, :-1]).astype(F) + pb[0, 1:]).astype(F)
    out[0, 0] = pa[0, 0] + pb[0, 0]
    return out


def estimate_u(ix, iy, grad, rho_c, p11, p12, p21, p22, u1, u2, l_t, theta):
    rho = (rho_c + (ix * u1 + iy * u2).astype(F)).astype(F)
    thr = (F(l_t) * grad).astype(F)
    with np.errstate(divide="ignore", invalid="ignore"):
        fi = (-rho / grad).astype(F)
    d1 = np.zeros_like(u1); d2 = np.zeros_like(u2)
    m3 = grad > np.finfo(F).eps
    d1 = np.where(m3, fi * ix, d1); d2 = np.where(m3, fi * iy, d2)
    m2 = rho > thr
    d1 = np.where(m2, -F(l_t) * ix, d1); d2 = np.where(m2, -F(l_t) * iy, d2)
    m1 = rho < -thr
    d1 = np.where(m1, F(l_t) * ix, d1); d2 = np.where(m1, F(l_t) * iy, d2)
    n1 = ((u1 + d1.astype(F)).astype(F) + (F(theta) * divergence(p11, p12)).astype(F)).astype(F)
    n2 = ((u2 + d2.astype(F)).astype(F) + (F(theta) * divergence(p21, p22)).astype(F)).astype(F)
    diff = (((u1 - n1) * (u1 - n1)).astype(F) + ((u2 - n2) * (u2 - n2)).astype(F)).astype(F)
    return n1, n2, float(diff.astype(np.float64).sum())


def estimate_dual(u1, u2, p11, p12, p21, p22, taut):
    def fwd(u):
        ux = (np.concatenate([u[:, 1:], u[:, -1:]], 1) - u).astype(F)
        uy = (np.concatenate([u[1:], u[-1:]], 0) - u).astype(F)
        return ux, uy
    u1x, u1y = fwd(u1)
    u2x, u2y = fwd(u2)
    g1 = np.hypot(u1x.astype(np.float64), u1y.astype(np.float64)).astype(F)
    g2 = np.hypot(u2x.astype(np.float64), u2y.astype(np.float64)).astype(F)
    ng1 = (F(1) + F(taut) * g1).astype(F)
    ng2 = (F(1) + F(taut) * g2).astype(F)
    return (((p11 + F(taut) * u1x).astype(F) / ng1).astype(F), ((p12 + F(taut) * u1y).astype(F) / ng1).astype(F),
            ((p21 + F(taut) * u2x).astype(F) / ng2).astype(F), ((p22 + F(taut) * u2y).astype(F) / ng2).astype(F))


def calc(I0u8, I1u8, tau=0.25, lam=0.15, theta=0.3, nscales=5, warps=5, epsilon=0.01, iterations=300, scale_step=0.8):
    """Returns (flow [H,W,2] float32, iteration log [levels, warps])."""
    I0s = [I0u8.astype(F)]
    I1s = [I1u8.astype(F)]
    finv = F(1.0 / scale_step)
    for s in range(1, nscales):
        ph, pw = I0s[-1].shape
        nw, nh = int(np.rint(pw * scale_step)), int(np.rint(ph * scale_step))
        if nw < 16 or nh < 16:
            break
        I0s.append(resize_cuda(I0s[-1], nw, nh, finv, finv))
        I1s.append(resize_cuda(I1s[-1], nw, nh, finv, finv))
    n = len(I0s)
    l_t, taut = F(lam * theta), F(tau / theta)
    u1 = np.zeros_like(I0s[-1]); u2 = np.zeros_like(I0s[-1])
    log = np.zeros((nscales, warps), np.int32)
    for s in range(n - 1, -1, -1):
        I0, I1 = I0s[s], I1s[s]
        h, w = I0.shape
        I1x, I1y = gradient(I1)
        p11 = np.zeros_like(I0); p12 = np.zeros_like(I0); p21 = np.zeros_like(I0); p22 = np.zeros_like(I0)
        scaled_eps = epsilon * epsilon * (w * h)
        for wi in range(warps):
            ix, iy, grad, rho_c = warp(I0, I1, I1x, I1y, u1, u2)
            error, prev, it = np.finfo(np.float64).max, 0.0, 0
            while error > scaled_eps and it < iterations:
                calc_err = epsilon > 0 and (it & 1) and prev < scaled_eps
                u1, u2, e = estimate_u(ix, iy, grad, rho_c, p11, p12, p21, p22, u1, u2, l_t, F(theta))
                if calc_err:
                    error = prev = e
                else:
                    error = np.finfo(np.float64).max
                    prev -= scaled_eps
                p11, p12, p21, p22 = estimate_dual(u1, u2, p11, p12, p21, p22, taut)
                it += 1
            log[s, wi] = it
        if s > 0:
            th, tw = I0s[s - 1].shape
            ufx, ufy = F(1.0 / (tw / w)), F(1.0 / (th / h))
            u1 = (resize_cuda(u1, tw, th, ufx, ufy) * F(1.0 / scale_step)).astype(F)
            u2 = (resize_cuda(u2, tw, th, ufx, ufy) * F(1.0 / scale_step)).astype(F)
    return np.stack([u1, u2], -1), log
