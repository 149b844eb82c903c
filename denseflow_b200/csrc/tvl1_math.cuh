// tvl1_math.cuh — per-pixel arithmetic of the TV-L1 inner loop, shared by the stand-alone
// kernels and the fused persistent kernel (upstream
// opencv_contrib/modules/cudaoptflow/src/cuda/tvl1flow.cu estimateUKernel /
// estimateDualVariablesKernel with gamma = 0).
#pragma once

#include <cfloat>

#include "tvl1.cuh"

namespace dfb {

// Primal half-step for one pixel. div1 = div(p11,p12), div2 = div(p21,p22) (backward differences,
// p outside the image = 0).
__device__ __forceinline__ void tvl1_primal_px(float ix, float iy, float g, float rc, float u1o, float u2o,
                                               float div1, float div2, const Tvl1Consts &c, float &u1n,
                                               float &u2n) {
    // The reference's three-way branch (rho < -l_t*g | rho > l_t*g | g > eps) as a branch-free factor f with
    // d = f * (I1wx, I1wy): f = +l_t, -l_t, -rho/g or 0.  Same products, no divergent code.
    const float rho = rc + (ix * u1o + iy * u2o);
    const float thr = c.l_t * g;
    float f = g > FLT_EPSILON ? f_div(-rho, g) : 0.f;
    f = rho > thr ? -c.l_t : f;
    f = rho < -thr ? c.l_t : f;
    const float d1 = f * ix, d2 = f * iy;
    u1n = (u1o + d1) + c.theta * div1;
    u2n = (u2o + d2) + c.theta * div2;
}

// Dual half-step for one pixel, from forward differences of the NEW u (index-clamped).
__device__ __forceinline__ void tvl1_dual_px(float u1x, float u1y, float u2x, float u2y, float taut, float &p11,
                                             float &p12, float &p21, float &p22) {
    const float g1 = f_hypot(u1x, u1y);
    const float g2 = f_hypot(u2x, u2y);
#ifdef DFB_STRICT_FP
    const float ng1 = 1.0f + taut * g1;
    const float ng2 = 1.0f + taut * g2;
    p11 = f_div(p11 + taut * u1x, ng1);
    p12 = f_div(p12 + taut * u1y, ng1);
    p21 = f_div(p21 + taut * u2x, ng2);
    p22 = f_div(p22 + taut * u2y, ng2);
#else
    const float r1 = f_rcp(fmaf(taut, g1, 1.0f));
    const float r2 = f_rcp(fmaf(taut, g2, 1.0f));
    p11 = fmaf(taut, u1x, p11) * r1;
    p12 = fmaf(taut, u1y, p12) * r1;
    p21 = fmaf(taut, u2x, p21) * r2;
    p22 = fmaf(taut, u2y, p22) * r2;
#endif
}

// Keys cubic (a = -0.5), SURVEY A.2 "Warp"
__device__ __forceinline__ float bicubic_coeff(float x) {
    x = fabsf(x);
    if (x <= 1.0f) return x * x * (1.5f * x - 2.5f) + 1.0f;
    if (x < 2.0f) return x * (x * (-0.5f * x + 2.5f) - 4.0f) + 2.0f;
    return 0.0f;
}

// A.2 "Warp (warpBackward)" for one pixel.  The reference sums taps cx = ceil(wx-2) .. floor(wx+2) (4, or 5 when wx
// is integral, in which case both end taps have weight k(+-2) = 0).  Anchored at xmin = ceil(wx-2) the distance to tap
// xmin+4 is in [2,3), so its weight is always exactly 0: a fixed 4x4 window with separable weights gives the same sums
// (tap order preserved: rows outer, columns inner).  I1 / I1x / I1y share one pitch; clamp addressing.
__device__ __forceinline__ void tvl1_warp_px(const float *__restrict__ I1, const float *I1x, const float *I1y, int W, int H, int P,
                                             int x, int y, float u1v, float u2v, float I0v, float &ix, float &iy, float &grad,
                                             float &rho_c) {
    const float wx = x + u1v, wy = y + u2v;
    const int xmin = (int)ceilf(wx - 2.0f), ymin = (int)ceilf(wy - 2.0f);
    float kx[4], ky[4];
    int cxs[4];
    size_t rows[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        kx[t] = bicubic_coeff(wx - (float)(xmin + t));
        ky[t] = bicubic_coeff(wy - (float)(ymin + t));
        cxs[t] = max(0, min(xmin + t, W - 1));
        rows[t] = (size_t)max(0, min(ymin + t, H - 1)) * P;
    }
    float sum = 0.f, sumx = 0.f, sumy = 0.f, wsum = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const float wgt = kx[b] * ky[a];
            const size_t t = rows[a] + cxs[b];
            sum = sum + wgt * __ldg(I1 + t);
            sumx = sumx + wgt * I1x[t];
            sumy = sumy + wgt * I1y[t];
            wsum = wsum + wgt;
        }
    }
    const float coeff = f_rcp(wsum);
    const float I1wv = sum * coeff;
    ix = sumx * coeff;
    iy = sumy * coeff;
    grad = ix * ix + iy * iy;
    rho_c = I1wv - ix * u1v - iy * u2v - I0v;
}

}  // namespace dfb
