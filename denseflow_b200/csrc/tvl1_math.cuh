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
    const float rho = rc + (ix * u1o + iy * u2o);
    const float thr = c.l_t * g;
    float d1 = 0.f, d2 = 0.f;
    if (rho < -thr) {
        d1 = c.l_t * ix;
        d2 = c.l_t * iy;
    } else if (rho > thr) {
        d1 = -c.l_t * ix;
        d2 = -c.l_t * iy;
    } else if (g > FLT_EPSILON) {
        const float fi = f_div(-rho, g);
        d1 = fi * ix;
        d2 = fi * iy;
    }
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

}  // namespace dfb
