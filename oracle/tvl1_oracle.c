/*
 * tvl1_oracle.c — CPU restatement (fp32 IEEE, no fast-math, no FMA contraction) of what
 * cv::cuda::OpticalFlowDual_TVL1::calc executes with the argument-less create() defaults the
 * reference uses (/root/reference/src/denseflow_gpu.cpp:299,327).
 *
 * TEST INFRASTRUCTURE — see oracle.h.  PARITY UNPINNED for TV-L1 (no external TV-L1 runs here).
 *
 * The arithmetic lives in OpenCV/opencv_contrib 4.5.2 (docker/Dockerfile:6), not in /root/reference:
 *   modules/cudaoptflow/src/tvl1flow.cpp        host control flow    -> SURVEY.md Appendix A.1, A.2, A.4
 *   modules/cudaoptflow/src/cuda/tvl1flow.cu    the four kernels     -> SURVEY.md Appendix A.2, A.3
 *   modules/cudawarping/src/cuda/resize.cu      resize_linear        -> SURVEY.md Appendix A.1
 * Each function below names the appendix paragraph it follows.
 */
#include "oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* create() defaults — SURVEY §2.2 / Appendix A header */
void orc_tvl1_default_params(orc_tvl1_params *p) {
    p->tau = 0.25;
    p->lambda = 0.15;
    p->theta = 0.3;
    p->nscales = 5;
    p->warps = 5;
    p->epsilon = 0.01;
    p->iterations = 300;
    p->scale_step = 0.8;
}

static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

/* A.1: I0s[0] = float(I0), scale 1.0 (intensities stay 0..255) */
void orc_u8_to_f32(const uint8_t *src, int w, int h, float *dst) {
    const long n = (long)w * h;
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i) dst[i] = (float)src[i];
}

/* A.1: CUDA bilinear resize (cudawarping resize_linear): src = dst * f, floor, x2/y2 reads clamped.
 * convention 1 restates the OpenCV-CPU INTER_LINEAR sampling (half-pixel centres, edge clamp) and is
 * used only to pin the Farneback restatement against cv2 (oracle.h). */
void orc_resize_linear(const float *src, int sw, int sh, float *dst, int dw, int dh, float fx, float fy,
                       int convention) {
#pragma omp parallel for schedule(static)
    for (int dy = 0; dy < dh; ++dy) {
        for (int dx = 0; dx < dw; ++dx) {
            if (convention == ORC_RESIZE_CUDA) {
                const float sx = (float)dx * fx;
                const float sy = (float)dy * fy;
                const int x1 = (int)floorf(sx);
                const int y1 = (int)floorf(sy);
                const int x2 = x1 + 1;
                const int y2 = y1 + 1;
                const int x2r = imin(x2, sw - 1);
                const int y2r = imin(y2, sh - 1);
                const int x1r = imin(x1, sw - 1); /* never binds for the sizes the algorithm produces */
                const int y1r = imin(y1, sh - 1);
                float out = 0.f;
                out = out + src[(long)y1r * sw + x1r] * (((float)x2 - sx) * ((float)y2 - sy));
                out = out + src[(long)y1r * sw + x2r] * ((sx - (float)x1) * ((float)y2 - sy));
                out = out + src[(long)y2r * sw + x1r] * (((float)x2 - sx) * (sy - (float)y1));
                out = out + src[(long)y2r * sw + x2r] * ((sx - (float)x1) * (sy - (float)y1));
                dst[(long)dy * dw + dx] = out;
            } else {
                /* cv::resize INTER_LINEAR (imgproc/resize.cpp): scale from the sizes, in double */
                const double scx = 1.0 / ((double)dw / (double)sw);
                const double scy = 1.0 / ((double)dh / (double)sh);
                float sx = (float)(((double)dx + 0.5) * scx - 0.5);
                float sy = (float)(((double)dy + 0.5) * scy - 0.5);
                int x1 = (int)floorf(sx);
                int y1 = (int)floorf(sy);
                float ax = sx - (float)x1;
                float ay = sy - (float)y1;
                if (x1 < 0) { x1 = 0; ax = 0.f; }
                if (x1 >= sw - 1) { x1 = sw - 1; ax = 0.f; }
                if (y1 < 0) { y1 = 0; ay = 0.f; }
                if (y1 >= sh - 1) { y1 = sh - 1; ay = 0.f; }
                const int x2 = imin(x1 + 1, sw - 1);
                const int y2 = imin(y1 + 1, sh - 1);
                const float top = src[(long)y1 * sw + x1] * (1.f - ax) + src[(long)y1 * sw + x2] * ax;
                const float bot = src[(long)y2 * sw + x1] * (1.f - ax) + src[(long)y2 * sw + x2] * ax;
                dst[(long)dy * dw + dx] = top * (1.f - ay) + bot * ay;
            }
        }
    }
}

/* A.2 step 1: centeredGradientKernel — half central difference, index-clamped borders */
void orc_centered_gradient(const float *src, int w, int h, float *dx, float *dy) {
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y) {
        for (int x = 0; x < w; ++x) {
            dx[(long)y * w + x] = 0.5f * (src[(long)y * w + imin(x + 1, w - 1)] - src[(long)y * w + imax(x - 1, 0)]);
            dy[(long)y * w + x] = 0.5f * (src[(long)imin(y + 1, h - 1) * w + x] - src[(long)imax(y - 1, 0) * w + x]);
        }
    }
}

/* A.2 "Warp": Keys cubic, a = -0.5 */
static inline float bicubic_coeff(float x) {
    x = fabsf(x);
    if (x <= 1.0f) return x * x * (1.5f * x - 2.5f) + 1.0f;
    if (x < 2.0f) return x * (x * (-0.5f * x + 2.5f) - 4.0f) + 2.0f;
    return 0.0f;
}

static inline float fetch_clamp(const float *img, int w, int h, int x, int y) {
    x = imax(0, imin(x, w - 1));
    y = imax(0, imin(y, h - 1));
    return img[(long)y * w + x];
}

/* A.2 "Warp (warpBackward)": weight-normalised 4x4 (up to 5x5 when on-grid) bicubic gather of
 * I1, I1x, I1y with clamp addressing; grad and rho_c. */
void orc_tvl1_warp_backward(const float *I0, const float *I1, const float *I1x, const float *I1y, const float *u1,
                            const float *u2, int w, int h, float *I1w, float *I1wx, float *I1wy, float *grad,
                            float *rho_c) {
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y) {
        for (int x = 0; x < w; ++x) {
            const long i = (long)y * w + x;
            const float u1v = u1[i], u2v = u2[i];
            const float wx = (float)x + u1v;
            const float wy = (float)y + u2v;
            const int xmin = (int)ceilf(wx - 2.0f);
            const int xmax = (int)floorf(wx + 2.0f);
            const int ymin = (int)ceilf(wy - 2.0f);
            const int ymax = (int)floorf(wy + 2.0f);
            float sum = 0.f, sumx = 0.f, sumy = 0.f, wsum = 0.f;
            for (int cy = ymin; cy <= ymax; ++cy) {
                for (int cx = xmin; cx <= xmax; ++cx) {
                    const float wgt = bicubic_coeff(wx - (float)cx) * bicubic_coeff(wy - (float)cy);
                    sum = sum + wgt * fetch_clamp(I1, w, h, cx, cy);
                    sumx = sumx + wgt * fetch_clamp(I1x, w, h, cx, cy);
                    sumy = sumy + wgt * fetch_clamp(I1y, w, h, cx, cy);
                    wsum = wsum + wgt;
                }
            }
            const float coeff = 1.0f / wsum;
            const float I1wv = sum * coeff;
            const float I1wxv = sumx * coeff;
            const float I1wyv = sumy * coeff;
            I1w[i] = I1wv;
            I1wx[i] = I1wxv;
            I1wy[i] = I1wyv;
            grad[i] = I1wxv * I1wxv + I1wyv * I1wyv;
            rho_c[i] = I1wv - I1wxv * u1v - I1wyv * u2v - I0[i];
        }
    }
}

/* A.3 div(pa,pb): backward differences, out-of-range p = 0 */
static inline float divergence(const float *v1, const float *v2, int w, int y, int x) {
    const long i = (long)y * w + x;
    if (x > 0 && y > 0) {
        const float v1x = v1[i] - v1[i - 1];
        const float v2y = v2[i] - v2[i - w];
        return v1x + v2y;
    }
    if (y > 0) return v1[i] + v2[i] - v2[i - w];
    if (x > 0) return v1[i] - v1[i - 1] + v2[i];
    return v1[i] + v2[i];
}

/* A.3 primal (estimateUKernel, gamma = 0).  In-place on u1,u2 (pointwise in u, reads p only). */
double orc_tvl1_estimate_u(const float *I1wx, const float *I1wy, const float *grad, const float *rho_c,
                           const float *p11, const float *p12, const float *p21, const float *p22, float *u1,
                           float *u2, int w, int h, float l_t, float theta, int calc_error) {
    double err = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : err)
    for (int y = 0; y < h; ++y) {
        double row_err = 0.0;
        for (int x = 0; x < w; ++x) {
            const long i = (long)y * w + x;
            const float ix = I1wx[i], iy = I1wy[i], g = grad[i];
            const float u1o = u1[i], u2o = u2[i];
            const float rho = rho_c[i] + (ix * u1o + iy * u2o);
            float d1 = 0.f, d2 = 0.f;
            if (rho < -l_t * g) {
                d1 = l_t * ix;
                d2 = l_t * iy;
            } else if (rho > l_t * g) {
                d1 = -l_t * ix;
                d2 = -l_t * iy;
            } else if (g > FLT_EPSILON) {
                const float fi = -rho / g;
                d1 = fi * ix;
                d2 = fi * iy;
            }
            const float v1 = u1o + d1;
            const float v2 = u2o + d2;
            const float div1 = divergence(p11, p12, w, y, x);
            const float div2 = divergence(p21, p22, w, y, x);
            const float u1n = v1 + theta * div1;
            const float u2n = v2 + theta * div2;
            u1[i] = u1n;
            u2[i] = u2n;
            if (calc_error) {
                const float n1 = (u1o - u1n) * (u1o - u1n);
                const float n2 = (u2o - u2n) * (u2o - u2n);
                row_err += (double)(n1 + n2);
            }
        }
        err += row_err;
    }
    return err;
}

/* A.3 dual (estimateDualVariablesKernel): forward differences of the NEW u, index-clamped.
 * In-place on p (pointwise in p, reads u only). */
void orc_tvl1_estimate_dual(const float *u1, const float *u2, float *p11, float *p12, float *p21, float *p22, int w,
                            int h, float taut) {
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y) {
        for (int x = 0; x < w; ++x) {
            const long i = (long)y * w + x;
            const long ir = (long)y * w + imin(x + 1, w - 1);
            const long id = (long)imin(y + 1, h - 1) * w + x;
            const float u1x = u1[ir] - u1[i];
            const float u1y = u1[id] - u1[i];
            const float u2x = u2[ir] - u2[i];
            const float u2y = u2[id] - u2[i];
            const float g1 = hypotf(u1x, u1y);
            const float g2 = hypotf(u2x, u2y);
            const float ng1 = 1.0f + taut * g1;
            const float ng2 = 1.0f + taut * g2;
            p11[i] = (p11[i] + taut * u1x) / ng1;
            p12[i] = (p12[i] + taut * u1y) / ng1;
            p21[i] = (p21[i] + taut * u2x) / ng2;
            p22[i] = (p22[i] + taut * u2y) / ng2;
        }
    }
}

/* A.1: level sizes = round-half-even(dim * scaleStep) (saturate_cast<int>(double)); stop before a
 * level with cols < 16 or rows < 16 (that level is dropped). */
int orc_tvl1_level_sizes(int w, int h, const orc_tvl1_params *p, int *ws, int *hs) {
    int n = 1;
    ws[0] = w;
    hs[0] = h;
    for (int s = 1; s < p->nscales; ++s) {
        const int nw = (int)nearbyint((double)ws[s - 1] * p->scale_step);
        const int nh = (int)nearbyint((double)hs[s - 1] * p->scale_step);
        if (nw < 16 || nh < 16) break;
        ws[s] = nw;
        hs[s] = nh;
        n = s + 1;
    }
    return n;
}

int orc_tvl1_calc(const uint8_t *I0u8, const uint8_t *I1u8, int w, int h, const orc_tvl1_params *p, float *flow_xy,
                  int *iter_log) {
    if (w <= 0 || h <= 0 || p->nscales < 1 || p->nscales > 16) return -1;
    int ws[16], hs[16];
    const int nscales = orc_tvl1_level_sizes(w, h, p, ws, hs);
    if (iter_log) memset(iter_log, 0, sizeof(int) * (size_t)p->nscales * (size_t)p->warps);

    float *I0s[16], *I1s[16], *u1s[16], *u2s[16];
    for (int s = 0; s < nscales; ++s) {
        const size_t n = (size_t)ws[s] * hs[s];
        I0s[s] = (float *)malloc(n * sizeof(float));
        I1s[s] = (float *)malloc(n * sizeof(float));
        u1s[s] = (float *)calloc(n, sizeof(float));
        u2s[s] = (float *)calloc(n, sizeof(float));
    }
    const size_t n0 = (size_t)w * h;
    float *I1x = (float *)malloc(n0 * sizeof(float)), *I1y = (float *)malloc(n0 * sizeof(float));
    float *I1w = (float *)malloc(n0 * sizeof(float)), *I1wx = (float *)malloc(n0 * sizeof(float));
    float *I1wy = (float *)malloc(n0 * sizeof(float)), *grad = (float *)malloc(n0 * sizeof(float));
    float *rho_c = (float *)malloc(n0 * sizeof(float));
    float *p11 = (float *)malloc(n0 * sizeof(float)), *p12 = (float *)malloc(n0 * sizeof(float));
    float *p21 = (float *)malloc(n0 * sizeof(float)), *p22 = (float *)malloc(n0 * sizeof(float));

    /* A.1 pyramid: no pre-smoothing; fx = float(1/scaleStep) because dsize is empty */
    orc_u8_to_f32(I0u8, w, h, I0s[0]);
    orc_u8_to_f32(I1u8, w, h, I1s[0]);
    const float finv = (float)(1.0 / p->scale_step);
    for (int s = 1; s < nscales; ++s) {
        orc_resize_linear(I0s[s - 1], ws[s - 1], hs[s - 1], I0s[s], ws[s], hs[s], finv, finv, ORC_RESIZE_CUDA);
        orc_resize_linear(I1s[s - 1], ws[s - 1], hs[s - 1], I1s[s], ws[s], hs[s], finv, finv, ORC_RESIZE_CUDA);
    }

    const float l_t = (float)(p->lambda * p->theta);
    const float taut = (float)(p->tau / p->theta);
    const float theta = (float)p->theta;

    /* A.2: coarse -> fine; u = 0 at the coarsest (useInitialFlow = false) */
    for (int s = nscales - 1; s >= 0; --s) {
        const int W = ws[s], H = hs[s];
        const size_t n = (size_t)W * H;
        float *u1 = u1s[s], *u2 = u2s[s];
        orc_centered_gradient(I1s[s], W, H, I1x, I1y);
        memset(p11, 0, n * sizeof(float)); /* once per scale, not per warp */
        memset(p12, 0, n * sizeof(float));
        memset(p21, 0, n * sizeof(float));
        memset(p22, 0, n * sizeof(float));
        /* A.4 inner-loop control */
        const double scaled_eps = p->epsilon * p->epsilon * (double)((long)W * H);
        for (int wi = 0; wi < p->warps; ++wi) {
            orc_tvl1_warp_backward(I0s[s], I1s[s], I1x, I1y, u1, u2, W, H, I1w, I1wx, I1wy, grad, rho_c);
            double error = DBL_MAX;
            double prev_error = 0.0;
            int n_it = 0;
            for (; error > scaled_eps && n_it < p->iterations; ++n_it) {
                const int calc_error = (p->epsilon > 0) && (n_it & 1) && (prev_error < scaled_eps);
                const double e = orc_tvl1_estimate_u(I1wx, I1wy, grad, rho_c, p11, p12, p21, p22, u1, u2, W, H, l_t,
                                                     theta, calc_error);
                if (calc_error) {
                    error = e;
                    prev_error = error;
                } else {
                    error = DBL_MAX;
                    prev_error -= scaled_eps;
                }
                orc_tvl1_estimate_dual(u1, u2, p11, p12, p21, p22, W, H, taut);
            }
            if (iter_log) iter_log[s * p->warps + wi] = n_it;
        }
        /* A.2 step 4: upsample with explicit dsize, then multiply by float(1/scaleStep) */
        if (s > 0) {
            const float ufx = (float)(1.0 / ((double)ws[s - 1] / (double)W));
            const float ufy = (float)(1.0 / ((double)hs[s - 1] / (double)H));
            orc_resize_linear(u1, W, H, u1s[s - 1], ws[s - 1], hs[s - 1], ufx, ufy, ORC_RESIZE_CUDA);
            orc_resize_linear(u2, W, H, u2s[s - 1], ws[s - 1], hs[s - 1], ufx, ufy, ORC_RESIZE_CUDA);
            const float mul = (float)(1.0 / p->scale_step);
            const size_t nn = (size_t)ws[s - 1] * hs[s - 1];
            for (size_t i = 0; i < nn; ++i) {
                u1s[s - 1][i] = u1s[s - 1][i] * mul;
                u2s[s - 1][i] = u2s[s - 1][i] * mul;
            }
        }
    }
    /* A.5: merge -> interleaved (u, v) */
    for (size_t i = 0; i < n0; ++i) {
        flow_xy[2 * i] = u1s[0][i];
        flow_xy[2 * i + 1] = u2s[0][i];
    }
    for (int s = 0; s < nscales; ++s) {
        free(I0s[s]);
        free(I1s[s]);
        free(u1s[s]);
        free(u2s[s]);
    }
    free(I1x); free(I1y); free(I1w); free(I1wx); free(I1wy); free(grad); free(rho_c);
    free(p11); free(p12); free(p21); free(p22);
    return 0;
}
