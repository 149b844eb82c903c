/*
 * farneback_oracle.c — CPU restatement (fp32 IEEE, no fast-math, no FMA contraction) of what
 * cv::cuda::FarnebackOpticalFlow::calc executes with the argument-less create() defaults the
 * reference uses (/root/reference/src/denseflow_gpu.cpp:301,329).
 *
 * TEST INFRASTRUCTURE — see oracle.h.  Pinned against cv2.calcOpticalFlowFarneback (real OpenCV
 * CPU code, same algorithm family and defaults) with resize_convention = ORC_RESIZE_HALF_PIXEL:
 * oracle/pin_farneback_cv2.py, tests/test_oracle_farneback.py.
 *
 * The arithmetic lives in OpenCV/opencv_contrib 4.5.2 (docker/Dockerfile:6), not in /root/reference:
 *   modules/cudaoptflow/src/farneback.cpp        level loop, constants  -> SURVEY.md Appendix B.1, B.2, B.5
 *   modules/cudaoptflow/src/cuda/farneback.cu    the five kernels       -> SURVEY.md Appendix B.2 - B.5
 *   modules/cudawarping/src/cuda/resize.cu       resize_linear          -> SURVEY.md Appendix A.1
 */
#include "oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

void orc_farn_default_params(orc_farn_params *p) {
    p->num_levels = 5;
    p->pyr_scale = 0.5;
    p->win_size = 13;
    p->num_iters = 10;
    p->poly_n = 5;
    p->poly_sigma = 1.1;
    p->resize_convention = ORC_RESIZE_CUDA;
}

/* cvRound == round-half-to-even on double */
static inline int cv_round(double v) { return (int)nearbyint(v); }

/* 6x6 Gauss-Jordan inverse in double (G is SPD; upstream uses Cholesky) */
static void inv6(double a[6][6], double inv[6][6]) {
    double m[6][12];
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) {
            m[i][j] = a[i][j];
            m[i][j + 6] = (i == j) ? 1.0 : 0.0;
        }
    for (int c = 0; c < 6; ++c) {
        int piv = c;
        for (int r = c + 1; r < 6; ++r)
            if (fabs(m[r][c]) > fabs(m[piv][c])) piv = r;
        if (piv != c)
            for (int j = 0; j < 12; ++j) {
                double t = m[c][j];
                m[c][j] = m[piv][j];
                m[piv][j] = t;
            }
        const double d = 1.0 / m[c][c];
        for (int j = 0; j < 12; ++j) m[c][j] *= d;
        for (int r = 0; r < 6; ++r)
            if (r != c) {
                const double f = m[r][c];
                if (f != 0.0)
                    for (int j = 0; j < 12; ++j) m[r][j] -= f * m[c][j];
            }
    }
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) inv[i][j] = m[i][j + 6];
}

/* B.3 constants (upstream prepareGaussian): 1-D weights normalised to sum 1 (stored fp32),
 * xg = j*g, xxg = j^2*g, 6x6 Gram matrix over (1, x, y, x^2, y^2, xy) inverted in double. */
void orc_farn_poly_constants(int n, double sigma, float *g_out, float *xg_out, float *xxg_out, float *ig) {
    float gbuf[2 * 16 + 1], xgbuf[2 * 16 + 1], xxgbuf[2 * 16 + 1];
    float *g = gbuf + n, *xg = xgbuf + n, *xxg = xxgbuf + n;
    if (sigma < 1.19209289550781250000e-7) sigma = n * 0.3;
    double s = 0.;
    for (int x = -n; x <= n; ++x) {
        g[x] = (float)exp(-x * x / (2 * sigma * sigma));
        s += g[x];
    }
    s = 1. / s;
    for (int x = -n; x <= n; ++x) {
        g[x] = (float)(g[x] * s);
        xg[x] = (float)(x * g[x]);
        xxg[x] = (float)(x * x * g[x]);
    }
    double G[6][6];
    memset(G, 0, sizeof(G));
    for (int y = -n; y <= n; ++y)
        for (int x = -n; x <= n; ++x) {
            G[0][0] += g[y] * g[x];
            G[1][1] += g[y] * g[x] * x * x;
            G[3][3] += g[y] * g[x] * x * x * x * x;
            G[5][5] += g[y] * g[x] * x * x * y * y;
        }
    G[2][2] = G[0][3] = G[0][4] = G[3][0] = G[4][0] = G[1][1];
    G[4][4] = G[3][3];
    G[3][4] = G[4][3] = G[5][5];
    double invG[6][6];
    inv6(G, invG);
    ig[0] = (float)invG[1][1]; /* ig11 */
    ig[1] = (float)invG[0][3]; /* ig03 */
    ig[2] = (float)invG[3][3]; /* ig33 */
    ig[3] = (float)invG[5][5]; /* ig55 */
    for (int k = 0; k <= n; ++k) {
        g_out[k] = g[k];
        xg_out[k] = xg[k];
        xxg_out[k] = xxg[k];
    }
}

/* B.1 level list.  Returned in processing order (index 0 = coarsest, last = full resolution). */
int orc_farn_levels(int w, int h, const orc_farn_params *p, int *ws, int *hs, int *smooth, double *sigmas) {
    int cropped = 0;
    double scale = 1.0;
    for (; cropped < p->num_levels; ++cropped) {
        scale *= p->pyr_scale;
        if (w * scale < 32 || h * scale < 32) break;
    }
    int cnt = 0;
    for (int k = cropped; k >= 0; --k) {
        scale = 1.0;
        for (int i = 0; i < k; ++i) scale *= p->pyr_scale;
        const double sigma = (1. / scale - 1) * 0.5;
        int ss = cv_round(sigma * 5) | 1;
        ss = imax(ss, 3);
        ws[cnt] = cv_round(w * scale);
        hs[cnt] = cv_round(h * scale);
        smooth[cnt] = ss;
        sigmas[cnt] = sigma;
        ++cnt;
    }
    return cnt;
}

/* cv::getGaussianKernel(ksize, sigma, CV_32F): fixed table for small odd ksize with sigma <= 0,
 * else exp(-x^2 / 2 sigma^2) normalised in double and cast to float. Returns half kernel k[0..half]. */
static void gaussian_half_kernel(int ksize, double sigma, float *half_k) {
    const int half = ksize / 2;
    static const float tab1[] = {1.f};
    static const float tab3[] = {0.25f, 0.5f, 0.25f};
    static const float tab5[] = {0.0625f, 0.25f, 0.375f, 0.25f, 0.0625f};
    static const float tab7[] = {0.03125f, 0.109375f, 0.21875f, 0.28125f, 0.21875f, 0.109375f, 0.03125f};
    const float *fixed = NULL;
    if (ksize % 2 == 1 && ksize <= 7 && sigma <= 0)
        fixed = ksize == 1 ? tab1 : ksize == 3 ? tab3 : ksize == 5 ? tab5 : tab7;
    if (fixed) {
        for (int i = 0; i <= half; ++i) half_k[i] = fixed[half + i];
        return;
    }
    const double sx = sigma > 0 ? sigma : ((ksize - 1) * 0.5 - 1) * 0.3 + 0.8;
    const double scale2x = -0.5 / (sx * sx);
    double tmp[129];
    double sum = 0;
    for (int i = 0; i < ksize; ++i) {
        const double x = i - (ksize - 1) * 0.5;
        tmp[i] = exp(scale2x * x * x);
        sum += tmp[i];
    }
    sum = 1. / sum;
    for (int i = 0; i <= half; ++i) half_k[i] = (float)(tmp[half + i] * sum);
}

static inline int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) {
        if (i < 0) i = -i;
        else i = 2 * (n - 1) - i;
    }
    return i;
}

/* B.2: separable Gaussian, BORDER_REFLECT_101; vertical pass first, symmetric taps paired as
 * (a + b) * k, centre tap first — the order of upstream's gaussianBlur kernel. */
void orc_farn_gaussian_blur(const float *src, int w, int h, int ksize, double sigma, float *dst) {
    const int half = ksize / 2;
    float kern[65];
    gaussian_half_kernel(ksize, sigma, kern);
    float *tmp = (float *)malloc((size_t)w * h * sizeof(float));
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            float acc = src[(long)y * w + x] * kern[0];
            for (int j = 1; j <= half; ++j)
                acc = acc + (src[(long)reflect101(y - j, h) * w + x] + src[(long)reflect101(y + j, h) * w + x]) * kern[j];
            tmp[(long)y * w + x] = acc;
        }
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            float acc = tmp[(long)y * w + x] * kern[0];
            for (int i = 1; i <= half; ++i)
                acc = acc + (tmp[(long)y * w + reflect101(x - i, w)] + tmp[(long)y * w + reflect101(x + i, w)]) * kern[i];
            dst[(long)y * w + x] = acc;
        }
    free(tmp);
}

/* B.3: polynomial expansion, index-clamped borders, vertical pass (t0,t1,t2) then horizontal. */
void orc_farn_poly_exp(const float *src, int w, int h, int n, double sigma, float *R) {
    float g[17], xg[17], xxg[17], ig[4];
    orc_farn_poly_constants(n, sigma, g, xg, xxg, ig);
    const float ig11 = ig[0], ig03 = ig[1], ig33 = ig[2], ig55 = ig[3];
    const size_t plane = (size_t)w * h;
    float *t = (float *)malloc(3 * plane * sizeof(float));
    float *t0p = t, *t1p = t + plane, *t2p = t + 2 * plane;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            float a0 = src[(long)y * w + x] * g[0];
            float a1 = 0.f, a2 = 0.f;
            for (int k = 1; k <= n; ++k) {
                const float s0 = src[(long)imax(y - k, 0) * w + x];
                const float s1 = src[(long)imin(y + k, h - 1) * w + x];
                a0 = a0 + g[k] * (s0 + s1);
                a1 = a1 + xg[k] * (s1 - s0);
                a2 = a2 + xxg[k] * (s0 + s1);
            }
            t0p[(long)y * w + x] = a0;
            t1p[(long)y * w + x] = a1;
            t2p[(long)y * w + x] = a2;
        }
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            const float *r0 = t0p + (long)y * w, *r1 = t1p + (long)y * w, *r2 = t2p + (long)y * w;
            float b1 = g[0] * r0[x], b3 = g[0] * r1[x], b5 = g[0] * r2[x];
            float b2 = 0.f, b4 = 0.f, b6 = 0.f;
            for (int k = 1; k <= n; ++k) {
                const int xl = imax(x - k, 0), xr = imin(x + k, w - 1);
                float s = r0[xr] + r0[xl];
                b1 = b1 + s * g[k];
                b4 = b4 + s * xxg[k];
                b2 = b2 + (r0[xr] - r0[xl]) * xg[k];
                s = r1[xr] + r1[xl];
                b3 = b3 + s * g[k];
                b6 = b6 + (r1[xr] - r1[xl]) * xg[k];
                s = r2[xr] + r2[xl];
                b5 = b5 + s * g[k];
            }
            const long i = (long)y * w + x;
            R[0 * plane + i] = b3 * ig11;
            R[1 * plane + i] = b2 * ig11;
            R[2 * plane + i] = b1 * ig03 + b5 * ig33;
            R[3 * plane + i] = b1 * ig03 + b4 * ig33;
            R[4 * plane + i] = b6 * ig55;
        }
    free(t);
}

/* B.4 */
void orc_farn_update_matrices(const float *flowx, const float *flowy, const float *R0, const float *R1, int w, int h,
                              float *M) {
    static const float c_border[6] = {0.14f, 0.14f, 0.4472f, 0.4472f, 0.4472f, 1.f};
    const size_t plane = (size_t)w * h;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            const long i = (long)y * w + x;
            const float dx = flowx[i], dy = flowy[i];
            float fx = (float)x + dx;
            float fy = (float)y + dy;
            const int x1 = (int)floorf(fx);
            const int y1 = (int)floorf(fy);
            fx -= (float)x1;
            fy -= (float)y1;
            float r2, r3, r4, r5, r6;
            if (x1 >= 0 && y1 >= 0 && x1 < w - 1 && y1 < h - 1) {
                const float a00 = (1.f - fx) * (1.f - fy);
                const float a01 = fx * (1.f - fy);
                const float a10 = (1.f - fx) * fy;
                const float a11 = fx * fy;
                const long j = (long)y1 * w + x1;
                r2 = a00 * R1[0 * plane + j] + a01 * R1[0 * plane + j + 1] + a10 * R1[0 * plane + j + w] + a11 * R1[0 * plane + j + w + 1];
                r3 = a00 * R1[1 * plane + j] + a01 * R1[1 * plane + j + 1] + a10 * R1[1 * plane + j + w] + a11 * R1[1 * plane + j + w + 1];
                r4 = a00 * R1[2 * plane + j] + a01 * R1[2 * plane + j + 1] + a10 * R1[2 * plane + j + w] + a11 * R1[2 * plane + j + w + 1];
                r5 = a00 * R1[3 * plane + j] + a01 * R1[3 * plane + j + 1] + a10 * R1[3 * plane + j + w] + a11 * R1[3 * plane + j + w + 1];
                r6 = a00 * R1[4 * plane + j] + a01 * R1[4 * plane + j + 1] + a10 * R1[4 * plane + j + w] + a11 * R1[4 * plane + j + w + 1];
                r4 = (R0[2 * plane + i] + r4) * 0.5f;
                r5 = (R0[3 * plane + i] + r5) * 0.5f;
                r6 = (R0[4 * plane + i] + r6) * 0.25f;
            } else {
                r2 = r3 = 0.f;
                r4 = R0[2 * plane + i];
                r5 = R0[3 * plane + i];
                r6 = R0[4 * plane + i] * 0.5f;
            }
            r2 = (R0[0 * plane + i] - r2) * 0.5f;
            r3 = (R0[1 * plane + i] - r3) * 0.5f;
            r2 = r2 + (r4 * dy + r6 * dx);
            r3 = r3 + (r6 * dy + r5 * dx);
            const float scale = c_border[imin(x, 5)] * c_border[imin(y, 5)] * c_border[imin(w - x - 1, 5)] *
                                c_border[imin(h - y - 1, 5)];
            r2 *= scale;
            r3 *= scale;
            r4 *= scale;
            r5 *= scale;
            r6 *= scale;
            M[0 * plane + i] = r4 * r4 + r6 * r6;
            M[1 * plane + i] = (r4 + r5) * r6;
            M[2 * plane + i] = r5 * r5 + r6 * r6;
            M[3 * plane + i] = r4 * r2 + r6 * r3;
            M[4 * plane + i] = r6 * r2 + r5 * r3;
        }
}

/* B.5 box mean (ksize x ksize) of each of the 5 planes, index-clamped; vertical sums first. */
void orc_farn_box_filter5(const float *src, int w, int h, int ksize, float *dst) {
    const int half = ksize / 2;
    const float area_inv = 1.f / (float)((1 + 2 * half) * (1 + 2 * half));
    const size_t plane = (size_t)w * h;
    float *tmp = (float *)malloc(plane * sizeof(float));
    for (int k = 0; k < 5; ++k) {
        const float *s = src + k * plane;
        float *d = dst + k * plane;
#pragma omp parallel for schedule(static)
        for (int y = 0; y < h; ++y)
            for (int x = 0; x < w; ++x) {
                float acc = s[(long)y * w + x];
                for (int j = 1; j <= half; ++j)
                    acc = acc + (s[(long)imax(y - j, 0) * w + x] + s[(long)imin(y + j, h - 1) * w + x]);
                tmp[(long)y * w + x] = acc;
            }
#pragma omp parallel for schedule(static)
        for (int y = 0; y < h; ++y)
            for (int x = 0; x < w; ++x) {
                float acc = tmp[(long)y * w + x];
                for (int i = 1; i <= half; ++i)
                    acc = acc + (tmp[(long)y * w + imax(x - i, 0)] + tmp[(long)y * w + imin(x + i, w - 1)]);
                d[(long)y * w + x] = acc * area_inv;
            }
    }
    free(tmp);
}

/* B.5 per-pixel 2x2 solve */
void orc_farn_update_flow(const float *M, int w, int h, float *flowx, float *flowy) {
    const size_t plane = (size_t)w * h;
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)plane; ++i) {
        const float g11 = M[i], g12 = M[plane + i], g22 = M[2 * plane + i];
        const float h1 = M[3 * plane + i], h2 = M[4 * plane + i];
        const float det_inv = 1.f / (g11 * g22 - g12 * g12 + 1e-3f);
        flowx[i] = (g11 * h2 - g12 * h1) * det_inv;
        flowy[i] = (g22 * h1 - g12 * h2) * det_inv;
    }
}

int orc_farn_calc(const uint8_t *I0u8, const uint8_t *I1u8, int w, int h, const orc_farn_params *p, float *flow_xy) {
    if (w <= 0 || h <= 0 || p->poly_n > 16 || p->num_levels > 15) return -1;
    int ws[16], hs[16], smooth[16];
    double sigmas[16];
    const int nlev = orc_farn_levels(w, h, p, ws, hs, smooth, sigmas);
    const size_t n0 = (size_t)w * h;
    float *frame[2], *blurred = (float *)malloc(n0 * sizeof(float)), *img = (float *)malloc(n0 * sizeof(float));
    frame[0] = (float *)malloc(n0 * sizeof(float));
    frame[1] = (float *)malloc(n0 * sizeof(float));
    orc_u8_to_f32(I0u8, w, h, frame[0]);
    orc_u8_to_f32(I1u8, w, h, frame[1]);
    float *R[2];
    R[0] = (float *)malloc(5 * n0 * sizeof(float));
    R[1] = (float *)malloc(5 * n0 * sizeof(float));
    float *M = (float *)malloc(5 * n0 * sizeof(float)), *bufM = (float *)malloc(5 * n0 * sizeof(float));
    float *fx = (float *)malloc(n0 * sizeof(float)), *fy = (float *)malloc(n0 * sizeof(float));
    float *pfx = (float *)malloc(n0 * sizeof(float)), *pfy = (float *)malloc(n0 * sizeof(float));
    int pw = 0, ph = 0;
    for (int l = 0; l < nlev; ++l) {
        const int W = ws[l], H = hs[l];
        const size_t n = (size_t)W * H;
        if (l == 0) {
            memset(fx, 0, n * sizeof(float));
            memset(fy, 0, n * sizeof(float));
        } else {
            const float rfx = (float)(1.0 / ((double)W / (double)pw));
            const float rfy = (float)(1.0 / ((double)H / (double)ph));
            orc_resize_linear(pfx, pw, ph, fx, W, H, rfx, rfy, p->resize_convention);
            orc_resize_linear(pfy, pw, ph, fy, W, H, rfx, rfy, p->resize_convention);
            const float mul = (float)(1.0 / p->pyr_scale);
            for (size_t i = 0; i < n; ++i) {
                fx[i] = fx[i] * mul;
                fy[i] = fy[i] * mul;
            }
        }
        for (int i = 0; i < 2; ++i) {
            orc_farn_gaussian_blur(frame[i], w, h, smooth[l], sigmas[l], blurred);
            const float rfx = (float)(1.0 / ((double)W / (double)w));
            const float rfy = (float)(1.0 / ((double)H / (double)h));
            orc_resize_linear(blurred, w, h, img, W, H, rfx, rfy, p->resize_convention);
            orc_farn_poly_exp(img, W, H, p->poly_n, p->poly_sigma, R[i]);
        }
        orc_farn_update_matrices(fx, fy, R[0], R[1], W, H, M);
        for (int it = 0; it < p->num_iters; ++it) {
            orc_farn_box_filter5(M, W, H, p->win_size, bufM);
            orc_farn_update_flow(bufM, W, H, fx, fy);
            if (it < p->num_iters - 1) orc_farn_update_matrices(fx, fy, R[0], R[1], W, H, M);
        }
        memcpy(pfx, fx, n * sizeof(float));
        memcpy(pfy, fy, n * sizeof(float));
        pw = W;
        ph = H;
    }
    for (size_t i = 0; i < n0; ++i) {
        flow_xy[2 * i] = fx[i];
        flow_xy[2 * i + 1] = fy[i];
    }
    free(frame[0]); free(frame[1]); free(blurred); free(img); free(R[0]); free(R[1]);
    free(M); free(bufM); free(fx); free(fy); free(pfx); free(pfy);
    return 0;
}
