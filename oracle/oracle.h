/*
 * oracle.h — CPU restatement of the reference's dense-optical-flow hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * library, and only as the checker / the timed CPU baseline.  The product path
 * (denseflow_b200/) never links, imports or falls back to it.
 *
 * What is restated: the arithmetic behind the two calls on the reference's hot path,
 *   alg_tvl1->calc(...)  /root/reference/src/denseflow_gpu.cpp:327  (cv::cuda::OpticalFlowDual_TVL1, defaults :299)
 *   alg_farn->calc(...)  /root/reference/src/denseflow_gpu.cpp:329  (cv::cuda::FarnebackOpticalFlow,  defaults :301)
 * and the host quantiser that follows it,
 *   convertFlowToImage   /root/reference/src/common.cpp:4-16.
 * The algorithm bodies live in a third-party dependency that is NOT under /root/reference:
 * OpenCV + opencv_contrib 4.5.2 (pinned at /root/reference/docker/Dockerfile:6), modules
 * cudaoptflow (tvl1flow.cpp, cuda/tvl1flow.cu, farneback.cpp, cuda/farneback.cu),
 * cudawarping (cuda/resize.cu), cudaarithm.  Their published algorithm is restated here as
 * recorded in SURVEY.md Appendix A (TV-L1) and Appendix B (Farneback).
 *
 * PINNING STATUS
 *   - quantiser:  pinned bit-exactly against the formula at src/common.cpp:6 (tests/golden/quantise_cases.npz, made by
 *                 tests/golden/make_golden.py; checked in tests/test_oracle_cpu.py).
 *   - Farneback:  pinned against real OpenCV code that runs in this image —
 *                 cv2.calcOpticalFlowFarneback(a,b,None,0.5,5,13,10,5,1.1,0) — with the resize
 *                 convention switched to OpenCV-CPU's half-pixel centres (ORC_RESIZE_HALF_PIXEL);
 *                 committed samples tests/golden/farneback_cv2_*.npz (tests/golden/make_golden.py) and a live comparison,
 *                 both in tests/test_oracle_cpu.py.
 *   - TV-L1:      PARITY UNPINNED.  No OpenCV build with cudaoptflow / optflow exists in this
 *                 image or on the GPU box, the reference has no tests or golden vectors, and
 *                 the reference itself cannot be compiled here (needs OpenCV-CUDA, Boost).
 *                 The restatement is only self-consistency checked (analytic ground-truth flow, an independently written
 *                 numpy restatement oracle/tvl1_numpy.py, and the frozen regression fixture tests/golden/tvl1_oracle_256.npz).
 */
#ifndef DENSEFLOW_ORACLE_H
#define DENSEFLOW_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- TV-L1 ------------------------------------------------------------ */

typedef struct {
    double tau;        /* 0.25  */
    double lambda;     /* 0.15  */
    double theta;      /* 0.3   */
    int    nscales;    /* 5     */
    int    warps;      /* 5     */
    double epsilon;    /* 0.01  */
    int    iterations; /* 300   */
    double scale_step; /* 0.8   */
} orc_tvl1_params;

void orc_tvl1_default_params(orc_tvl1_params *p);

/* building blocks, exposed so GPU kernels can be unit-checked one at a time */
void orc_u8_to_f32(const uint8_t *src, int w, int h, float *dst);
/* convention 0 = OpenCV-CUDA (src = dst*f, no half-pixel offset); 1 = OpenCV-CPU half-pixel centres */
#define ORC_RESIZE_CUDA 0
#define ORC_RESIZE_HALF_PIXEL 1
void orc_resize_linear(const float *src, int sw, int sh, float *dst, int dw, int dh, float fx, float fy,
                       int convention);
void orc_centered_gradient(const float *src, int w, int h, float *dx, float *dy);
void orc_tvl1_warp_backward(const float *I0, const float *I1, const float *I1x, const float *I1y, const float *u1,
                            const float *u2, int w, int h, float *I1w, float *I1wx, float *I1wy, float *grad,
                            float *rho_c);
/* returns sum of diff (double) when calc_error != 0, else 0 */
double orc_tvl1_estimate_u(const float *I1wx, const float *I1wy, const float *grad, const float *rho_c,
                           const float *p11, const float *p12, const float *p21, const float *p22, float *u1,
                           float *u2, int w, int h, float l_t, float theta, int calc_error);
void orc_tvl1_estimate_dual(const float *u1, const float *u2, float *p11, float *p12, float *p21, float *p22, int w,
                            int h, float taut);
/* pyramid level sizes; returns number of levels actually used (<= nscales) */
int orc_tvl1_level_sizes(int w, int h, const orc_tvl1_params *p, int *ws, int *hs);

/* full calc: I0,I1 u8 w*h (dense rows); flow_xy interleaved (u,v) float w*h*2.
 * iter_log (may be NULL): int[nscales*warps], executed inner iterations per (scale, warp),
 * index = s*warps + w with s = pyramid level (0 = finest).  Returns 0 on success. */
int orc_tvl1_calc(const uint8_t *I0, const uint8_t *I1, int w, int h, const orc_tvl1_params *p, float *flow_xy,
                  int *iter_log);

/* ---- Farneback -------------------------------------------------------- */

typedef struct {
    int    num_levels; /* 5   */
    double pyr_scale;  /* 0.5 */
    int    win_size;   /* 13  */
    int    num_iters;  /* 10  */
    int    poly_n;     /* 5   */
    double poly_sigma; /* 1.1 */
    int    resize_convention; /* ORC_RESIZE_CUDA for GPU parity, ORC_RESIZE_HALF_PIXEL to pin against cv2 CPU */
} orc_farn_params;

void orc_farn_default_params(orc_farn_params *p);
/* g[0..n], xg[0..n], xxg[0..n], ig = {ig11, ig03, ig33, ig55} */
void orc_farn_poly_constants(int n, double sigma, float *g, float *xg, float *xxg, float *ig);
/* level list: returns count; k index 0 = first processed (coarsest). */
int orc_farn_levels(int w, int h, const orc_farn_params *p, int *ws, int *hs, int *smooth, double *sigma);
void orc_farn_gaussian_blur(const float *src, int w, int h, int ksize, double sigma, float *dst);
void orc_farn_poly_exp(const float *src, int w, int h, int n, double sigma, float *R /* 5*h*w */);
void orc_farn_update_matrices(const float *flowx, const float *flowy, const float *R0, const float *R1, int w, int h,
                              float *M /* 5*h*w */);
void orc_farn_box_filter5(const float *src, int w, int h, int ksize, float *dst);
void orc_farn_update_flow(const float *M, int w, int h, float *flowx, float *flowy);
int orc_farn_calc(const uint8_t *I0, const uint8_t *I1, int w, int h, const orc_farn_params *p, float *flow_xy);

/* ---- quantiser (/root/reference/src/common.cpp:4-16) -------------------------------------- */
void orc_convert_flow_to_image(const float *flow_x, const float *flow_y, int w, int h, double lower, double higher,
                               uint8_t *img_x, uint8_t *img_y);
/* same on an interleaved CV_32FC2 flow (the split at /root/reference/src/denseflow_gpu.cpp:418 folded in) */
void orc_quantise_flow_xy(const float *flow_xy, int w, int h, int bound, uint8_t *img_x, uint8_t *img_y);

int orc_num_threads(void);
void orc_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
