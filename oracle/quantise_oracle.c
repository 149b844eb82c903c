/*
 * quantise_oracle.c — restatement of convertFlowToImage, /root/reference/src/common.cpp:4-16
 * (CAST macro at :6), reached from encodeFlowMap (src/common.cpp:48-64, bounds passed as
 * -bound / +bound ints promoted to double at :53) after cv::split (src/denseflow_gpu.cpp:418).
 *
 * TEST INFRASTRUCTURE — see oracle.h.  Pinned bit-exactly: the macro is plain C arithmetic
 * (float promoted to double, left-to-right, cvRound = round-half-to-even) so this IS the formula.
 * NaN: both comparisons are false, cvRound(NaN) is INT_MIN on x86 (cvtsd2si), stored to uchar as 0.
 */
#include "oracle.h"

#include <math.h>

static inline uint8_t cast_px(float v, double L, double H) {
    if (v > H) return 255;
    if (v < L) return 0;
    const double q = 255 * ((double)v - L) / (H - L);
    if (q != q) return 0;              /* NaN: documented, not relied on */
    return (uint8_t)(int)nearbyint(q); /* cvRound: round-half-to-even under the default rounding mode */
}

void orc_convert_flow_to_image(const float *flow_x, const float *flow_y, int w, int h, double lower, double higher,
                               uint8_t *img_x, uint8_t *img_y) {
    const long n = (long)w * h;
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i) {
        img_x[i] = cast_px(flow_x[i], lower, higher);
        img_y[i] = cast_px(flow_y[i], lower, higher);
    }
}

void orc_quantise_flow_xy(const float *flow_xy, int w, int h, int bound, uint8_t *img_x, uint8_t *img_y) {
    const long n = (long)w * h;
    const double L = -bound, H = bound;
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i) {
        img_x[i] = cast_px(flow_xy[2 * i], L, H);
        img_y[i] = cast_px(flow_xy[2 * i + 1], L, H);
    }
}
