// preproc.cu — the per-frame conversions the reference's decode stage applies before the hot path
// (/root/reference/src/denseflow_gpu.cpp:163-170): cvtColor(BGR2GRAY) and cv::resize(INTER_LINEAR) on uint8
// (SURVEY §8 f3).  Both are bit-exact restatements of OpenCV's CPU fixed-point arithmetic, because the flow is
// computed on these pixels: a one-level difference in the gray frame moves the flow.
//   gray   = (B*3735 + G*19235 + R*9798 + 2^14) >> 15
//   resize : 11-bit coefficients cvRound(alpha*2048) (x coefficients reset at the borders, y rows index-clamped
//            with beta kept), horizontal pass in int, vertical (((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2) >> 2
#include <cmath>
#include <vector>

#include "preproc.h"

namespace dfb {

namespace {

__global__ void __launch_bounds__(256) k_bgr_to_gray(const uint8_t *__restrict__ bgr, size_t bgr_pitch, int w, int h,
                                                     uint8_t *__restrict__ gray, size_t gray_pitch) {
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= w || y >= h) return;
    const uint8_t *p = bgr + (size_t)y * bgr_pitch + 3 * x;
    gray[(size_t)y * gray_pitch + x] = (uint8_t)((p[0] * 3735 + p[1] * 19235 + p[2] * 9798 + 16384) >> 15);
}

__global__ void __launch_bounds__(256) k_resize_u8(const uint8_t *__restrict__ src, size_t src_pitch, int sw, int sh,
                                                   uint8_t *__restrict__ dst, size_t dst_pitch, int dw, int dh,
                                                   const ResizeTap *__restrict__ xt, const ResizeTap *__restrict__ yt) {
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= dw || y >= dh) return;
    const ResizeTap tx = xt[x], ty = yt[y];
    const int x1 = min(tx.idx + 1, sw - 1);
    const uint8_t *r0 = src + (size_t)min(max(ty.idx, 0), sh - 1) * src_pitch;
    const uint8_t *r1 = src + (size_t)min(max(ty.idx + 1, 0), sh - 1) * src_pitch;
    const int S0 = r0[tx.idx] * tx.a0 + r0[x1] * tx.a1;
    const int S1 = r1[tx.idx] * tx.a0 + r1[x1] * tx.a1;
    const int v = (((ty.a0 * (S0 >> 4)) >> 16) + ((ty.a1 * (S1 >> 4)) >> 16) + 2) >> 2;
    dst[(size_t)y * dst_pitch + x] = (uint8_t)min(max(v, 0), 255);
}

}  // namespace

void build_resize_taps(int dn, int sn, bool reset_at_borders, std::vector<ResizeTap> &out) {
    out.resize(dn);
    const double scale = 1.0 / ((double)dn / (double)sn);
    for (int d = 0; d < dn; ++d) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = (int)std::floor(f);
        f -= (float)s;
        if (reset_at_borders) {
            if (s < 0) {
                f = 0.f;
                s = 0;
            }
            if (s >= sn - 1) {
                f = 0.f;
                s = sn - 1;
            }
        }
        out[d].idx = s;
        out[d].a0 = (int)std::lrintf((1.f - f) * 2048.f);
        out[d].a1 = (int)std::lrintf(f * 2048.f);
    }
}

void launch_bgr_to_gray(const uint8_t *bgr, size_t bgr_pitch, int w, int h, uint8_t *gray, size_t gray_pitch, cudaStream_t s) {
    k_bgr_to_gray<<<dim3(ceil_div(w, 32), ceil_div(h, 8)), dim3(32, 8), 0, s>>>(bgr, bgr_pitch, w, h, gray, gray_pitch);
    DFB_KERNEL_CHECK();
}

void launch_resize_u8(const uint8_t *src, size_t src_pitch, int sw, int sh, uint8_t *dst, size_t dst_pitch, int dw, int dh,
                      const ResizeTap *xt, const ResizeTap *yt, cudaStream_t s) {
    k_resize_u8<<<dim3(ceil_div(dw, 32), ceil_div(dh, 8)), dim3(32, 8), 0, s>>>(src, src_pitch, sw, sh, dst, dst_pitch, dw, dh, xt, yt);
    DFB_KERNEL_CHECK();
}

}  // namespace dfb
