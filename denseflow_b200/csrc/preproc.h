// preproc.h — BGR->gray and INTER_LINEAR uint8 resize, bit-exact to OpenCV's CPU path (see preproc.cu).
#pragma once

#include <vector>

#include "common.cuh"

namespace dfb {

struct ResizeTap {
    int idx;     // first source index (unclamped for rows)
    int a0, a1;  // 11-bit fixed-point weights of idx and idx + 1
};

void build_resize_taps(int dn, int sn, bool reset_at_borders, std::vector<ResizeTap> &out);
void launch_bgr_to_gray(const uint8_t *bgr, size_t bgr_pitch, int w, int h, uint8_t *gray, size_t gray_pitch, cudaStream_t s);
void launch_resize_u8(const uint8_t *src, size_t src_pitch, int sw, int sh, uint8_t *dst, size_t dst_pitch, int dw, int dh,
                      const ResizeTap *xt, const ResizeTap *yt, cudaStream_t s);

}  // namespace dfb
