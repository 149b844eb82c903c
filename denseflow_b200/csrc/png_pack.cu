// png_pack.cu — the GPU part of the packed-PNG flow format: convertFlowToPngImage,
// /root/reference/src/common.cpp:18-46.  For one CV_32FC2 flow field it produces the CV_8UC3 image the reference then
// hands to imencode(".png") (src/common.cpp:66-71):
//   minMaxLoc(flow_x) / minMaxLoc(flow_y)                                          (:23,:25)
//   bound = min(255*4, ceil((min(extent, max(|min|,|max|)) * 128/127) / 4) * 4), +4 when int(bound) % 8 == 0   (:24-32)
//   x = convertTo(CV_8U, alpha = float(1/(bound_x/128)), beta = 128), y likewise   (:33-39)
//   third channel: bound_x/4 in rows 0..int(h/2), bound_y/4 below                  (:40-42)
//   mixChannels -> (x, y, b) interleaved                                            (:43-45)
// Two kernels: a min/max reduction whose last block turns the extrema into the two bounds (double arithmetic, as the
// reference's), and the pack.  OpenCV's convertTo for CV_32F -> CV_8U computes saturate_cast<uchar>(src*alpha + beta) in
// float with a fused multiply-add on every x86 build with FMA3 (v_fma) and cvRound = round-half-to-even (the tests pin
// exactly that against cv2).  OpenCV's scalar tail (the last < 32 elements of a continuous image) multiplies and adds
// separately; the two differ only when they straddle a rounding boundary (3 values in 824 000 in a measurement here).
#include <cfloat>

#include "common.cuh"
#include "png_pack.h"

namespace dfb {

namespace {

struct MinMax {
    float min_x, max_x, min_y, max_y;
};

__device__ __forceinline__ MinMax mm_merge(MinMax a, MinMax b) {
    return MinMax{fminf(a.min_x, b.min_x), fmaxf(a.max_x, b.max_x), fminf(a.min_y, b.min_y), fmaxf(a.max_y, b.max_y)};
}

// scratch: [0 .. nblocks) partial MinMax, then one unsigned ticket; out: PngBounds
__global__ void __launch_bounds__(256) k_flow_minmax(const float *flow, size_t pitch_bytes, int w, int h, MinMax *partials, unsigned *ticket,
                                                    PngBounds *out) {
    MinMax m{FLT_MAX, -FLT_MAX, FLT_MAX, -FLT_MAX};
    const int total = w * h;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int y = i / w, x = i - y * w;
        const float2 f = reinterpret_cast<const float2 *>(reinterpret_cast<const char *>(flow) + (size_t)y * pitch_bytes)[x];
        m = mm_merge(m, MinMax{f.x, f.x, f.y, f.y});
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        MinMax n{__shfl_xor_sync(0xffffffffu, m.min_x, o), __shfl_xor_sync(0xffffffffu, m.max_x, o), __shfl_xor_sync(0xffffffffu, m.min_y, o),
                 __shfl_xor_sync(0xffffffffu, m.max_y, o)};
        m = mm_merge(m, n);
    }
    __shared__ MinMax sm[8];
    __shared__ bool last;
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < (int)(blockDim.x >> 5); ++k) m = mm_merge(m, sm[k]);
        partials[blockIdx.x] = m;
        __threadfence();
        last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {  // min / max are exact and order-free: any reduction order gives minMaxLoc's values
        __threadfence();
        MinMax t = partials[0];
        for (unsigned k = 1; k < gridDim.x; ++k) t = mm_merge(t, *(volatile MinMax *)&partials[k]);
        // src/common.cpp:24-32, in double like the reference (w, h, min_v, max_v are doubles there)
        auto bound_of = [](double extent, double mn, double mx) {
            double b = fmin(255. * 4, ceil((fmin(extent, fmax(fabs(mn), fabs(mx))) * 128. / 127.) / 4) * 4);
            if ((int)b % 8 == 0) b += 4;
            return b;
        };
        PngBounds o;
        o.bound_x = bound_of((double)w, (double)t.min_x, (double)t.max_x);
        o.bound_y = bound_of((double)h, (double)t.min_y, (double)t.max_y);
        o.min_x = t.min_x;
        o.max_x = t.max_x;
        o.min_y = t.min_y;
        o.max_y = t.max_y;
        *out = o;
        *ticket = 0;  // ready for the next launch
    }
}

// saturate_cast<uchar>(v * alpha + 128.f) as cv::Mat::convertTo does it for CV_32F sources: one fused multiply-add in
// float, cvRound (round-half-to-even, cvtss2si), then clamp to [0, 255]
__device__ __forceinline__ uint8_t cvt_u8(float v, float alpha) {
    const float t = fmaf(v, alpha, 128.f);
    const int q = __float2int_rn(t);
    return (uint8_t)min(max(q, 0), 255);
}

__global__ void k_flow_pack_png(const float *flow, size_t pitch_bytes, int w, int h, const PngBounds *bounds, uint8_t *bgr, size_t bgr_pitch) {
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= w || y >= h) return;
    const double base = 1. / 128.;
    const float ax = (float)(1. / (base * bounds->bound_x)), ay = (float)(1. / (base * bounds->bound_y));  // :33-34
    const float2 f = reinterpret_cast<const float2 *>(reinterpret_cast<const char *>(flow) + (size_t)y * pitch_bytes)[x];
    // rectangle(b, (0,0)-(w-1, half_h), bound_x/4); rectangle(b, (0, half_h+1)-(w-1, h-1), bound_y/4) with half_h = h / 2 as a
    // double truncated to int by cv::Point (:40-42); Scalar -> uchar is saturate_cast<uchar>(double)
    const int half_h = (int)((double)h / 2);
    const int split = (int)((double)h / 2 + 1);  // Point(0, half_h + 1): the double sum is truncated
    const double bv = y <= half_h && y < split ? bounds->bound_x / 4 : bounds->bound_y / 4;
    uint8_t *px = bgr + (size_t)y * bgr_pitch + 3 * (size_t)x;
    px[0] = cvt_u8(f.x, ax);
    px[1] = cvt_u8(f.y, ay);
    px[2] = (uint8_t)min(max(__double2int_rn(bv), 0), 255);
}

}  // namespace

size_t png_pack_scratch_bytes() { return sizeof(MinMax) * kPngPackBlocks + 64 + sizeof(PngBounds); }

void launch_flow_to_png_image(const float *flow_xy, size_t flow_pitch_bytes, int w, int h, uint8_t *bgr, size_t bgr_pitch, void *scratch,
                              PngBounds **bounds_dev, cudaStream_t s) {
    char *sc = static_cast<char *>(scratch);
    MinMax *partials = reinterpret_cast<MinMax *>(sc);
    unsigned *ticket = reinterpret_cast<unsigned *>(sc + sizeof(MinMax) * kPngPackBlocks);
    PngBounds *out = reinterpret_cast<PngBounds *>(sc + sizeof(MinMax) * kPngPackBlocks + 64);
    const int blocks = std::min(kPngPackBlocks, ceil_div(w * h, 256));
    k_flow_minmax<<<blocks, 256, 0, s>>>(flow_xy, flow_pitch_bytes, w, h, partials, ticket, out);
    DFB_KERNEL_CHECK();
    k_flow_pack_png<<<dim3(ceil_div(w, 32), ceil_div(h, 8)), dim3(32, 8), 0, s>>>(flow_xy, flow_pitch_bytes, w, h, out, bgr, bgr_pitch);
    DFB_KERNEL_CHECK();
    if (bounds_dev) *bounds_dev = out;
}

}  // namespace dfb
