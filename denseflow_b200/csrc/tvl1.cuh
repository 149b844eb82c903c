// tvl1.cuh — kernels of the TV-L1 path (what cv::cuda::OpticalFlowDual_TVL1::calc executes,
// /root/reference/src/denseflow_gpu.cpp:327; arithmetic per SURVEY.md Appendix A).
#pragma once

#include "common.cuh"

namespace dfb {

// ---- numeric policy -----------------------------------------------------------------------
// The reference's OpenCV is built with CUDA_FAST_MATH=ON (/root/reference/docker/Dockerfile:70):
// its divisions / hypotf are approximate.  Default build: FMA contraction on, rcp.approx /
// sqrt.approx in the inner loop.  -DDFB_STRICT_FP (+ -fmad=false): IEEE everywhere, used by the
// tests to separate restatement bugs from fp noise (bit-compatible with the CPU oracle up to the
// summation order of the convergence error).
#ifdef DFB_STRICT_FP
__device__ __forceinline__ float f_div(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float f_rcp(float a) { return __fdiv_rn(1.0f, a); }
__device__ __forceinline__ float f_hypot(float a, float b) { return hypotf(a, b); }
#else
__device__ __forceinline__ float f_rcp(float a) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
    return r;
}
__device__ __forceinline__ float f_div(float a, float b) { return a * f_rcp(b); }
__device__ __forceinline__ float f_sqrt(float s) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(s));
    return r;
}
__device__ __forceinline__ float f_hypot(float a, float b) { return f_sqrt(fmaf(a, a, b * b)); }
#endif

struct Tvl1Consts {
    float l_t;    // float(lambda * theta)
    float taut;   // float(tau / theta)
    float theta;  // float(theta)
};

// ---- stand-alone kernels (one launch per half-step: the reference's launch structure) -------
void launch_u8_to_f32(const uint8_t *src, size_t src_pitch_bytes, Plane dst, cudaStream_t s);
void launch_resize_linear(Plane src, Plane dst, float fx, float fy, float post_mul, cudaStream_t s);
// the two per-frame stages for several frames of one size at once (blockIdx.z = frame): each launch is latency-bound, a
// 64-frame clip needs 5 launches instead of 320
constexpr int kMaxFrameBatch = 16;
struct FramePtrs {
    const uint8_t *src[kMaxFrameBatch];
    float *base[kMaxFrameBatch];  // the frame's pyramid slot
};
void launch_u8_to_f32_batch(const FramePtrs &fp, int n, size_t src_pitch_bytes, int w, int h, int pitch, cudaStream_t s);
// level (src_off, sw, sh, sp) -> level (dst_off, dw, dh, dp) of every frame's slot
void launch_resize_linear_batch(const FramePtrs &fp, int n, size_t src_off, int sw, int sh, int sp, size_t dst_off, int dw, int dh, int dp, float fx,
                                float fy, cudaStream_t s);
void launch_centered_gradient(Plane src, Plane dx, Plane dy, cudaStream_t s);
void launch_warp_backward(Plane I0, Plane I1, Plane I1x, Plane I1y, Plane u1, Plane u2, Plane I1wx, Plane I1wy,
                          Plane grad, Plane rho_c, cudaStream_t s);
// err_partials: one double per block (grid size returned by estimate_u_blocks), or nullptr
int estimate_u_blocks(int w, int h);
void launch_estimate_u(Plane I1wx, Plane I1wy, Plane grad, Plane rho_c, Plane p11, Plane p12, Plane p21, Plane p22,
                       Plane u1, Plane u2, Tvl1Consts c, double *err_partials, cudaStream_t s);
void launch_estimate_dual(Plane u1, Plane u2, Plane p11, Plane p12, Plane p21, Plane p22, Tvl1Consts c,
                          cudaStream_t s);
// fixed-order sum of n partials into *out (device or mapped host memory)
void launch_sum_partials(const double *partials, int n, double *out, cudaStream_t s);
void launch_fill(Plane dst, float v, cudaStream_t s);
void launch_merge_flow(Plane u1, Plane u2, float *flow_xy, size_t flow_pitch_bytes, cudaStream_t s);
void launch_quantise(const float *flow_xy, size_t flow_pitch_bytes, int w, int h, int bound, uint8_t *qx, uint8_t *qy,
                     size_t q_pitch_bytes, cudaStream_t s);

}  // namespace dfb
