// tvl1_fused.cuh — the persistent fused TV-L1 pair kernel (tvl1_fused.cu): job description shared
// between the host engine and the device code.
#pragma once

#include "tvl1.cuh"

namespace dfb {

constexpr int kFusedMaxK = 8;  // most inner iterations kept on chip per tile visit (halo width)

struct FusedLevel {
    int w, h, pitch;
    const float *I0, *I1;  // pyramid level of frame a / frame b
    float *u1[2], *u2[2];  // flow, ping-pong
    float up_fx, up_fy;    // resize factors for upsampling THIS level's flow to level-1 (A.2 step 4)
};

// Written by the kernel into mapped host memory (no memcpy, no sync on the pair path).
struct FusedHostCtl {
    double error;         // unfused engine: convergence sum read by the host state machine
    int iters[16 * 16];   // executed inner iterations per (scale, warp) of the last pair
    unsigned long long px_iters_total;  // sum over pairs of (level pixels x executed iterations)
};

struct FusedJob {
    int nscales, warps, iterations, k;
    double epsilon;
    Tvl1Consts c;
    float up_mul;
    FusedLevel lv[16];
    float *I1x, *I1y, *I1wx, *I1wy, *grad, *rho_c;
    float *p[2][4];  // p11,p12,p21,p22 ping-pong
    double *partials;
    unsigned *sync;
    FusedHostCtl *ctl;
    float *flow_xy;
    size_t flow_pitch_bytes;
};

// returns the number of kernels launched
int launch_tvl1_fused(const FusedJob &job, int device, cudaStream_t s);

}  // namespace dfb
