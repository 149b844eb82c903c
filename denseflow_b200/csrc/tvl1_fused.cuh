// tvl1_fused.cuh — the persistent fused TV-L1 pair kernel (tvl1_fused.cu): job description shared
// between the host engine and the device code.
#pragma once

#include "tma.cuh"
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
    unsigned long long px_iters_total;  // sum over pairs of (level pixels x executed iterations)
    unsigned long long px_chunks_total; // sum over pairs of (level pixels x tile visits): 64 B of state/constant traffic each
    // CTA-0 view of where the last pair's time went, in ns (globaltimer): [0] level start, [1] warps,
    // [2] tile compute, [3] grid barriers, [4] upsample+merge, [8+s] tile compute at scale s, [16+s] chunks at scale s
    unsigned long long prof[32];
    // watchdog of the grid barrier: a CTA that polls the arrival counter kFusedStallPolls times (an L2 round trip each: 10-25 s)
    // records {1, blockIdx.x, barrier target, counter value, CTAs per lane, lane} here and traps, so a lost arrival surfaces
    // as a CUDA error instead of a hung device
    unsigned stall[8];
};
constexpr unsigned kFusedStallPolls = 1u << 25;

struct FusedJob {
    int nscales, warps, iterations, k;
    int flag_sync;  // 1: neighbour-warp progress counters in the tile loop, 0: CTA-wide barriers
    int use_tma;    // 1: constants + u tiles staged by TMA (cp.async.bulk.tensor.2d), 0: LDG -> STS
    int prefetch;   // 1: while a tile iterates, the CTA's next tile of the chunk is prefetched into L2 (TMA prefetch + prefetch.global.L2)
    // CUtensorMap[nscales][kFusedMapsPerLevel] in global memory: I1wx, I1wy, grad, rho_c, u1[0], u2[0], u1[1], u2[1]
    const void *tmaps;
    double epsilon;
    Tvl1Consts c;
    float up_mul;
    FusedLevel lv[16];
    float *I1x, *I1y, *I1wx, *I1wy, *grad, *rho_c;
    float *p[2][4];  // p11,p12,p21,p22 ping-pong
    double *partials;
    unsigned *sync;
    FusedHostCtl *ctl;
    int *iters_log;  // [16 * 16] executed inner iterations per (scale, warp) of THIS pair, index s * warps + w (mapped host memory)
    float *flow_xy;
    size_t flow_pitch_bytes;
    // bound > 0: emit the two quantised uint8 planes instead of the float2 field (PairJob)
    int bound;
    uint8_t *qx, *qy;
    size_t q_pitch;
};

// Several independent pairs ("lanes") per launch: lane i is solved by CTAs [i*group, (i+1)*group) with
// their own barrier words, partials and workspace.  Coarse pyramid levels have fewer tiles than the
// GPU has SMs; running pairs side by side keeps every SM busy without touching a pair's arithmetic.
constexpr int kFusedMaxLanes = 64;  // job descriptions travel through a device-memory ring, not the parameter bank
constexpr int kFusedMapsPerLevel = 8;
constexpr int kFusedParamLanes = 16;  // 16 jobs x 1.4 KB fit the kernel parameter bank (32 KB)
struct FusedBatch {
    int njobs, group;
    const FusedJob *jobs;  // njobs > kFusedParamLanes: [njobs] in device memory, each CTA copies its lane's job into shared memory
    const FusedJob *host_jobs;  // the same descriptions in host memory (always); copied into the parameter bank when they fit
};
// njobs <= kFusedParamLanes: the jobs travel in the parameter bank (uniform loads, no registers or shared memory spent on them —
// 4 % faster at 1080p than reading them from shared memory)
struct FusedBatchParams {
    int njobs, group;
    FusedJob job[kFusedParamLanes];
};

// tiles needed along one axis (see process_tile): region origins at multiples of T - 2h
__host__ __device__ inline int fused_tiles_along(int n, int T, int h) {
    const int stride = T - 2 * h;
    return n <= T ? 1 : (n - 2 * h + stride - 1) / stride;
}
#ifndef DFB_FUSED_THREADS
#define DFB_FUSED_THREADS 512
#endif
#ifndef DFB_FUSED_RPT
#define DFB_FUSED_RPT 4
#endif
constexpr int kFusedTileW = 128, kFusedTileH = DFB_FUSED_THREADS / 32 * DFB_FUSED_RPT;
// CTAs of the persistent kernel resident per SM: 512-thread CTAs own the whole SM (192 KB tile); 256-thread CTAs
// (128 x 32 tiles, 96 KB) run two to an SM, so one CTA's MUFU-bound dual half-steps and tile loads overlap the
// other's FMA-bound primal half-steps (the warps of ONE tile are forced into lock-step by their neighbour dependencies)
constexpr int kFusedCtasPerSm = DFB_FUSED_THREADS <= 256 ? 2 : 1;
constexpr int kFusedDefaultK = kFusedTileH >= 64 ? 8 : 4;

int fused_num_sms(int device);
inline int fused_cta_slots(int device) { return fused_num_sms(device) * kFusedCtasPerSm; }
// Encodes one 2-D fp32 tile descriptor (box 128 x 64, no swizzle, zero fill out of bounds) into out[128 bytes].
// plane: base pointer, extent w x h (elements / rows), row pitch in elements.
void fused_encode_tensor_map(void *out, const float *plane, int w, int h, int pitch);
// returns the number of kernels launched
int launch_tvl1_fused(const FusedBatch &batch, int device, cudaStream_t s, bool serialise);

}  // namespace dfb
