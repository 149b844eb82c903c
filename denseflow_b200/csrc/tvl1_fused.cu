// tvl1_fused.cu — the persistent fused TV-L1 pair kernel for sm_100a.
//
// One cooperative launch computes a whole flow field: everything procOneScale does for every
// scale of cv::cuda::OpticalFlowDual_TVL1::calc (/root/reference/src/denseflow_gpu.cpp:327;
// SURVEY.md Appendix A.2 - A.5) — centred gradients, 5 bicubic warps per scale, the primal/dual
// inner loop with its data-dependent length, the flow upsampling between scales and the final
// merge — with the A.4 convergence state machine evaluated on the device.  The reference needs
// ~2 000 launches and a host sync per convergence check for the same work.
//
// Inner loop (the hot part).  The image is cut into 128 x 64 register tiles.  A CTA of 16 warps
// owns one tile at a time: lane l of warp q holds pixels x = 4l..4l+3 of rows 4q..4q+3 — the six
// state planes (u1,u2,p11,p12,p21,p22) live in registers, the four per-warp constants
// (I1wx,I1wy,grad,rho_c) in shared memory.  Horizontal neighbours come from warp shuffles,
// vertical neighbours from the thread's own next/previous row or, across warps, from a one-row
// shared-memory exchange.  k primal+dual iterations run on chip per tile visit (halo = k, the
// valid region shrinks by one pixel per half-step pair, exactly preserving the reference's Jacobi
// ordering: dual sees the fully updated u, the next primal the fully updated p), then the
// interior is written to the other half of a ping-pong pair.  Compulsory HBM traffic is
// 64 B/px per k iterations instead of 88 B/px per iteration.
//
// Grid-wide ordering uses a monotonically counting barrier in global memory (all CTAs are
// co-resident: cooperative launch).  The convergence error is reduced in a fixed order
// (per-CTA partial -> every CTA sums all partials identically), so every CTA takes the same
// branch of the A.4 state machine and results are run-to-run deterministic.
#include <cfloat>

#include "tvl1_fused.cuh"
#include "tvl1_math.cuh"

namespace dfb {

namespace {

constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
constexpr int RPT = 4;            // rows per thread
constexpr int TW = 128;           // tile width  = 32 lanes x 4 px
constexpr int TH = kWarps * RPT;  // tile height = 64
constexpr int kConstPlane = TW * TH;

struct Smem {
    float consts[4][kConstPlane];  // I1wx, I1wy, grad, rho_c of the current tile (thread-private slots)
    float u_top[2][kWarps][TW];    // row 0 of every warp's u1/u2 (read by the warp above as "down")
    float p_bot[2][kWarps][TW];    // row 3 of every warp's p12/p22 (read by the warp below as "up")
    double red[kWarps];
    double bcast[4];
    int ibcast[4];
};

__device__ __forceinline__ float4 ld_cg4(const float *p) { return __ldcg(reinterpret_cast<const float4 *>(p)); }
__device__ __forceinline__ void st4(float *p, float4 v) { *reinterpret_cast<float4 *>(p) = v; }
__device__ __forceinline__ float4 zero4() { return make_float4(0.f, 0.f, 0.f, 0.f); }

// ---- grid barrier ---------------------------------------------------------------------------
// sync[0] counts arrivals monotonically: barrier number b completes when it reaches b * gridDim.x.
// bar.sync orders the CTA's writes before thread 0's gpu-scope fence + atomic (release); the
// fence after the spin makes other CTAs' writes visible and invalidates this SM's L1.
__device__ __forceinline__ void grid_barrier(unsigned *counter, unsigned &epoch) {
    __syncthreads();
    if (threadIdx.x == 0) {
        epoch += gridDim.x;
        __threadfence();
        atomicAdd(counter, 1u);
        unsigned v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
        } while ((int)(v - epoch) < 0);
        __threadfence();
    }
    __syncthreads();
}

// ---- pixel-parallel phases (grid-stride over the level) --------------------------------------
__device__ __forceinline__ void phase_level_start(const FusedJob &job, const FusedLevel &L, bool coarsest) {
    const int W = L.w, H = L.h, P = L.pitch;
    const int total = H * (P >> 2);
    for (int i = blockIdx.x * kThreads + threadIdx.x; i < total; i += gridDim.x * kThreads) {
        const int y = i / (P >> 2), x0 = (i - y * (P >> 2)) << 2;
        const size_t o = (size_t)y * P + x0;
        // A.2 step 2: p = 0 once per scale; A.2: u = 0 at the coarsest scale
        st4(job.p[0][0] + o, zero4());
        st4(job.p[0][1] + o, zero4());
        st4(job.p[0][2] + o, zero4());
        st4(job.p[0][3] + o, zero4());
        if (coarsest) {
            st4(L.u1[0] + o, zero4());
            st4(L.u2[0] + o, zero4());
        }
        // A.2 step 1: centred gradient of I1, index-clamped
        if (x0 < W) {
            const float *row = L.I1 + (size_t)y * P;
            const float *up = L.I1 + (size_t)max(y - 1, 0) * P;
            const float *dn = L.I1 + (size_t)min(y + 1, H - 1) * P;
            float gx[4], gy[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int x = min(x0 + j, W - 1);
                gx[j] = 0.5f * (__ldg(row + min(x + 1, W - 1)) - __ldg(row + max(x - 1, 0)));
                gy[j] = 0.5f * (__ldg(dn + x) - __ldg(up + x));
            }
            st4(job.I1x + o, make_float4(gx[0], gx[1], gx[2], gx[3]));
            st4(job.I1y + o, make_float4(gy[0], gy[1], gy[2], gy[3]));
        }
    }
}

// A.2 "Warp (warpBackward)" for the whole level: reads u[cur], writes the four per-warp constants.
__device__ __forceinline__ void phase_warp(const FusedJob &job, const FusedLevel &L, int cur) {
    const int W = L.w, H = L.h, P = L.pitch;
    const float *u1 = L.u1[cur], *u2 = L.u2[cur];
    const int total = H * W;
    for (int i = blockIdx.x * kThreads + threadIdx.x; i < total; i += gridDim.x * kThreads) {
        const int y = i / W, x = i - y * W;
        const size_t o = (size_t)y * P + x;
        const float u1v = __ldcg(u1 + o), u2v = __ldcg(u2 + o);
        const float wx = x + u1v, wy = y + u2v;
        const int xmin = (int)ceilf(wx - 2.0f), xmax = (int)floorf(wx + 2.0f);
        const int ymin = (int)ceilf(wy - 2.0f), ymax = (int)floorf(wy + 2.0f);
        float sum = 0.f, sumx = 0.f, sumy = 0.f, wsum = 0.f;
        for (int cy = ymin; cy <= ymax; ++cy) {
            const float wyc = bicubic_coeff(wy - cy);
            const size_t ro = (size_t)max(0, min(cy, H - 1)) * P;
            for (int cx = xmin; cx <= xmax; ++cx) {
                const float wgt = bicubic_coeff(wx - cx) * wyc;
                const size_t t = ro + max(0, min(cx, W - 1));
                sum = sum + wgt * __ldg(L.I1 + t);
                sumx = sumx + wgt * __ldcg(job.I1x + t);
                sumy = sumy + wgt * __ldcg(job.I1y + t);
                wsum = wsum + wgt;
            }
        }
        const float coeff = f_rcp(wsum);
        const float I1wv = sum * coeff, ix = sumx * coeff, iy = sumy * coeff;
        job.I1wx[o] = ix;
        job.I1wy[o] = iy;
        job.grad[o] = ix * ix + iy * iy;
        job.rho_c[o] = I1wv - ix * u1v - iy * u2v - __ldg(L.I0 + o);
    }
}

// A.2 step 4: upsample this level's flow to the next finer level (explicit dsize), x float(1/scaleStep)
__device__ __forceinline__ void phase_upsample(const FusedJob &job, const FusedLevel &L, const FusedLevel &F, int cur) {
    const int total = F.h * F.w;
    for (int i = blockIdx.x * kThreads + threadIdx.x; i < total; i += gridDim.x * kThreads) {
        const int dy = i / F.w, dx = i - dy * F.w;
        const float sx = dx * L.up_fx, sy = dy * L.up_fy;
        const int x1 = __float2int_rd(sx), y1 = __float2int_rd(sy);
        const int x2 = x1 + 1, y2 = y1 + 1;
        const size_t r1 = (size_t)min(y1, L.h - 1) * L.pitch, r2 = (size_t)min(y2, L.h - 1) * L.pitch;
        const int x1r = min(x1, L.w - 1), x2r = min(x2, L.w - 1);
        const float w11 = (x2 - sx) * (y2 - sy), w12 = (sx - x1) * (y2 - sy);
        const float w21 = (x2 - sx) * (sy - y1), w22 = (sx - x1) * (sy - y1);
        const float *a = L.u1[cur], *b = L.u2[cur];
        float o1 = 0.f, o2 = 0.f;
        o1 = o1 + __ldcg(a + r1 + x1r) * w11;
        o1 = o1 + __ldcg(a + r1 + x2r) * w12;
        o1 = o1 + __ldcg(a + r2 + x1r) * w21;
        o1 = o1 + __ldcg(a + r2 + x2r) * w22;
        o2 = o2 + __ldcg(b + r1 + x1r) * w11;
        o2 = o2 + __ldcg(b + r1 + x2r) * w12;
        o2 = o2 + __ldcg(b + r2 + x1r) * w21;
        o2 = o2 + __ldcg(b + r2 + x2r) * w22;
        F.u1[0][(size_t)dy * F.pitch + dx] = o1 * job.up_mul;
        F.u2[0][(size_t)dy * F.pitch + dx] = o2 * job.up_mul;
    }
}

// A.5: merge(u1,u2) -> CV_32FC2
__device__ __forceinline__ void phase_merge(const FusedJob &job, const FusedLevel &L, int cur) {
    const int total = L.h * L.w;
    for (int i = blockIdx.x * kThreads + threadIdx.x; i < total; i += gridDim.x * kThreads) {
        const int y = i / L.w, x = i - y * L.w;
        const size_t o = (size_t)y * L.pitch + x;
        float2 *row = reinterpret_cast<float2 *>(reinterpret_cast<char *>(job.flow_xy) + (size_t)y * job.flow_pitch_bytes);
        row[x] = make_float2(__ldcg(L.u1[cur] + o), __ldcg(L.u2[cur] + o));
    }
}

// ---- the register-tile inner loop -------------------------------------------------------------
// Runs kk primal+dual iterations on the tile whose region origin is (rx0, ry0) (image coords, may be
// negative), reading state from buffers [cur] and writing the interior (region shrunk by hx / hy) to
// [cur^1].  Returns this thread's share of sum(diff) of the last primal step when `check`.
__device__ __forceinline__ double process_tile(const FusedJob &job, const FusedLevel &L, int cur, int rx0, int ry0,
                                               int kk, int hx, int hy, bool check, Smem &sm) {
    const int lane = threadIdx.x & 31, wq = threadIdx.x >> 5;
    const int W = L.w, H = L.h, P = L.pitch;
    const int gx0 = rx0 + 4 * lane;
    const int gy0 = ry0 + RPT * wq;
    const Tvl1Consts c = job.c;

    const float *src[6] = {L.u1[cur], L.u2[cur], job.p[cur][0], job.p[cur][1], job.p[cur][2], job.p[cur][3]};
    float4 u1[RPT], u2[RPT], p11[RPT], p12[RPT], p21[RPT], p22[RPT];
    const bool col_ok = gx0 >= 0 && gx0 < P;
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
        const int gy = gy0 + r;
        const bool ok = col_ok && gy >= 0 && gy < H;
        const size_t o = ok ? (size_t)gy * P + gx0 : 0;
        u1[r] = ok ? ld_cg4(src[0] + o) : zero4();
        u2[r] = ok ? ld_cg4(src[1] + o) : zero4();
        p11[r] = ok ? ld_cg4(src[2] + o) : zero4();
        p12[r] = ok ? ld_cg4(src[3] + o) : zero4();
        p21[r] = ok ? ld_cg4(src[4] + o) : zero4();
        p22[r] = ok ? ld_cg4(src[5] + o) : zero4();
        const int so = (RPT * wq + r) * TW + 4 * lane;
        st4(&sm.consts[0][so], ok ? ld_cg4(job.I1wx + o) : zero4());
        st4(&sm.consts[1][so], ok ? ld_cg4(job.I1wy + o) : zero4());
        st4(&sm.consts[2][so], ok ? ld_cg4(job.grad + o) : zero4());
        st4(&sm.consts[3][so], ok ? ld_cg4(job.rho_c + o) : zero4());
    }
    // border flags (image coordinates).  gx0 is a multiple of 4, so x == 0 can only be lane-pixel 0.
    const bool left_edge = gx0 == 0;
    const int jlast = W - 1 - gx0;  // pixel j == jlast is the last image column (forward diff clamps)
    bool top_edge[RPT], bot_edge[RPT];
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
        top_edge[r] = gy0 + r == 0;
        bot_edge[r] = gy0 + r == H - 1;
    }

    // make row 3 of p12/p22 visible to the warp below before the first primal step
    st4(&sm.p_bot[0][wq][4 * lane], p12[RPT - 1]);
    st4(&sm.p_bot[1][wq][4 * lane], p22[RPT - 1]);
    __syncthreads();

    double err = 0.0;
    for (int it = 0; it < kk; ++it) {
        const bool do_err = check && it == kk - 1;
        // -------- primal: u <- u + d(rho) + theta * div p --------------------------------------
        float4 up12 = wq > 0 ? *reinterpret_cast<const float4 *>(&sm.p_bot[0][wq - 1][4 * lane]) : zero4();
        float4 up22 = wq > 0 ? *reinterpret_cast<const float4 *>(&sm.p_bot[1][wq - 1][4 * lane]) : zero4();
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            const int so = (RPT * wq + r) * TW + 4 * lane;
            const float4 ix = *reinterpret_cast<const float4 *>(&sm.consts[0][so]);
            const float4 iy = *reinterpret_cast<const float4 *>(&sm.consts[1][so]);
            const float4 g = *reinterpret_cast<const float4 *>(&sm.consts[2][so]);
            const float4 rc = *reinterpret_cast<const float4 *>(&sm.consts[3][so]);
            float l11 = __shfl_up_sync(0xffffffffu, p11[r].w, 1);
            float l21 = __shfl_up_sync(0xffffffffu, p21[r].w, 1);
            if (left_edge || lane == 0) l11 = l21 = 0.f;  // p outside the image is 0 (lane 0: region edge, halo)
            if (top_edge[r]) up12 = up22 = zero4();
            float4 n1, n2;
            tvl1_primal_px(ix.x, iy.x, g.x, rc.x, u1[r].x, u2[r].x, (p11[r].x - l11) + (p12[r].x - up12.x), (p21[r].x - l21) + (p22[r].x - up22.x), c, n1.x, n2.x);
            tvl1_primal_px(ix.y, iy.y, g.y, rc.y, u1[r].y, u2[r].y, (p11[r].y - p11[r].x) + (p12[r].y - up12.y), (p21[r].y - p21[r].x) + (p22[r].y - up22.y), c, n1.y, n2.y);
            tvl1_primal_px(ix.z, iy.z, g.z, rc.z, u1[r].z, u2[r].z, (p11[r].z - p11[r].y) + (p12[r].z - up12.z), (p21[r].z - p21[r].y) + (p22[r].z - up22.z), c, n1.z, n2.z);
            tvl1_primal_px(ix.w, iy.w, g.w, rc.w, u1[r].w, u2[r].w, (p11[r].w - p11[r].z) + (p12[r].w - up12.w), (p21[r].w - p21[r].z) + (p22[r].w - up22.w), c, n1.w, n2.w);
            if (do_err) {
                // interior pixels inside the image only: every pixel is counted by exactly one tile
                const int ry = RPT * wq + r, gy = gy0 + r;
                const bool row_in = ry >= hy && ry < TH - hy && gy < H;
                const bool lane_in = 4 * lane >= hx && 4 * lane < TW - hx;
                if (row_in && lane_in) {
                    float d;
                    if (gx0 + 0 < W) { d = (u1[r].x - n1.x) * (u1[r].x - n1.x) + (u2[r].x - n2.x) * (u2[r].x - n2.x); err += (double)d; }
                    if (gx0 + 1 < W) { d = (u1[r].y - n1.y) * (u1[r].y - n1.y) + (u2[r].y - n2.y) * (u2[r].y - n2.y); err += (double)d; }
                    if (gx0 + 2 < W) { d = (u1[r].z - n1.z) * (u1[r].z - n1.z) + (u2[r].z - n2.z) * (u2[r].z - n2.z); err += (double)d; }
                    if (gx0 + 3 < W) { d = (u1[r].w - n1.w) * (u1[r].w - n1.w) + (u2[r].w - n2.w) * (u2[r].w - n2.w); err += (double)d; }
                }
            }
            u1[r] = n1;
            u2[r] = n2;
            up12 = p12[r];
            up22 = p22[r];
        }
        st4(&sm.u_top[0][wq][4 * lane], u1[0]);
        st4(&sm.u_top[1][wq][4 * lane], u2[0]);
        __syncthreads();
        // -------- dual: p <- (p + taut * grad u) / (1 + taut * |grad u|) ------------------------
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            float4 d1, d2;
            if (r < RPT - 1) {
                d1 = u1[r + 1];
                d2 = u2[r + 1];
            } else if (wq < kWarps - 1) {
                d1 = *reinterpret_cast<const float4 *>(&sm.u_top[0][wq + 1][4 * lane]);
                d2 = *reinterpret_cast<const float4 *>(&sm.u_top[1][wq + 1][4 * lane]);
            } else {
                d1 = u1[r];
                d2 = u2[r];
            }
            if (bot_edge[r]) {  // u(y+1) = u(y) on the last image row
                d1 = u1[r];
                d2 = u2[r];
            }
            const float r1 = __shfl_down_sync(0xffffffffu, u1[r].x, 1);
            const float r2 = __shfl_down_sync(0xffffffffu, u2[r].x, 1);
            // u(x+1) - u(x), zero on the last image column
            const float e1 = jlast == 0 ? 0.f : u1[r].y - u1[r].x, e2 = jlast == 0 ? 0.f : u2[r].y - u2[r].x;
            const float f1 = jlast == 1 ? 0.f : u1[r].z - u1[r].y, f2 = jlast == 1 ? 0.f : u2[r].z - u2[r].y;
            const float g1 = jlast == 2 ? 0.f : u1[r].w - u1[r].z, g2 = jlast == 2 ? 0.f : u2[r].w - u2[r].z;
            const float h1 = jlast == 3 ? 0.f : r1 - u1[r].w, h2 = jlast == 3 ? 0.f : r2 - u2[r].w;
            tvl1_dual_px(e1, d1.x - u1[r].x, e2, d2.x - u2[r].x, c.taut, p11[r].x, p12[r].x, p21[r].x, p22[r].x);
            tvl1_dual_px(f1, d1.y - u1[r].y, f2, d2.y - u2[r].y, c.taut, p11[r].y, p12[r].y, p21[r].y, p22[r].y);
            tvl1_dual_px(g1, d1.z - u1[r].z, g2, d2.z - u2[r].z, c.taut, p11[r].z, p12[r].z, p21[r].z, p22[r].z);
            tvl1_dual_px(h1, d1.w - u1[r].w, h2, d2.w - u2[r].w, c.taut, p11[r].w, p12[r].w, p21[r].w, p22[r].w);
        }
        st4(&sm.p_bot[0][wq][4 * lane], p12[RPT - 1]);
        st4(&sm.p_bot[1][wq][4 * lane], p22[RPT - 1]);
        __syncthreads();
    }

    // -------- write the interior to the other buffer ---------------------------------------------
    float *dst[6] = {L.u1[cur ^ 1], L.u2[cur ^ 1], job.p[cur ^ 1][0], job.p[cur ^ 1][1], job.p[cur ^ 1][2], job.p[cur ^ 1][3]};
    const bool lane_in = 4 * lane >= hx && 4 * lane < TW - hx && gx0 < W;
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
        const int ry = RPT * wq + r, gy = gy0 + r;
        if (lane_in && ry >= hy && ry < TH - hy && gy < H) {
            const size_t o = (size_t)gy * P + gx0;
            st4(dst[0] + o, u1[r]);
            st4(dst[1] + o, u2[r]);
            st4(dst[2] + o, p11[r]);
            st4(dst[3] + o, p12[r]);
            st4(dst[4] + o, p21[r]);
            st4(dst[5] + o, p22[r]);
        }
    }
    return err;
}

__device__ __forceinline__ double block_sum(double v, Smem &sm) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sm.red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < kWarps; ++k) s += sm.red[k];
    }
    __syncthreads();
    return s;  // valid in thread 0
}

__global__ void __launch_bounds__(kThreads, 1) k_tvl1_pair(const __grid_constant__ FusedJob job) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem &sm = *reinterpret_cast<Smem *>(smem_raw);
    unsigned epoch = 0;
    unsigned *bar = job.sync;
    int cur = 0;
    unsigned long long px_iters = 0;

    for (int s = job.nscales - 1; s >= 0; --s) {
        const FusedLevel &L = job.lv[s];
        cur = 0;  // level start: u[0] holds the upsampled (or zero) flow, p[0] is zeroed
        phase_level_start(job, L, s == job.nscales - 1);
        grid_barrier(bar, epoch);
        const double scaled_eps = job.epsilon * job.epsilon * (double)((long long)L.w * L.h);  // A.4
        for (int wi = 0; wi < job.warps; ++wi) {
            phase_warp(job, L, cur);
            grid_barrier(bar, epoch);
            double error = DBL_MAX, prev_error = 0.0;
            int n = 0;
            while (error > scaled_eps && n < job.iterations) {
                // plan the epoch: iterations up to and including the next convergence check (A.4)
                int K = 0, nn = n;
                bool check = false;
                double pp = prev_error;
                for (;;) {
                    const bool calc = job.epsilon > 0 && (nn & 1) && pp < scaled_eps;
                    ++K;
                    ++nn;
                    if (calc) {
                        check = true;
                        break;
                    }
                    pp -= scaled_eps;
                    if (nn >= job.iterations) break;
                }
                int remaining = K;
                double cta_err = 0.0;
                while (remaining > 0) {
                    const int nch = (remaining + job.k - 1) / job.k;
                    const int kk = (remaining + nch - 1) / nch;
                    const bool chk = check && kk == remaining;
                    const int hx = (kk + 3) & ~3, hy = kk;
                    const int iw = TW - 2 * hx, ih = TH - 2 * hy;
                    const int ntx = (L.w + iw - 1) / iw, nty = (L.h + ih - 1) / ih;
                    const int ntiles = ntx * nty;
                    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
                        const int ty = t / ntx, tx = t - ty * ntx;
                        const double e = process_tile(job, L, cur, tx * iw - hx, ty * ih - hy, kk, hx, hy, chk, sm);
                        if (chk) {
                            const double bs = block_sum(e, sm);
                            if (threadIdx.x == 0) cta_err += bs;
                        }
                    }
                    if (chk && threadIdx.x == 0) job.partials[blockIdx.x] = cta_err;
                    grid_barrier(bar, epoch);
                    cur ^= 1;
                    remaining -= kk;
                }
                n = nn;
                if (check) {
                    // every CTA sums all partials in the same fixed order -> identical decisions everywhere
                    if (threadIdx.x < 32) {
                        double v = 0.0;
                        for (int i = threadIdx.x; i < (int)gridDim.x; i += 32) v += __ldcg(job.partials + i);
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                        if (threadIdx.x == 0) sm.bcast[0] = v;
                    }
                    __syncthreads();
                    error = sm.bcast[0];
                    __syncthreads();
                    prev_error = error;
                } else {
                    error = DBL_MAX;
                    prev_error = pp;
                }
            }
            if (blockIdx.x == 0 && threadIdx.x == 0) job.ctl->iters[s * job.warps + wi] = n;
            px_iters += (unsigned long long)n * (unsigned long long)(L.w * L.h);
        }
        if (s > 0) {
            phase_upsample(job, L, job.lv[s - 1], cur);
            grid_barrier(bar, epoch);
        }
    }
    phase_merge(job, job.lv[0], cur);
    if (blockIdx.x == 0 && threadIdx.x == 0) job.ctl->px_iters_total += px_iters;  // single writer, launches are serialised

    // last CTA out resets the barrier words for the next launch
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned done = atomicAdd(bar + 1, 1u);
        if (done == gridDim.x - 1) {
            bar[0] = 0;
            bar[1] = 0;
            __threadfence();
        }
    }
}

}  // namespace

int launch_tvl1_fused(const FusedJob &job, int device, cudaStream_t s) {
    static int num_sms[64] = {};
    static bool configured = false;
    if (!configured) {
        DFB_CUDA(cudaFuncSetAttribute(k_tvl1_pair, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
        configured = true;
    }
    if (device < 64 && num_sms[device] == 0)
        DFB_CUDA(cudaDeviceGetAttribute(&num_sms[device], cudaDevAttrMultiProcessorCount, device));
    const int sms = device < 64 ? num_sms[device] : 148;
    // enough CTAs for the busiest phase, never more than are co-resident (1 CTA / SM)
    const FusedLevel &L0 = job.lv[0];
    const int iw = TW - 8, ih = TH - 4;  // smallest halo => most tiles
    const int max_tiles = ceil_div(L0.w, iw) * ceil_div(L0.h, ih);
    const int grid = std::max(1, std::min(sms, max_tiles));
    void *args[] = {const_cast<FusedJob *>(&job)};
    DFB_CUDA(cudaLaunchCooperativeKernel((const void *)k_tvl1_pair, dim3(grid), dim3(kThreads), args, sizeof(Smem), s));
    return 1;
}

}  // namespace dfb
