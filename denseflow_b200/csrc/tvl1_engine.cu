// tvl1_engine.cu — host orchestration of the TV-L1 path: the replacement for
// cv::cuda::OpticalFlowDual_TVL1 as created/called at /root/reference/src/denseflow_gpu.cpp:299,327
// (control flow per SURVEY.md Appendix A.1, A.2, A.4, A.5).
//
// Two schedules over the same arithmetic:
//   fused = 1 (default)  one persistent cooperative kernel per pair (tvl1_fused.cu): gradients,
//                        warps, the primal/dual iterations with k steps kept on chip per tile, the
//                        A.4 convergence state machine and the flow upsampling all run on-device
//                        with no host round trip.
//   fused = 0            one kernel per half-step, the reference's launch structure, with the A.4
//                        state machine on the host (one stream sync per convergence check).
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "engine.h"
#include "tvl1.cuh"
#include "tvl1_fused.cuh"

namespace dfb {

namespace {

constexpr int kMaxScales = 16;
constexpr double kAllocScaleStep = 0.8;  // scaleStep the slot / lane pyramid layout is sized for (the create() default)

struct Tvl1Params {
    double tau = 0.25, lambda = 0.15, theta = 0.3;
    int nscales = 5, warps = 5;
    double epsilon = 0.01;
    int iterations = 300;
    double scale_step = 0.8;
    int fused = 1;
    int fused_k = kFusedDefaultK;
    int flag_sync = 1;
    int time_kernels = 0;
    int use_tma = 1;
    int prefetch = 1;
    int lanes = 0;  // pairs solved side by side per fused launch; 0 = choose from the tile counts
    int serial_launches = 0;  // 1: fused launches of all handles on one device execute strictly one after another (event chain)
};

class Tvl1 final : public FlowAlgorithm {
  public:
    Tvl1(int device, int max_w, int max_h) : device_(device), max_w_(max_w), max_h_(max_h) {
        DFB_CUDA(cudaSetDevice(device_));
        if (const char *e = std::getenv("DFB_TVL1_LANES")) prm_.lanes = std::max(0, std::min(atoi(e), kFusedMaxLanes));  // debugging aid
        if (const char *e = std::getenv("DFB_TVL1_SERIAL_LAUNCHES")) prm_.serial_launches = atoi(e) != 0;         // debugging aid
        allocate();
    }
    ~Tvl1() override {
        cudaSetDevice(device_);
        for (auto p : extra_slots_) cudaFree(p);
        for (auto &l : lanes_) {
            if (l.host_ctl) cudaFreeHost(l.host_ctl);
            if (l.own) cudaFree(l.own);
            if (l.tmaps) cudaFree(l.tmaps);
        }
        if (stats_event_) cudaEventDestroy(stats_event_);
        if (pair_log_) cudaFreeHost(pair_log_);
        if (unfused_scratch_) cudaFree(unfused_scratch_);
        for (int r = 0; r < kJobRing; ++r) {
            if (job_ring_h_[r]) cudaFreeHost(job_ring_h_[r]);
            if (job_ring_d_[r]) cudaFree(job_ring_d_[r]);
            if (job_ring_ev_[r]) cudaEventDestroy(job_ring_ev_[r]);
        }
        for (auto &e : timing_ev_)
            for (auto ev : e)
                if (ev) cudaEventDestroy(ev);
    }
    const char *name() const override { return "tvl1"; }
    int num_slots() const override { return (int)slots_.size(); }

    void ensure_slots(int n) override {
        while ((int)slots_.size() < n) {
            float *p = nullptr;
            DFB_CUDA(cudaMalloc(&p, pyr_elems_ * sizeof(float)));
            // planes start finite: padding columns are read (never used).  The memset runs on the null stream, which the
            // handle's non-blocking streams do not wait for: finish it before any of them can touch the slot.
            DFB_CUDA(cudaMemset(p, 0, pyr_elems_ * sizeof(float)));
            DFB_CUDA(cudaDeviceSynchronize());
            extra_slots_.push_back(p);
            slots_.push_back(p);
        }
    }

    bool set_param(const std::string &k, double v) override {
        if (k == "tau") prm_.tau = v;
        else if (k == "lambda") prm_.lambda = v;
        else if (k == "theta") prm_.theta = v;
        else if (k == "nscales") { if (v < 1 || v > kMaxScales) return false; prm_.nscales = (int)v; }
        else if (k == "warps") { if (v < 1 || v > 16) return false; prm_.warps = (int)v; }
        else if (k == "epsilon") prm_.epsilon = v;
        else if (k == "iterations") { if (v < 1) return false; prm_.iterations = (int)v; }
        else if (k == "scale_step") {
            // the pyramid slots were laid out at create time for the default 0.8: a larger factor would not fit them
            if (!(v > 0 && v <= kAllocScaleStep)) return false;
            prm_.scale_step = v;
        }
        else if (k == "fused") prm_.fused = v != 0;
        else if (k == "fused_k") { if (v < 1 || v > kFusedMaxK || 2 * v >= kFusedTileH) return false; prm_.fused_k = (int)v; }
        else if (k == "flag_sync") prm_.flag_sync = (int)v;  // 0 CTA barriers, 1 spin on neighbour flags, n > 1: spin with n ns back-off
        else if (k == "time_kernels") prm_.time_kernels = v != 0;
        else if (k == "use_tma") prm_.use_tma = v != 0;
        else if (k == "prefetch") prm_.prefetch = v != 0;
        else if (k == "serial_launches") prm_.serial_launches = v != 0;
        else if (k == "lanes") { if (v < 0 || v > kFusedMaxLanes) return false; prm_.lanes = (int)v; }
        else return false;
        return true;
    }
    bool get_param(const std::string &k, double *v) const override {
        if (k == "tau") *v = prm_.tau;
        else if (k == "lambda") *v = prm_.lambda;
        else if (k == "theta") *v = prm_.theta;
        else if (k == "nscales") *v = prm_.nscales;
        else if (k == "warps") *v = prm_.warps;
        else if (k == "epsilon") *v = prm_.epsilon;
        else if (k == "iterations") *v = prm_.iterations;
        else if (k == "scale_step") *v = prm_.scale_step;
        else if (k == "fused") *v = prm_.fused;
        else if (k == "fused_k") *v = prm_.fused_k;
        else if (k == "flag_sync") *v = prm_.flag_sync;
        else if (k == "time_kernels") *v = prm_.time_kernels;
        else if (k == "use_tma") *v = prm_.use_tma;
        else if (k == "prefetch") *v = prm_.prefetch;
        else if (k == "serial_launches") *v = prm_.serial_launches;
        else if (k == "lanes") *v = prm_.lanes;
        else return false;
        return true;
    }

    // A.1: level sizes = round-half-even(dim * scaleStep); a level with cols<16 || rows<16 is dropped.
    int level_geometry(int w, int h, LevelGeom *lv) const {
        int n = 1;
        lv[0] = {w, h, round_up(w, 32)};
        for (int s = 1; s < prm_.nscales; ++s) {
            const int nw = (int)std::nearbyint((double)lv[s - 1].w * prm_.scale_step);
            const int nh = (int)std::nearbyint((double)lv[s - 1].h * prm_.scale_step);
            if (nw < 16 || nh < 16) break;
            lv[s] = {nw, nh, round_up(nw, 32)};
            n = s + 1;
        }
        return n;
    }

    void prepare_frame(const uint8_t *src, size_t pitch_bytes, int w, int h, int slot, cudaStream_t s) override {
        LevelGeom lv[kMaxScales];
        const int n = level_geometry(w, h, lv);
        float *base = slots_.at(slot);
        launch_u8_to_f32(src, pitch_bytes, level_plane(base, lv, 0), s);
        ++launches;
        const float finv = (float)(1.0 / prm_.scale_step);  // dsize empty => fx = float(1/scaleStep)
        for (int l = 1; l < n; ++l) {
            launch_resize_linear(level_plane(base, lv, l - 1), level_plane(base, lv, l), finv, finv, 1.0f, s);
            ++launches;
        }
    }

    void prepare_frames(int n, const uint8_t *const *srcs, size_t pitch_bytes, int w, int h, const int *slots, cudaStream_t s) override {
        LevelGeom lv[kMaxScales];
        const int nl = level_geometry(w, h, lv);
        const float finv = (float)(1.0 / prm_.scale_step);
        for (int f0 = 0; f0 < n; f0 += kMaxFrameBatch) {
            const int nf = std::min(kMaxFrameBatch, n - f0);
            FramePtrs fp{};
            for (int i = 0; i < nf; ++i) {
                fp.src[i] = srcs[f0 + i];
                fp.base[i] = slots_.at(slots[f0 + i]);
            }
            launch_u8_to_f32_batch(fp, nf, pitch_bytes, w, h, lv[0].pitch, s);
            ++launches;
            size_t off = 0;
            for (int l = 1; l < nl; ++l) {
                const size_t next = off + level_stride(l - 1);
                launch_resize_linear_batch(fp, nf, off, lv[l - 1].w, lv[l - 1].h, lv[l - 1].pitch, next, lv[l].w, lv[l].h, lv[l].pitch, finv, finv, s);
                ++launches;
                off = next;
            }
        }
    }

    void solve(int slot_a, int slot_b, int w, int h, float *flow_xy, size_t flow_pitch_bytes,
               cudaStream_t s) override {
        LevelGeom lv[kMaxScales];
        const int n = level_geometry(w, h, lv);
        last_nscales_ = n;
        std::memcpy(last_lv_, lv, sizeof(lv));
        if (prm_.fused) {
            PairJob one{};
            one.slot_a = slot_a;
            one.slot_b = slot_b;
            one.flow_xy = flow_xy;
            one.flow_pitch_bytes = flow_pitch_bytes;
            solve_fused(&one, 1, lv, n, s);
        } else {
            solve_unfused(slot_a, slot_b, lv, n, flow_xy, flow_pitch_bytes, s);
        }
    }

    // pairs per fused launch.  Utilisation model: a lane of G = SMs / B CTAs visits tiles(level) tiles
    // in ceil(tiles / G) rounds; weight the levels by where the iterations are (the coarsest scale runs
    // most of them), pick the B with the fewest idle CTA-rounds.
    int max_concurrent_pairs(int w, int h) override {
        if (!prm_.fused) return 1;
        if (prm_.lanes > 0) return prm_.lanes;
        LevelGeom lv[kMaxScales];
        const int n = level_geometry(w, h, lv);
        const int sms = fused_cta_slots(device_);
        const int hy = prm_.fused_k, hx = (hy + 3) & ~3;
        // lanes are sized for the handle's MAXIMUM frame: keep their total workspace under 8 GB
        const size_t lane_bytes = 4 * Slab::padded(pyr_elems_, 4) + 14 * Slab::padded(plane_elems_, 4);
        const int mem_cap = (int)std::max<size_t>(1, (size_t(8) << 30) / std::max<size_t>(lane_bytes, 1));
        const int maxB = std::min(kFusedMaxLanes, mem_cap);
        double util[kFusedMaxLanes + 1] = {};
        double best_u = 0;
        for (int B = 1; B <= maxB; ++B) {
            const int G = sms / B;
            if (G < 1) break;
            double num = 0, den = 0;
            for (int l = 0; l < n; ++l) {
                const double wgt = l == n - 1 ? 7.0 : (l == n - 2 ? 1.5 : 1.0);
                const int tiles = fused_tiles_along(lv[l].w, kFusedTileW, hx) * fused_tiles_along(lv[l].h, kFusedTileH, hy);
                const int rounds = (tiles + G - 1) / G;
                num += wgt * tiles;
                den += wgt * (double)rounds * G;
            }
            util[B] = num / den * (double)(G * B) / sms;
            best_u = std::max(best_u, util[B]);
        }
        // the smallest lane count within 5 % of the best modelled utilisation: more lanes than needed cost workspace, L2 hits
        // (7.9 -> 10.8 GB of DRAM traffic per 1080p pair from 7 to 16 lanes) and a longer un-overlapped download of the last group
        int best = 1;
        for (int B = 1; B <= maxB; ++B)
            if (util[B] >= 0.95 * best_u) {
                best = B;
                break;
            }
        return best;
    }

    void solve_batch(const PairJob *jobs, int count, int w, int h, cudaStream_t s) override {
        LevelGeom lv[kMaxScales];
        const int n = level_geometry(w, h, lv);
        last_nscales_ = n;
        std::memcpy(last_lv_, lv, sizeof(lv));
        if (!prm_.fused) {
            for (int i = 0; i < count; ++i) {
                float *flow = jobs[i].flow_xy;
                size_t pitch = jobs[i].flow_pitch_bytes;
                if (!flow) {  // quantised output only: the float2 field lives in an engine-owned scratch plane
                    if (!unfused_scratch_) DFB_CUDA(cudaMalloc(&unfused_scratch_, (size_t)max_w_ * max_h_ * 2 * sizeof(float)));
                    flow = unfused_scratch_;
                    pitch = (size_t)w * 2 * sizeof(float);
                }
                solve_unfused(jobs[i].slot_a, jobs[i].slot_b, lv, n, flow, pitch, s);
                if (jobs[i].bound > 0) {  // the reference's structure: a separate pass over the finished field
                    launch_quantise(flow, pitch, lv[0].w, lv[0].h, jobs[i].bound, jobs[i].qx, jobs[i].qy, jobs[i].q_pitch, s);
                    ++launches;
                }
            }
            return;
        }
        const int B = std::min(max_concurrent_pairs(w, h), kFusedMaxLanes);
        for (int i = 0; i < count; i += B) solve_fused(jobs + i, std::min(B, count - i), lv, n, s);
    }

    void phase_ns(uint64_t *out) override {
        if (stats_pending_) DFB_CUDA(cudaEventSynchronize(stats_event_));
        for (int i = 0; i < 32; ++i) out[i] = prm_.fused ? lanes_[last_lane_].host_ctl->prof[i] : 0;
    }
    void tvl1_stats(dfb_tvl1_stats *out) override {
        *out = dfb_tvl1_stats{};
        if (stats_pending_) {  // fused engine: the log was written by the kernel into mapped host memory
            DFB_CUDA(cudaEventSynchronize(stats_event_));
            std::memcpy(last_iters_, pair_log_ + (size_t)last_log_slot_ * kLogInts, sizeof(last_iters_));
            stats_pending_ = false;
        }
        pixel_iters = unfused_px_iters_;
        pixel_chunks = 0;
        for (auto &l : lanes_) {
            pixel_iters += l.host_ctl->px_iters_total;
            pixel_chunks += l.host_ctl->px_chunks_total;
        }
        out->nscales = last_nscales_;
        out->warps = prm_.warps;
        for (int s = 0; s < last_nscales_; ++s) {
            out->level_w[s] = last_lv_[s].w;
            out->level_h[s] = last_lv_[s].h;
        }
        std::memcpy(out->iters, last_iters_, sizeof(last_iters_));
    }

    void begin_batch() override { pair_log_count_ = 0; }
    std::string fault_info() override {
        for (size_t i = 0; i < lanes_.size(); ++i) {
            const volatile unsigned *st = lanes_[i].host_ctl->stall;
            if (st[0])
                return "k_tvl1_pair grid-barrier watchdog: CTA " + std::to_string(st[1]) + " (lane " + std::to_string(st[5]) + ", " +
                       std::to_string(st[4]) + " CTAs per lane) waited for arrival count " + std::to_string(st[2]) + ", saw " + std::to_string(st[3]);
        }
        return {};
    }
    bool pair_stats(int idx, dfb_tvl1_stats *out) override {
        *out = dfb_tvl1_stats{};
        if (idx < 0 || idx >= pair_log_count_ || idx < pair_log_count_ - kPairLog) return false;
        if (stats_pending_) DFB_CUDA(cudaEventSynchronize(stats_event_));
        out->nscales = last_nscales_;
        out->warps = prm_.warps;
        for (int s = 0; s < last_nscales_; ++s) {
            out->level_w[s] = last_lv_[s].w;
            out->level_h[s] = last_lv_[s].h;
        }
        std::memcpy(out->iters, pair_log_ + (size_t)(idx % kPairLog) * kLogInts, sizeof(out->iters));
        return true;
    }

  private:
    int next_log_slot() {
        last_log_slot_ = pair_log_count_ % kPairLog;
        ++pair_log_count_;
        return last_log_slot_;
    }
    Plane level_plane(float *base, const LevelGeom *lv, int l) const {
        size_t off = 0;
        for (int i = 0; i < l; ++i) off += level_stride(i);
        return Plane{base + off, lv[l].w, lv[l].h, lv[l].pitch};
    }
    // slot layout is fixed by the MAX geometry so any smaller frame fits
    size_t level_stride(int l) const { return max_lv_elems_[l]; }

    void allocate() {
        // geometry of the largest frame, with the full default pyramid depth available
        Tvl1Params keep = prm_;
        prm_.nscales = kMaxScales;
        prm_.scale_step = kAllocScaleStep;
        LevelGeom lv[kMaxScales];
        const int n = level_geometry(max_w_, max_h_, lv);
        prm_ = keep;
        pyr_elems_ = 0;
        for (int l = 0; l < kMaxScales; ++l) {
            // +pitch: one spare row so float4 tile reads one row past the image stay inside the slab
            max_lv_elems_[l] = l < n ? (size_t)lv[l].pitch * (lv[l].h + 1) : 0;
            max_lv_elems_[l] = (max_lv_elems_[l] + 63) & ~size_t(63);
            pyr_elems_ += max_lv_elems_[l];
        }
        plane_elems_ = (size_t)lv[0].pitch * (lv[0].h + 1);
        constexpr int kInitialSlots = 4;
        slab_.reserve(kInitialSlots * Slab::padded(pyr_elems_, 4) + (1 << 12));
        for (int i = 0; i < kInitialSlots; ++i) slots_.push_back(slab_.take<float>(pyr_elems_));
        ensure_lanes(1);
        DFB_CUDA(cudaEventCreateWithFlags(&stats_event_, cudaEventDisableTiming));
        DFB_CUDA(cudaHostAlloc(&pair_log_, sizeof(int) * kLogInts * kPairLog, cudaHostAllocMapped));
        std::memset(pair_log_, 0, sizeof(int) * kLogInts * kPairLog);
        DFB_CUDA(cudaHostGetDevicePointer(&pair_log_dev_, pair_log_, 0));
        // every plane starts finite: padding columns are read (never used) by vectorised kernels
        slab_.zero();
    }

    // per-pair workspace: flow pyramids (ping-pong), gradient / warp planes, dual variables (ping-pong),
    // convergence partials, barrier words and the mapped control block
    struct Lane {
        void *own = nullptr;
        float *u1pyr[2] = {}, *u2pyr[2] = {};
        float *I1x = nullptr, *I1y = nullptr, *I1wx = nullptr, *I1wy = nullptr, *grad = nullptr, *rho_c = nullptr;
        float *p[2][4] = {};
        double *partials = nullptr;
        unsigned *sync = nullptr;
        FusedHostCtl *host_ctl = nullptr, *dev_ctl = nullptr;
        void *tmaps = nullptr;  // device array CUtensorMap[kMaxScales][kFusedMapsPerLevel]
        int tmap_w = 0, tmap_h = 0, tmap_n = 0;
        double tmap_ss = 0;
    };

    void ensure_lanes(int n) {
        while ((int)lanes_.size() < n) {
            Lane l;
            const size_t pyr = Slab::padded(pyr_elems_, 4), pl = Slab::padded(plane_elems_, 4);
            const size_t bytes = 4 * pyr + 14 * pl + Slab::padded(kMaxPartials, 8) + 256;
            DFB_CUDA(cudaMalloc(&l.own, bytes));
            DFB_CUDA(cudaMemset(l.own, 0, bytes));  // planes start finite: padding is read (never used) by vector loads
            DFB_CUDA(cudaDeviceSynchronize());      // null-stream memset vs the caller's (possibly non-blocking) stream
            char *c = static_cast<char *>(l.own);
            auto take = [&](size_t b) { char *r = c; c += b; return r; };
            for (int b = 0; b < 2; ++b) {
                l.u1pyr[b] = reinterpret_cast<float *>(take(pyr));
                l.u2pyr[b] = reinterpret_cast<float *>(take(pyr));
            }
            l.I1x = reinterpret_cast<float *>(take(pl));
            l.I1y = reinterpret_cast<float *>(take(pl));
            l.I1wx = reinterpret_cast<float *>(take(pl));
            l.I1wy = reinterpret_cast<float *>(take(pl));
            l.grad = reinterpret_cast<float *>(take(pl));
            l.rho_c = reinterpret_cast<float *>(take(pl));
            for (int b = 0; b < 2; ++b)
                for (int k = 0; k < 4; ++k) l.p[b][k] = reinterpret_cast<float *>(take(pl));
            l.partials = reinterpret_cast<double *>(take(Slab::padded(kMaxPartials, 8)));
            l.sync = reinterpret_cast<unsigned *>(take(256));
            DFB_CUDA(cudaHostAlloc(&l.host_ctl, sizeof(FusedHostCtl), cudaHostAllocMapped));
            std::memset(l.host_ctl, 0, sizeof(FusedHostCtl));
            DFB_CUDA(cudaHostGetDevicePointer(&l.dev_ctl, l.host_ctl, 0));
            DFB_CUDA(cudaMalloc(&l.tmaps, (size_t)kMaxScales * kFusedMapsPerLevel * kTensorMapBytes));
            lanes_.push_back(l);
        }
    }

    Plane work(float *p, const LevelGeom &g) const { return Plane{p, g.w, g.h, g.pitch}; }

    // ---- fused = 0: reference launch structure, A.4 state machine on the host ------------------
    void solve_unfused(int slot_a, int slot_b, const LevelGeom *lv, int n, float *flow_xy, size_t flow_pitch_bytes,
                       cudaStream_t s) {
        const Tvl1Consts c{(float)(prm_.lambda * prm_.theta), (float)(prm_.tau / prm_.theta), (float)prm_.theta};
        std::memset(last_iters_, 0, sizeof(last_iters_));
        Lane &wk = lanes_[0];
        float *I0b = slots_.at(slot_a), *I1b = slots_.at(slot_b);
        for (int l = n - 1; l >= 0; --l) {
            const LevelGeom &g = lv[l];
            const Plane I0 = level_plane(I0b, lv, l), I1 = level_plane(I1b, lv, l);
            const Plane u1 = level_plane(wk.u1pyr[0], lv, l), u2 = level_plane(wk.u2pyr[0], lv, l);
            const Plane I1x = work(wk.I1x, g), I1y = work(wk.I1y, g), I1wx = work(wk.I1wx, g), I1wy = work(wk.I1wy, g);
            const Plane grad = work(wk.grad, g), rho_c = work(wk.rho_c, g);
            const Plane p11 = work(wk.p[0][0], g), p12 = work(wk.p[0][1], g), p21 = work(wk.p[0][2], g), p22 = work(wk.p[0][3], g);
            if (l == n - 1) {  // useInitialFlow = false
                launch_fill(u1, 0.f, s);
                launch_fill(u2, 0.f, s);
                launches += 2;
            }
            launch_centered_gradient(I1, I1x, I1y, s);
            launch_fill(p11, 0.f, s);  // once per scale, not per warp (A.2 step 2)
            launch_fill(p12, 0.f, s);
            launch_fill(p21, 0.f, s);
            launch_fill(p22, 0.f, s);
            launches += 5;
            const double scaled_eps = prm_.epsilon * prm_.epsilon * (double)((long)g.w * g.h);
            const int nblk = estimate_u_blocks(g.w, g.h);
            for (int wi = 0; wi < prm_.warps; ++wi) {
                launch_warp_backward(I0, I1, I1x, I1y, u1, u2, I1wx, I1wy, grad, rho_c, s);
                ++launches;
                double error = DBL_MAX, prev_error = 0.0;
                int it = 0;
                for (; error > scaled_eps && it < prm_.iterations; ++it) {
                    const bool calc_error = prm_.epsilon > 0 && (it & 1) && prev_error < scaled_eps;
                    launch_estimate_u(I1wx, I1wy, grad, rho_c, p11, p12, p21, p22, u1, u2, c,
                                      calc_error ? wk.partials : nullptr, s);
                    ++launches;
                    if (calc_error) {
                        launch_sum_partials(wk.partials, nblk, &wk.dev_ctl->error, s);
                        ++launches;
                        DFB_CUDA(cudaStreamSynchronize(s));  // the reference syncs here too (stream.waitForCompletion)
                        error = wk.host_ctl->error;
                        prev_error = error;
                    } else {
                        error = DBL_MAX;
                        prev_error -= scaled_eps;
                    }
                    launch_estimate_dual(u1, u2, p11, p12, p21, p22, c, s);
                    ++launches;
                }
                last_iters_[l * prm_.warps + wi] = it;
            }
            if (l > 0) {  // A.2 step 4: explicit dsize => f = float(1 / (dst/src)); then * float(1/scaleStep)
                const float ufx = (float)(1.0 / ((double)lv[l - 1].w / (double)g.w));
                const float ufy = (float)(1.0 / ((double)lv[l - 1].h / (double)g.h));
                const float mul = (float)(1.0 / prm_.scale_step);
                launch_resize_linear(u1, level_plane(wk.u1pyr[0], lv, l - 1), ufx, ufy, mul, s);
                launch_resize_linear(u2, level_plane(wk.u2pyr[0], lv, l - 1), ufx, ufy, mul, s);
                launches += 2;
            }
        }
        launch_merge_flow(level_plane(wk.u1pyr[0], lv, 0), level_plane(wk.u2pyr[0], lv, 0), flow_xy, flow_pitch_bytes, s);
        ++launches;
        stats_pending_ = false;
        std::memcpy(pair_log_ + (size_t)next_log_slot() * kLogInts, last_iters_, sizeof(last_iters_));
        accumulate_pixel_iters();
    }

    // TMA descriptors of the lane's shared-memory-resident planes, rebuilt only when the frame geometry changes
    void ensure_tensor_maps(Lane &wk, const LevelGeom *lv, int n) {
        if (wk.tmap_w == lv[0].w && wk.tmap_h == lv[0].h && wk.tmap_n == n && wk.tmap_ss == prm_.scale_step) return;
        std::vector<char> host((size_t)kMaxScales * kFusedMapsPerLevel * kTensorMapBytes, 0);
        for (int l = 0; l < n; ++l) {
            char *m = host.data() + (size_t)l * kFusedMapsPerLevel * kTensorMapBytes;
            float *planes[kFusedMapsPerLevel] = {wk.I1wx, wk.I1wy, wk.grad, wk.rho_c, level_plane(wk.u1pyr[0], lv, l).p,
                                                 level_plane(wk.u2pyr[0], lv, l).p, level_plane(wk.u1pyr[1], lv, l).p,
                                                 level_plane(wk.u2pyr[1], lv, l).p};
            for (int k = 0; k < kFusedMapsPerLevel; ++k)
                fused_encode_tensor_map(m + (size_t)k * kTensorMapBytes, planes[k], lv[l].w, lv[l].h, lv[l].pitch);
        }
        DFB_CUDA(cudaDeviceSynchronize());  // the previous maps may still be in use by a running launch
        DFB_CUDA(cudaMemcpy(wk.tmaps, host.data(), host.size(), cudaMemcpyHostToDevice));
        wk.tmap_w = lv[0].w;
        wk.tmap_h = lv[0].h;
        wk.tmap_n = n;
        wk.tmap_ss = prm_.scale_step;
    }

    // ---- fused = 1 -------------------------------------------------------------------------------
    void solve_fused(const PairJob *jobs, int count, const LevelGeom *lv, int n, cudaStream_t s) {
        ensure_lanes(count);
        FusedBatch batch{};
        batch.njobs = count;
        batch.group = std::max(1, fused_cta_slots(device_) / count);
        // job descriptions: filled in a pinned host slot of a small ring, copied to its device twin on the launch stream
        if (!job_ring_h_[0]) {
            for (int r = 0; r < kJobRing; ++r) {
                DFB_CUDA(cudaHostAlloc(&job_ring_h_[r], sizeof(FusedJob) * kFusedMaxLanes, cudaHostAllocDefault));
                DFB_CUDA(cudaMalloc(&job_ring_d_[r], sizeof(FusedJob) * kFusedMaxLanes));
                DFB_CUDA(cudaEventCreateWithFlags(&job_ring_ev_[r], cudaEventDisableTiming));
            }
        }
        const int jr = job_ring_next_++ % kJobRing;
        if (job_ring_next_ > kJobRing) DFB_CUDA(cudaEventSynchronize(job_ring_ev_[jr]));  // the copy out of this host slot has run (no-op if never recorded)
        FusedJob *host_jobs = job_ring_h_[jr];
        for (int i = 0; i < count; ++i) {
            Lane &wk = lanes_[i];
            FusedJob &job = host_jobs[i];
            job = FusedJob{};
            job.nscales = n;
            job.warps = prm_.warps;
            job.iterations = prm_.iterations;
            job.epsilon = prm_.epsilon;
            job.k = prm_.fused_k;
            job.flag_sync = prm_.flag_sync;
            job.use_tma = prm_.use_tma;
            job.prefetch = prm_.prefetch;
            if (prm_.use_tma) ensure_tensor_maps(wk, lv, n);
            job.tmaps = wk.tmaps;
            job.c = Tvl1Consts{(float)(prm_.lambda * prm_.theta), (float)(prm_.tau / prm_.theta), (float)prm_.theta};
            job.up_mul = (float)(1.0 / prm_.scale_step);
            for (int l = 0; l < n; ++l) {
                FusedLevel &L = job.lv[l];
                L.w = lv[l].w;
                L.h = lv[l].h;
                L.pitch = lv[l].pitch;
                L.I0 = level_plane(slots_.at(jobs[i].slot_a), lv, l).p;
                L.I1 = level_plane(slots_.at(jobs[i].slot_b), lv, l).p;
                for (int b = 0; b < 2; ++b) {
                    L.u1[b] = level_plane(wk.u1pyr[b], lv, l).p;
                    L.u2[b] = level_plane(wk.u2pyr[b], lv, l).p;
                }
                if (l > 0) {
                    L.up_fx = (float)(1.0 / ((double)lv[l - 1].w / (double)lv[l].w));
                    L.up_fy = (float)(1.0 / ((double)lv[l - 1].h / (double)lv[l].h));
                }
            }
            job.I1x = wk.I1x;
            job.I1y = wk.I1y;
            job.I1wx = wk.I1wx;
            job.I1wy = wk.I1wy;
            job.grad = wk.grad;
            job.rho_c = wk.rho_c;
            for (int b = 0; b < 2; ++b)
                for (int k = 0; k < 4; ++k) job.p[b][k] = wk.p[b][k];
            job.partials = wk.partials;
            job.sync = wk.sync;
            job.ctl = wk.dev_ctl;
            job.iters_log = pair_log_dev_ + (size_t)next_log_slot() * kLogInts;
            job.flow_xy = jobs[i].flow_xy;
            job.flow_pitch_bytes = jobs[i].flow_pitch_bytes;
            job.bound = jobs[i].bound;
            job.qx = jobs[i].qx;
            job.qy = jobs[i].qy;
            job.q_pitch = jobs[i].q_pitch;
        }
        // one CTA per SM at most: the tile counts of the largest level bound the useful group size
        const int hx = 4, hy = 1;
        const int max_tiles = fused_tiles_along(lv[0].w, kFusedTileW, hx) * fused_tiles_along(lv[0].h, kFusedTileH, hy);
        batch.group = std::max(1, std::min(batch.group, max_tiles));
        if (prm_.time_kernels) {
            if (timing_used_ == kTimingRing) drain_timing();
            if (!timing_ev_[0][0])
                for (int i = 0; i < kTimingRing; ++i) {
                    DFB_CUDA(cudaEventCreate(&timing_ev_[i][0]));
                    DFB_CUDA(cudaEventCreate(&timing_ev_[i][1]));
                }
            DFB_CUDA(cudaEventRecord(timing_ev_[timing_used_][0], s));
        }
        if (count > kFusedParamLanes) {  // too many for the parameter bank: through the device-memory ring
            DFB_CUDA(cudaMemcpyAsync(job_ring_d_[jr], host_jobs, sizeof(FusedJob) * count, cudaMemcpyHostToDevice, s));
            DFB_CUDA(cudaEventRecord(job_ring_ev_[jr], s));
        }
        batch.jobs = job_ring_d_[jr];
        batch.host_jobs = host_jobs;
        launches += launch_tvl1_fused(batch, device_, s, prm_.serial_launches != 0);
        if (prm_.time_kernels) {
            DFB_CUDA(cudaEventRecord(timing_ev_[timing_used_][1], s));
            ++timing_used_;
            timed_pairs_ += count;
        }
        last_lane_ = count - 1;
        DFB_CUDA(cudaEventRecord(stats_event_, s));
        stats_pending_ = true;
    }

    // CUDA-event timing of the dominant kernel (events on the launching stream, drained lazily)
    void drain_timing() {
        for (int i = 0; i < timing_used_; ++i) {
            DFB_CUDA(cudaEventSynchronize(timing_ev_[i][1]));
            float ms = 0.f;
            DFB_CUDA(cudaEventElapsedTime(&ms, timing_ev_[i][0], timing_ev_[i][1]));
            timed_ns_ += (uint64_t)((double)ms * 1e6);
            ++timed_launches_;
        }
        timing_used_ = 0;
    }
    void kernel_timing(uint64_t *launches_, uint64_t *ns, uint64_t *pairs) override {
        drain_timing();
        *launches_ = timed_launches_;
        *ns = timed_ns_;
        *pairs = timed_pairs_;
    }
    void accumulate_pixel_iters() {
        for (int l = 0; l < last_nscales_; ++l)
            for (int wi = 0; wi < prm_.warps; ++wi)
                unfused_px_iters_ += (uint64_t)last_iters_[l * prm_.warps + wi] * (uint64_t)last_lv_[l].w * last_lv_[l].h;
    }
    void reset_counters() override {
        if (stats_pending_) {
            cudaEventSynchronize(stats_event_);
            stats_pending_ = false;
        }
        launches = 0;
        pixel_iters = 0;
        unfused_px_iters_ = 0;
        drain_timing();
        timed_launches_ = timed_ns_ = timed_pairs_ = 0;
        for (auto &l : lanes_) {
            l.host_ctl->px_iters_total = 0;
            l.host_ctl->px_chunks_total = 0;
        }
        pixel_chunks = 0;
    }
    uint64_t unfused_px_iters_ = 0;
    float *unfused_scratch_ = nullptr;
    static constexpr int kJobRing = 8;
    FusedJob *job_ring_h_[kJobRing] = {}, *job_ring_d_[kJobRing] = {};
    cudaEvent_t job_ring_ev_[kJobRing] = {};
    long job_ring_next_ = 0;

    static constexpr size_t kMaxPartials = 1 << 16;

    int device_, max_w_, max_h_;
    Tvl1Params prm_;
    Slab slab_;
    std::vector<float *> slots_;
    std::vector<float *> extra_slots_;
    size_t max_lv_elems_[kMaxScales] = {};
    size_t pyr_elems_ = 0, plane_elems_ = 0;
    std::vector<Lane> lanes_;
    static constexpr int kTimingRing = 256;
    cudaEvent_t timing_ev_[kTimingRing][2] = {};
    int timing_used_ = 0;
    uint64_t timed_launches_ = 0, timed_ns_ = 0, timed_pairs_ = 0;
    int last_lane_ = 0;
    // iteration logs of the last kPairLog pairs, written by the kernel into mapped host memory
    static constexpr int kPairLog = 256, kLogInts = 16 * 16;
    int *pair_log_ = nullptr, *pair_log_dev_ = nullptr;
    int pair_log_count_ = 0, last_log_slot_ = 0;
    cudaEvent_t stats_event_ = nullptr;
    bool stats_pending_ = false;
    int last_nscales_ = 0;
    LevelGeom last_lv_[kMaxScales] = {};
    int last_iters_[16 * 16] = {};
};

}  // namespace

std::unique_ptr<FlowAlgorithm> make_tvl1(int device, int max_w, int max_h) {
    return std::unique_ptr<FlowAlgorithm>(new Tvl1(device, max_w, max_h));
}

}  // namespace dfb
