// list_runner.cu — video-list dispatch over several GPUs (SURVEY §8 e; BASELINE.json configs[4]).
//
// The reference turns a list.txt into a vector of videos (/root/reference/tools/denseflow.cpp:54-81) and one DenseFlow
// object walks it on GPU 0 (src/denseflow_gpu.cpp:482-489), handing FlowBuffers (item_data, base_start, last_buffer;
// include/dense_flow.h:10-18) to the writer thread, which marks a video done only after its LAST buffer has been
// written (src/denseflow_gpu.cpp:456-470).  Here the same list is drained by W workers — one host thread per entry of
// `devices`, each with its own engine handle, streams and pinned staging — from ONE work queue: an in-process atomic
// counter, or a counter in POSIX shared memory when the workers are separate processes (one process per GPU).  The
// unit of work is one video (its outputs and its completion mark stay together), there is no data-path collective,
// and completion is reported per chunk with the reference's `last_buffer` meaning.
//
// This file is a client of the C ABI (dfb_create / dfb_calc_batch_host[_u8] / dfb_destroy): it adds no arithmetic.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/denseflow_b200.h"

struct dfb_queue {
    std::atomic<long> *counter = nullptr;  // lives in the shared mapping
    std::string name;
    int fd = -1;
};

namespace {
using clk = std::chrono::steady_clock;
double seconds_since(clk::time_point t0) { return std::chrono::duration<double>(clk::now() - t0).count(); }
}  // namespace

extern "C" {

int dfb_queue_open(const char *name, int create, dfb_queue **out) {
    if (!name || !out || name[0] != '/') return DFB_ERR_INVALID_ARG;
    *out = nullptr;
    const int fd = shm_open(name, create ? (O_CREAT | O_RDWR) : O_RDWR, 0600);
    if (fd < 0) return DFB_ERR_INVALID_ARG;
    if (create && ftruncate(fd, 64) != 0) {
        close(fd);
        return DFB_ERR_INVALID_ARG;
    }
    void *p = mmap(nullptr, 64, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    if (p == MAP_FAILED) {
        close(fd);
        return DFB_ERR_INVALID_ARG;
    }
    static_assert(sizeof(std::atomic<long>) == sizeof(long) && std::atomic<long>::is_always_lock_free, "plain word in shared memory");
    dfb_queue *q = new dfb_queue();
    q->counter = reinterpret_cast<std::atomic<long> *>(p);
    q->name = name;
    q->fd = fd;
    if (create) q->counter->store(0);
    *out = q;
    return DFB_OK;
}

long dfb_queue_next(dfb_queue *q) { return q ? q->counter->fetch_add(1) : -1; }

void dfb_queue_reset(dfb_queue *q) {
    if (q) q->counter->store(0);
}

void dfb_queue_close(dfb_queue *q, int unlink_name) {
    if (!q) return;
    munmap(q->counter, 64);
    close(q->fd);
    if (unlink_name) shm_unlink(q->name.c_str());
    delete q;
}

struct dfb_list_runner {
    std::string algorithm;
    int max_w = 0, max_h = 0;
    struct Worker {
        int device = 0;
        dfb_handle *h = nullptr;
        void *pinned = nullptr;  // output ring of one chunk: the copy engines write the results straight into it
        size_t pinned_bytes = 0;
    };
    std::vector<Worker> workers;
};

int dfb_list_open(const char *algorithm, const int *devices, int n_workers, int max_width, int max_height, dfb_list_runner **out, char *err,
                  size_t err_len) {
    auto set_err = [&](const std::string &m) {
        if (err && err_len) std::snprintf(err, err_len, "%s", m.c_str());
    };
    if (!out) return DFB_ERR_INVALID_ARG;
    *out = nullptr;
    if (!algorithm || !devices || n_workers < 1 || n_workers > DFB_LIST_MAX_WORKERS || max_width <= 0 || max_height <= 0) {
        set_err("dfb_list_open: bad arguments");
        return DFB_ERR_INVALID_ARG;
    }
    dfb_list_runner *r = new dfb_list_runner();
    r->algorithm = algorithm;
    r->max_w = max_width;
    r->max_h = max_height;
    for (int wi = 0; wi < n_workers; ++wi) {
        dfb_list_runner::Worker w;
        w.device = devices[wi];
        if (int rc = dfb_create(algorithm, w.device, max_width, max_height, &w.h)) {
            set_err("worker " + std::to_string(wi) + ": " + dfb_last_error(nullptr));
            dfb_list_close(r);
            return rc;
        }
        r->workers.push_back(w);
    }
    *out = r;
    return DFB_OK;
}

void dfb_list_close(dfb_list_runner *r) {
    if (!r) return;
    for (auto &w : r->workers) {
        cudaSetDevice(w.device);
        if (w.pinned) cudaFreeHost(w.pinned);
        dfb_destroy(w.h);
    }
    delete r;
}

int dfb_list_run(dfb_list_runner *r, const dfb_clip *clips, int n_clips, int step, int bound, int chunk_flows, dfb_queue *queue,
                 dfb_chunk_done_fn done, void *user, dfb_list_stats *stats, char *err, size_t err_len) {
    auto set_err = [&](const std::string &m) {
        if (err && err_len) std::snprintf(err, err_len, "%s", m.c_str());
    };
    if (!r || (!clips && n_clips > 0) || n_clips < 0 || step == 0 || bound < 0) {
        set_err("dfb_list_run: bad arguments (step must be non-zero: step 0 is the frame-extraction mode)");
        return DFB_ERR_INVALID_ARG;
    }
    if (chunk_flows <= 0) chunk_flows = 64;
    const int astep = std::abs(step);
    const int n_workers = (int)r->workers.size();
    for (int i = 0; i < n_clips; ++i) {
        if (clips[i].n_frames < 0 || clips[i].width <= 0 || clips[i].height <= 0 || (clips[i].n_frames > 0 && !clips[i].frames)) {
            set_err("dfb_list_run: clip " + std::to_string(i) + " is malformed");
            return DFB_ERR_INVALID_ARG;
        }
        if (clips[i].width > r->max_w || clips[i].height > r->max_h) {
            set_err("dfb_list_run: clip " + std::to_string(i) + " exceeds the size given to dfb_list_open");
            return DFB_ERR_SIZE;
        }
    }
    dfb_list_stats st{};
    st.workers = n_workers;
    std::atomic<long> local_next{0};
    std::atomic<int> failed{0};
    std::mutex mtx;
    std::string first_error;
    auto fail = [&](int code, const std::string &m) {
        std::lock_guard<std::mutex> lk(mtx);
        if (!failed.load()) {
            failed.store(code);
            first_error = m;
        }
    };
    const size_t px_max = (size_t)r->max_w * r->max_h;
    const size_t out_bytes = (size_t)chunk_flows * px_max * (bound > 0 ? 2 : 8);
    for (auto &w : r->workers) dfb_reset_counters(w.h);
    const auto t_start = clk::now();

    auto worker = [&](int wi) {
        dfb_list_runner::Worker &w = r->workers[wi];
        cudaSetDevice(w.device);
        if (w.pinned_bytes < out_bytes) {
            if (w.pinned) cudaFreeHost(w.pinned);
            w.pinned = nullptr;
            w.pinned_bytes = 0;
            if (cudaHostAlloc(&w.pinned, out_bytes, cudaHostAllocDefault) != cudaSuccess) {
                fail(DFB_ERR_CUDA, "worker " + std::to_string(wi) + ": pinned output allocation failed");
                return;
            }
            w.pinned_bytes = out_bytes;
        }
        uint8_t *qbuf = bound > 0 ? static_cast<uint8_t *>(w.pinned) : nullptr;
        float *fbuf = bound > 0 ? nullptr : static_cast<float *>(w.pinned);
        std::vector<uint8_t *> qx(chunk_flows), qy(chunk_flows);
        std::vector<float *> fl(chunk_flows);
        double busy = 0;
        while (!failed.load()) {
            const long idx = queue ? dfb_queue_next(queue) : local_next.fetch_add(1);
            if (idx >= n_clips) break;
            const dfb_clip &c = clips[idx];
            const auto t0 = clk::now();
            const size_t px = (size_t)c.width * c.height;
            const int M = std::max(c.n_frames - astep, 0);  // src/denseflow_gpu.cpp:308
            for (int i = 0; i < chunk_flows; ++i) {
                qx[i] = qbuf ? qbuf + (size_t)(2 * i) * px : nullptr;
                qy[i] = qbuf ? qbuf + (size_t)(2 * i + 1) * px : nullptr;
                fl[i] = fbuf ? fbuf + (size_t)i * px * 2 : nullptr;
            }
            // chunks of flows [f0, f0 + m): frames f0 .. f0 + m - 1 + |step| (the overlap the reference keeps between its
            // own <= 512-frame batches, src/denseflow_gpu.cpp:182-189,204-205), so flow indices stay global (base_start)
            int f0 = 0;
            for (;;) {
                const int m = std::max(std::min(chunk_flows, M - f0), 0);
                int rc = DFB_OK;
                if (m > 0)
                    rc = bound > 0 ? dfb_calc_batch_host_u8(w.h, c.frames + f0, m + astep, step, c.width, c.height, bound, qx.data(), qy.data())
                                   : dfb_calc_batch_host(w.h, c.frames + f0, m + astep, step, c.width, c.height, fl.data());
                if (rc != DFB_OK) {
                    fail(rc, "clip " + std::to_string(idx) + ": " + dfb_last_error(w.h));
                    break;
                }
                const int last = f0 + m >= M;
                // every output of this chunk is in host memory now; `last` is the reference's FlowBuffer::last_buffer,
                // the only point at which a video may be marked done
                if (done) done(user, (int)idx, w.device, f0, m, last, bound > 0 ? qx.data() : nullptr, bound > 0 ? qy.data() : nullptr,
                               bound > 0 ? nullptr : fl.data());
                f0 += m;
                if (last) break;
            }
            if (failed.load()) break;
            busy += seconds_since(t0);
            std::lock_guard<std::mutex> lk(mtx);
            ++st.clips;
            st.flows += (uint64_t)M;
            st.frames += (uint64_t)c.n_frames;
            ++st.clips_per_worker[wi];
            st.flows_per_worker[wi] += (uint64_t)M;
        }
        std::lock_guard<std::mutex> lk(mtx);
        st.busy_seconds_per_worker[wi] = busy;
        st.finish_seconds_per_worker[wi] = seconds_since(t_start);
    };

    if (n_clips > 0) {
        std::vector<std::thread> threads;
        for (int wi = 1; wi < n_workers; ++wi) threads.emplace_back(worker, wi);
        worker(0);
        for (auto &t : threads) t.join();
    }
    st.seconds = seconds_since(t_start);
    for (auto &w : r->workers) {
        dfb_counters c{};
        if (dfb_get_counters(w.h, &c) == DFB_OK) {
            st.kernel_launches += c.kernel_launches;
            st.h2d_bytes += c.h2d_bytes;
            st.d2h_bytes += c.d2h_bytes;
        }
    }
    if (stats) *stats = st;
    if (failed.load()) {
        set_err(first_error);
        return failed.load();
    }
    return DFB_OK;
}

int dfb_run_list(const char *algorithm, const int *devices, int n_workers, const dfb_clip *clips, int n_clips, int step, int bound,
                 int chunk_flows, dfb_queue *queue, dfb_chunk_done_fn done, void *user, dfb_list_stats *stats, char *err, size_t err_len) {
    auto set_err = [&](const std::string &m) {
        if (err && err_len) std::snprintf(err, err_len, "%s", m.c_str());
    };
    if (!algorithm || !devices || n_workers < 1 || n_workers > DFB_LIST_MAX_WORKERS || (!clips && n_clips > 0) || n_clips < 0 || step == 0 ||
        bound < 0) {
        set_err("dfb_run_list: bad arguments (step must be non-zero: step 0 is the frame-extraction mode)");
        return DFB_ERR_INVALID_ARG;
    }
    if (n_clips == 0) {  // tools/denseflow.cpp:86: nothing is created when every video of the list is already done
        dfb_list_stats st{};
        st.workers = n_workers;
        if (stats) *stats = st;
        return DFB_OK;
    }
    int max_w = 1, max_h = 1;
    for (int i = 0; i < n_clips; ++i) {
        max_w = std::max(max_w, clips[i].width);
        max_h = std::max(max_h, clips[i].height);
    }
    dfb_list_runner *r = nullptr;
    if (int rc = dfb_list_open(algorithm, devices, n_workers, max_w, max_h, &r, err, err_len)) return rc;
    const int rc = dfb_list_run(r, clips, n_clips, step, bound, chunk_flows, queue, done, user, stats, err, err_len);
    dfb_list_close(r);
    return rc;
}

}  // extern "C"
