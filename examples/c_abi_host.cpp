// c_abi_host.cpp — a C++ host that uses ONLY include/denseflow_b200.h, shaped like the reference's flow + encode
// stages for one batch (DenseFlow::calc_optflows_imp, /root/reference/src/denseflow_gpu.cpp:282-370, and encodeFlowMap +
// writeFlowImages, src/common.cpp:48-64,84-100).  It is the binding of INTEGRATION.md §2 as a stand-alone program:
//
//   c_abi_host <frames.raw> <width> <height> <n_frames> <algorithm: tvl1|farn> <step> <bound> <outdir>
//
// frames.raw holds n_frames dense 8-bit gray frames.  Writes the quantised planes (what imencode would compress) as
// flow_x_%05d.pgm / flow_y_%05d.pgm with the reference's file naming (src/common.cpp:85-93: _p<step>_ / _m<|step|>_
// infixes, 0-based indices) and prints the reference's summary line format (src/denseflow_gpu.cpp:494-496).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>

#include "denseflow_b200.h"

static void check(int rc, dfb_handle *h) {
    if (rc != DFB_OK) throw std::runtime_error(dfb_last_error(h));  // reference: what() + exit 1 (tools/denseflow.cpp:93-96)
}

int main(int argc, char **argv) {
    try {
        if (argc != 9) {
            std::fprintf(stderr, "usage: %s frames.raw width height n_frames tvl1|farn step bound outdir\n", argv[0]);
            return 0;
        }
        const std::string path = argv[1], algorithm = argv[5], outdir = argv[8];
        const int w = std::atoi(argv[2]), h = std::atoi(argv[3]), n = std::atoi(argv[4]), step = std::atoi(argv[6]), bound = std::atoi(argv[7]);
        if (bound <= 0) throw std::runtime_error("bound should > 0!");  // check_param, src/denseflow_gpu.cpp:15-18
        std::vector<uint8_t> raw((size_t)w * h * n);
        FILE *f = std::fopen(path.c_str(), "rb");
        if (!f || std::fread(raw.data(), 1, raw.size(), f) != raw.size()) throw std::runtime_error("cannot read " + path);
        std::fclose(f);

        dfb_handle *alg = nullptr;
        check(dfb_create(algorithm.c_str(), 0, w, h, &alg), nullptr);
        const int astep = step < 0 ? -step : step;
        const int M = n - astep > 0 ? n - astep : 0;
        std::vector<const uint8_t *> frames(n);
        for (int i = 0; i < n; ++i) frames[i] = raw.data() + (size_t)i * w * h;
        std::vector<std::vector<uint8_t>> qx(M, std::vector<uint8_t>((size_t)w * h)), qy(M, std::vector<uint8_t>((size_t)w * h));
        std::vector<uint8_t *> px(M), py(M);
        for (int i = 0; i < M; ++i) {
            px[i] = qx[i].data();
            py[i] = qy[i].data();
        }
        const auto t0 = std::chrono::steady_clock::now();
        // flow stage + convertFlowToImage: the quantised planes come back, not the float field
        check(dfb_calc_batch_host_u8(alg, frames.data(), n, step, w, h, bound, px.data(), py.data()), alg);
        const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

        // write stage: files named as writeFlowImages does (PGM instead of JPEG keeps the example free of codecs;
        // dfb_process_bgr_batch_host / dfb_encode_jpeg_gray_device produce the JPEG bitstreams on the GPU)
        for (int i = 0; i < M; ++i) {
            char name[512];
            const int idx = step > 0 ? i : i + astep;
            for (int c = 0; c < 2; ++c) {
                if (step > 1) std::snprintf(name, sizeof name, "%s/flow_%c_p%d_%05d.pgm", outdir.c_str(), c ? 'y' : 'x', step, idx);
                else if (step < 0) std::snprintf(name, sizeof name, "%s/flow_%c_m%d_%05d.pgm", outdir.c_str(), c ? 'y' : 'x', astep, idx);
                else std::snprintf(name, sizeof name, "%s/flow_%c_%05d.pgm", outdir.c_str(), c ? 'y' : 'x', idx);
                FILE *o = std::fopen(name, "wb");
                if (!o) throw std::runtime_error(std::string("cannot write ") + name);
                std::fprintf(o, "P5\n%d %d\n255\n", w, h);
                std::fwrite(c ? py[i] : px[i], 1, (size_t)w * h, o);
                std::fclose(o);
            }
        }
        dfb_counters ctr;
        check(dfb_get_counters(alg, &ctr), alg);
        std::printf("1 videos (%d frames, %d %s flows) processed, using %fs, decoding speed %ffps, flow speed %ffps\n", n, M, algorithm.c_str(),
                    secs, n / secs, M / secs);
        std::printf("kernels launched: %llu\n", (unsigned long long)ctr.kernel_launches);
        dfb_destroy(alg);
        return 0;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 1;
    }
}
