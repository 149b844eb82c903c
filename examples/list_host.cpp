// list_host.cpp — the list mode of the reference (`denseflow list.txt -a=tvl1 -s=1 -b=32`) on the bare C ABI, INTEGRATION.md §2c as a
// program.  The reference turns list.txt into a vector of videos (/root/reference/tools/denseflow.cpp:54-81), skips those whose
// .done marker exists (:66-73), and its writer creates the marker after a video's last buffer (src/denseflow_gpu.cpp:456-470).
//
//   list_host <clips.raw> <width> <height> <frames_per_clip> <n_clips> <tvl1|farn> <step> <bound> <outdir> <workers> [device ...]
//
// clips.raw holds n_clips x frames_per_clip dense 8-bit gray frames (the decode stage's output).  The clips are drained by
// `workers` host threads (worker i on the i-th listed device, round robin; default device 0) from one dynamic queue.  For every
// chunk the callback writes the quantised planes as <outdir>/<clip>/flow_x_%05d.pgm / flow_y_%05d.pgm (global indices) and, on
// the video's last chunk, <outdir>/.done/<clip>.  Videos whose marker already exists are not put on the list.  Prints the
// reference's summary line (src/denseflow_gpu.cpp:494-496).
#include <sys/stat.h>

#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>

#include "denseflow_b200.h"

struct Job {
    std::string outdir;
    int w, h;
    std::vector<int> list_to_clip;  // list entry -> clip number (entries of finished videos are left out)
};

static void write_pgm(const std::string &path, const uint8_t *p, int w, int h) {
    FILE *f = std::fopen(path.c_str(), "wb");
    if (!f) return;
    std::fprintf(f, "P5\n%d %d\n255\n", w, h);
    std::fwrite(p, 1, (size_t)w * h, f);
    std::fclose(f);
}

static void on_chunk(void *user, int entry, int /*device*/, int first_flow, int n_flows, int last_chunk, uint8_t *const *qx, uint8_t *const *qy,
                     float *const * /*flows*/) {
    const Job &job = *static_cast<const Job *>(user);
    char name[512];
    const int clip = job.list_to_clip[entry];
    for (int i = 0; i < n_flows; ++i) {  // writeFlowImages: index = base_start + i (src/common.cpp:84-100)
        std::snprintf(name, sizeof name, "%s/%04d/flow_x_%05d.pgm", job.outdir.c_str(), clip, first_flow + i);
        write_pgm(name, qx[i], job.w, job.h);
        std::snprintf(name, sizeof name, "%s/%04d/flow_y_%05d.pgm", job.outdir.c_str(), clip, first_flow + i);
        write_pgm(name, qy[i], job.w, job.h);
    }
    if (last_chunk) {  // FlowBuffer::last_buffer: only now may the video be marked done
        std::snprintf(name, sizeof name, "%s/.done/%04d", job.outdir.c_str(), clip);
        FILE *f = std::fopen(name, "wb");
        if (f) std::fclose(f);
        std::printf("done video %04d\n", clip);
    }
}

int main(int argc, char **argv) {
    try {
        if (argc < 11) {
            std::fprintf(stderr, "usage: %s clips.raw width height frames_per_clip n_clips tvl1|farn step bound outdir workers [device ...]\n", argv[0]);
            return 0;
        }
        const std::string path = argv[1], algorithm = argv[6], outdir = argv[9];
        const int w = std::atoi(argv[2]), h = std::atoi(argv[3]), fpc = std::atoi(argv[4]), n_clips = std::atoi(argv[5]);
        const int step = std::atoi(argv[7]), bound = std::atoi(argv[8]), workers = std::atoi(argv[10]);
        if (bound <= 0) throw std::runtime_error("bound should > 0!");  // check_param, src/denseflow_gpu.cpp:15-18
        std::vector<uint8_t> raw((size_t)w * h * fpc * n_clips);
        FILE *f = std::fopen(path.c_str(), "rb");
        if (!f || std::fread(raw.data(), 1, raw.size(), f) != raw.size()) throw std::runtime_error("cannot read " + path);
        std::fclose(f);

        Job job{outdir, w, h, {}};
        mkdir((outdir + "/.done").c_str(), 0755);
        std::vector<std::vector<const uint8_t *>> frames;
        std::vector<dfb_clip> clips;
        for (int c = 0; c < n_clips; ++c) {
            char name[512];
            std::snprintf(name, sizeof name, "%s/.done/%04d", outdir.c_str(), c);
            struct stat st;
            if (stat(name, &st) == 0) continue;  // tools/denseflow.cpp:66-73: skip finished videos (no -f)
            std::snprintf(name, sizeof name, "%s/%04d", outdir.c_str(), c);
            mkdir(name, 0755);
            frames.emplace_back(fpc);
            for (int i = 0; i < fpc; ++i) frames.back()[i] = raw.data() + ((size_t)c * fpc + i) * w * h;
            job.list_to_clip.push_back(c);
        }
        for (auto &fr : frames) clips.push_back(dfb_clip{fr.data(), fpc, w, h});
        std::vector<int> devices(workers, 0);
        for (int i = 0; i < workers; ++i)
            if (argc > 11) devices[i] = std::atoi(argv[11 + i % (argc - 11)]);
        dfb_list_stats st{};
        char err[512] = "";
        const int rc = dfb_run_list(algorithm.c_str(), devices.data(), workers, clips.data(), (int)clips.size(), step, bound, /*chunk_flows=*/4,
                                    /*queue=*/nullptr, on_chunk, &job, &st, err, sizeof err);
        if (rc != DFB_OK) throw std::runtime_error(err);
        std::printf("%zu videos (%llu frames, %llu %s flows) processed, using %gs, decoding speed %gfps, flow speed %gfps\n", clips.size(),
                    (unsigned long long)st.frames, (unsigned long long)st.flows, algorithm.c_str(), st.seconds,
                    st.seconds > 0 ? st.frames / st.seconds : 0.0, st.seconds > 0 ? st.flows / st.seconds : 0.0);
        return 0;
    } catch (const std::exception &ex) {
        std::printf("%s\n", ex.what());  // tools/denseflow.cpp:93-96
        return 1;
    }
}
