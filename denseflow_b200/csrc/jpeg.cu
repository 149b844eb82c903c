// jpeg.cu — GPU JPEG encode of the quantised flow planes through nvJPEG: replaces the two
// imencode(".jpg", flow_img_x / flow_img_y, ...) calls of encodeFlowMap (/root/reference/src/common.cpp:56-57) with
// OpenCV's defaults (quality 95, baseline sequential, one gray component).  Byte identity with libjpeg-turbo is
// not a goal (lossy; the reference pins no bytes) — the decoded planes are compared in the tests.
// nvJPEG is loaded lazily with dlopen so the engine itself never depends on it.
#include <dlfcn.h>
#include <nvjpeg.h>

#include <mutex>

#include "jpeg.h"

namespace dfb {

namespace {

struct Api {
    void *lib = nullptr;
    decltype(&nvjpegCreateSimple) CreateSimple = nullptr;
    decltype(&nvjpegDestroy) Destroy = nullptr;
    decltype(&nvjpegEncoderStateCreate) StateCreate = nullptr;
    decltype(&nvjpegEncoderStateDestroy) StateDestroy = nullptr;
    decltype(&nvjpegEncoderParamsCreate) ParamsCreate = nullptr;
    decltype(&nvjpegEncoderParamsDestroy) ParamsDestroy = nullptr;
    decltype(&nvjpegEncoderParamsSetQuality) SetQuality = nullptr;
    decltype(&nvjpegEncoderParamsSetSamplingFactors) SetSampling = nullptr;
    decltype(&nvjpegEncoderParamsSetOptimizedHuffman) SetOptHuff = nullptr;
    decltype(&nvjpegEncodeYUV) EncodeYUV = nullptr;
    decltype(&nvjpegEncodeRetrieveBitstream) Retrieve = nullptr;
};

Api &api() {
    static Api a;
    static std::once_flag once;
    std::call_once(once, [] {
        for (const char *name : {"libnvjpeg.so.12", "libnvjpeg.so", "/usr/local/cuda/lib64/libnvjpeg.so.12"}) {
            a.lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
            if (a.lib) break;
        }
        if (!a.lib) return;
#define DFB_SYM(field, sym) a.field = reinterpret_cast<decltype(a.field)>(dlsym(a.lib, #sym))
        DFB_SYM(CreateSimple, nvjpegCreateSimple);
        DFB_SYM(Destroy, nvjpegDestroy);
        DFB_SYM(StateCreate, nvjpegEncoderStateCreate);
        DFB_SYM(StateDestroy, nvjpegEncoderStateDestroy);
        DFB_SYM(ParamsCreate, nvjpegEncoderParamsCreate);
        DFB_SYM(ParamsDestroy, nvjpegEncoderParamsDestroy);
        DFB_SYM(SetQuality, nvjpegEncoderParamsSetQuality);
        DFB_SYM(SetSampling, nvjpegEncoderParamsSetSamplingFactors);
        DFB_SYM(SetOptHuff, nvjpegEncoderParamsSetOptimizedHuffman);
        DFB_SYM(EncodeYUV, nvjpegEncodeYUV);
        DFB_SYM(Retrieve, nvjpegEncodeRetrieveBitstream);
#undef DFB_SYM
    });
    return a;
}

void check(nvjpegStatus_t st, const char *what) {
    if (st != NVJPEG_STATUS_SUCCESS) throw std::runtime_error(std::string("nvjpeg: ") + what + " failed with status " + std::to_string((int)st));
}

}  // namespace

struct JpegEncoder::Impl {
    nvjpegHandle_t handle = nullptr;
    nvjpegEncoderState_t state = nullptr;
    nvjpegEncoderParams_t params = nullptr;
    int quality = -1;
};

JpegEncoder::JpegEncoder() : impl_(new Impl) {
    Api &a = api();
    if (!a.lib || !a.CreateSimple || !a.EncodeYUV || !a.Retrieve) throw std::runtime_error("nvjpeg: libnvjpeg.so.12 could not be loaded");
    check(a.CreateSimple(&impl_->handle), "nvjpegCreateSimple");
    check(a.StateCreate(impl_->handle, &impl_->state, nullptr), "nvjpegEncoderStateCreate");
    check(a.ParamsCreate(impl_->handle, &impl_->params, nullptr), "nvjpegEncoderParamsCreate");
}

JpegEncoder::~JpegEncoder() {
    Api &a = api();
    if (impl_->params) a.ParamsDestroy(impl_->params);
    if (impl_->state) a.StateDestroy(impl_->state);
    if (impl_->handle) a.Destroy(impl_->handle);
    delete impl_;
}

size_t JpegEncoder::encode_gray(const uint8_t *gray, size_t pitch, int w, int h, int quality, uint8_t *out, size_t out_cap,
                                cudaStream_t s) {
    Api &a = api();
    if (quality != impl_->quality) {
        check(a.SetQuality(impl_->params, quality, s), "SetQuality");
        check(a.SetSampling(impl_->params, NVJPEG_CSS_GRAY, s), "SetSamplingFactors");
        check(a.SetOptHuff(impl_->params, 0, s), "SetOptimizedHuffman");
        impl_->quality = quality;
    }
    nvjpegImage_t img{};
    img.channel[0] = const_cast<unsigned char *>(gray);
    img.pitch[0] = pitch;
    check(a.EncodeYUV(impl_->handle, impl_->state, impl_->params, &img, NVJPEG_CSS_GRAY, w, h, s), "nvjpegEncodeYUV");
    size_t len = 0;
    check(a.Retrieve(impl_->handle, impl_->state, nullptr, &len, s), "RetrieveBitstream(size)");
    DFB_CUDA(cudaStreamSynchronize(s));
    if (len > out_cap) throw std::runtime_error("nvjpeg: output buffer too small (" + std::to_string(len) + " > " + std::to_string(out_cap) + ")");
    check(a.Retrieve(impl_->handle, impl_->state, out, &len, s), "RetrieveBitstream");
    DFB_CUDA(cudaStreamSynchronize(s));
    return len;
}

}  // namespace dfb
