// jpeg.h — nvJPEG-backed encoder for the quantised flow planes (see jpeg.cu).
#pragma once

#include "common.cuh"

namespace dfb {

class JpegEncoder {
  public:
    JpegEncoder();
    ~JpegEncoder();
    JpegEncoder(const JpegEncoder &) = delete;
    JpegEncoder &operator=(const JpegEncoder &) = delete;
    // gray: device pointer; out: host buffer. Returns the JPEG length in bytes. Blocks until the bitstream is on the host.
    size_t encode_gray(const uint8_t *gray, size_t pitch, int w, int h, int quality, uint8_t *out, size_t out_cap, cudaStream_t s);
    // The same in three steps over a pool of encoder states, so many planes can be in flight and the host synchronises once per
    // group instead of twice per plane:
    //   enqueue(state, ...)   asynchronous: the encode is enqueued on s
    //   length(state, s)      after s has been synchronised past the enqueue (event or stream sync): size of the bitstream
    //   fetch(state, out, s)  enqueues the copy of the bitstream into the host buffer on s (any stream; complete after a sync of s)
    void ensure_states(int n);
    void enqueue(int state, const uint8_t *gray, size_t pitch, int w, int h, int quality, cudaStream_t s);
    size_t length(int state, cudaStream_t s);
    void fetch(int state, uint8_t *out, size_t len, cudaStream_t s);

    // ---- decode (SURVEY §8 f3 remainder): imread(".jpg") of an `-if` frame folder, /root/reference/src/denseflow_gpu.cpp:154-162 ----
    // Reads width / height (and the component count) from the bitstream header.
    void image_info(const uint8_t *jpeg, size_t len, int *w, int *h, int *components);
    // Decodes a baseline / progressive JPEG held in host memory into packed 8-bit BGR in device memory (what imread returns,
    // IMREAD_COLOR), enqueued on s.  nvJPEG's IDCT / chroma upsampling are not bit-identical to libjpeg-turbo's.
    void decode_bgr(const uint8_t *jpeg, size_t len, uint8_t *bgr, size_t pitch, int w, int h, cudaStream_t s);

  private:
    struct Impl;
    Impl *impl_;
};

}  // namespace dfb
