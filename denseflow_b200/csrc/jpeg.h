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

  private:
    struct Impl;
    Impl *impl_;
};

}  // namespace dfb
