// farneback.cu — Farneback path (placeholder until the kernels land).
#include "engine.h"

namespace dfb {
std::unique_ptr<FlowAlgorithm> make_farneback(int, int, int) {
    throw std::runtime_error("farn: kernels not built yet");
}
}  // namespace dfb
