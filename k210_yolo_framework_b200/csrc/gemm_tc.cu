// Placeholder until the tcgen05 kernels land: reports "unsupported" so every layer runs on the
// fp32 CUDA-core path.
#include "gemm_tc.h"

namespace k2y {
int tc_pack(TcWeights &, const float *, int, int) { return K2Y_OK; }
void tc_free(TcWeights &) {}
bool tc_supported(const ConvArgs &, const TcWeights &) { return false; }
cudaError_t launch_conv_tc(const ConvArgs &, const TcWeights &, int, cudaStream_t) { return cudaErrorNotSupported; }
}  // namespace k2y
