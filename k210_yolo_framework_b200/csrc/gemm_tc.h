// tcgen05 (5th-gen tensor core) convolution path — interface used by net.cu.
#pragma once
#include "common.h"

namespace k2y {

// Weights re-packed for the tensor-core kernels: B operand, K-major [Npad][Kpad] fp32, split into a
// tf32-exact "hi" plane and the fp32 remainder "lo" (for the 3xTF32 scheme).
struct TcWeights {
    float *d_hi = nullptr;
    float *d_lo = nullptr;
    unsigned short *d_bh = nullptr;  // bf16x3: bf16(w) and bf16(w - hi) planes, K-major [Npad][Kpad64]
    unsigned short *d_bm = nullptr;
    // the same two planes with the k index permuted inside every 32-channel chunk (k = 16u + 4m + e holds channel 8m + 4u + e):
    // the order in which the fused depthwise producer (dwpw_tc.cu) lays channels out in tensor memory; 1x1 convs only
    unsigned short *d_bh_p = nullptr;
    unsigned short *d_bm_p = nullptr;
    int K = 0, N = 0, Kpad = 0, Kpad64 = 0, Npad = 0;
};

int tc_pack(TcWeights &w, const float *kernel_kn, int K, int N, bool with_permuted = false);  // kernel_kn: [K][N] row-major (Keras HWIO flattened)
void tc_free(TcWeights &w);
bool tc_supported(const ConvArgs &a, const TcWeights &w);
size_t tc_scratch_bound(const ConvArgs &a, const TcWeights &w);           // split-K scratch a net must provide for this layer
int tc_launch_count(const ConvArgs &a, const TcWeights &w, int math_mode);  // kernels one launch_conv_tc issues
// dw != nullptr: `a` is the plain 1x1 conv that follows the depthwise layer *dw; both run in one launch (tc_dw_fusable)
cudaError_t launch_conv_tc(const ConvArgs &a, const TcWeights &w, int math_mode, cudaStream_t st, const DwArgs *dw = nullptr);
bool tc_dw_fusable(const DwArgs &dw, const ConvArgs &pw, const TcWeights &w, int math_mode);

// Fused MobileNet block (dwpw_tc.cu): depthwise 3x3 stride 1 + BN + act computed from a TMA-staged shared-memory window
// straight into tensor memory, then the 1x1 conv on the tensor cores — the depthwise output never reaches HBM.
bool dwpw_supported(const DwArgs &dw, const ConvArgs &pw, const TcWeights &w, int math_mode);
cudaError_t launch_dwpw_tc(const DwArgs &dw, const ConvArgs &pw, const TcWeights &w, cudaStream_t st);

}  // namespace k2y
