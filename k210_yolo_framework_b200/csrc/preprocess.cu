// Image pre-processing on the GPU: the reference's aspect-preserving "letterbox" resize (tools/utils.py:372-400) for
// uint8 HWC images of any size -> uint8 [in_h, in_w, 3], the tensor the uint8 front end of the network consumes
// (k2y_net_bind_u8: per-image max + `img / np.max(img)` fused into the first convolution).
//
// The resampling restates skimage 0.15 `warp(img, aff.inverse, output_shape, order=1, mode='constant', cval=0, clip=True,
// preserve_range=True).astype('uint8')` (third-party, unpinned — see oracle/preprocess_ref.py): float64 bilinear
// interpolation between floor/ceil neighbours with zero outside the image, no half-pixel shift, clip to the input's
// [min, max] (exact-zero fill pixels stay zero when 0 is outside that range), truncation to uint8.  Every double
// operation is an explicit round-to-nearest intrinsic so that nvcc cannot contract a*b+c into an FMA: results are
// bit-identical to the numpy oracle.
#include "common.h"

namespace k2y {
namespace {

__global__ void __launch_bounds__(256) image_minmax_u8_kernel(const unsigned char *__restrict__ x, size_t n, int *__restrict__ minmax) {
    int lo = 255, hi = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int v = x[i];
        lo = min(lo, v);
        hi = max(hi, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&minmax[0], lo);
        atomicMax(&minmax[1], hi);
    }
}

struct LetterboxParams {
    const unsigned char *src;
    unsigned char *dst;
    const int *minmax;
    int sh, sw, dh, dw;
    double m00, m01, m02, m10, m11, m12;
};

__device__ __forceinline__ double pix_or_zero(const unsigned char *src, int sh, int sw, double r, double c, int ch) {
    if (!(r >= 0.0 && r < (double)sh && c >= 0.0 && c < (double)sw)) return 0.0;
    return (double)src[((size_t)(int)r * sw + (int)c) * 3 + ch];
}

__global__ void __launch_bounds__(256) letterbox_u8_kernel(const LetterboxParams p) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= p.dh * p.dw) return;
    const double r = (double)(idx / p.dw), c = (double)(idx % p.dw);
    const double x = __dadd_rn(__dadd_rn(__dmul_rn(p.m00, c), __dmul_rn(p.m01, r)), p.m02);
    const double y = __dadd_rn(__dadd_rn(__dmul_rn(p.m10, c), __dmul_rn(p.m11, r)), p.m12);
    const double minr = floor(y), minc = floor(x), maxr = ceil(y), maxc = ceil(x);
    const double dr = __dsub_rn(y, minr), dc = __dsub_rn(x, minc);
    const double wr = __dsub_rn(1.0, dr), wc = __dsub_rn(1.0, dc);
    const double lo = (double)p.minmax[0], hi = (double)p.minmax[1];
    const bool keep_fill = lo > 0.0;  // 0 (cval) lies outside [min, max]: exact-zero pixels are preserved by the clip
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        const double p00 = pix_or_zero(p.src, p.sh, p.sw, minr, minc, ch), p01 = pix_or_zero(p.src, p.sh, p.sw, minr, maxc, ch);
        const double p10 = pix_or_zero(p.src, p.sh, p.sw, maxr, minc, ch), p11 = pix_or_zero(p.src, p.sh, p.sw, maxr, maxc, ch);
        const double top = __dadd_rn(__dmul_rn(wc, p00), __dmul_rn(dc, p01));
        const double bot = __dadd_rn(__dmul_rn(wc, p10), __dmul_rn(dc, p11));
        double o = __dadd_rn(__dmul_rn(wr, top), __dmul_rn(dr, bot));
        if (!(keep_fill && o == 0.0)) o = fmin(fmax(o, lo), hi);
        p.dst[(size_t)idx * 3 + ch] = (unsigned char)(int)o;
    }
}

}  // namespace
}  // namespace k2y

extern "C" int k2y_letterbox_u8(const unsigned char *src_dev, int src_h, int src_w, const double *inv_matrix_host,
                                unsigned char *dst_dev, int dst_h, int dst_w, int *minmax_dev, void *stream) {
    using namespace k2y;
    if (!src_dev || !dst_dev || !inv_matrix_host || !minmax_dev || src_h <= 0 || src_w <= 0 || dst_h <= 0 || dst_w <= 0) {
        set_error("k2y_letterbox_u8: bad arguments");
        return K2Y_ERR_INVALID;
    }
    cudaStream_t st = (cudaStream_t)stream;
    static const int init[2] = {255, 0};
    K2Y_CUDA_CHECK(cudaMemcpyAsync(minmax_dev, init, sizeof(init), cudaMemcpyHostToDevice, st));
    const size_t n = (size_t)src_h * src_w * 3;
    int blocks = (int)((n + 256 * 16 - 1) / (256 * 16));
    if (blocks > 592) blocks = 592;
    image_minmax_u8_kernel<<<blocks, 256, 0, st>>>(src_dev, n, minmax_dev);
    LetterboxParams p;
    p.src = src_dev;
    p.dst = dst_dev;
    p.minmax = minmax_dev;
    p.sh = src_h;
    p.sw = src_w;
    p.dh = dst_h;
    p.dw = dst_w;
    p.m00 = inv_matrix_host[0];
    p.m01 = inv_matrix_host[1];
    p.m02 = inv_matrix_host[2];
    p.m10 = inv_matrix_host[3];
    p.m11 = inv_matrix_host[4];
    p.m12 = inv_matrix_host[5];
    letterbox_u8_kernel<<<(dst_h * dst_w + 255) / 256, 256, 0, st>>>(p);
    K2Y_CUDA_CHECK(cudaGetLastError());
    return K2Y_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// Precision / recall counters of the reference's training-time metrics (tools/custom.py:13-75 Yolo_Precision / Yolo_Recall),
// evaluated on device head tensors: per box, true_conf = y_true[..., 4], pred_conf = y_pred[..., 4];
//   tp += (true > thr) & (pred > thr);  fp += !(true > thr) & (pred > thr);  fn += (true > thr) & !(pred > thr).
// The reference compares the RAW predicted logit with the threshold (it computes the sigmoid and never uses it, :33-35);
// apply_sigmoid = 1 evaluates the evidently intended sigmoid(pred) > thr instead.
// ---------------------------------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256) pr_counts_kernel(const float *__restrict__ y_true, const float *__restrict__ y_pred, long long n_boxes,
                                                        int entry, float thr, int apply_sigmoid, unsigned long long *__restrict__ counts) {
    unsigned tp = 0, fp = 0, fn = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_boxes; i += (long long)gridDim.x * blockDim.x) {
        const float t = __ldg(y_true + i * entry + 4);
        float p = __ldg(y_pred + i * entry + 4);
        if (apply_sigmoid) p = __fdiv_rn(1.0f, __fadd_rn(1.0f, __double2float_rn(exp(-(double)p))));
        const bool tt = t > thr, pp = p > thr;
        tp += tt && pp;
        fp += !tt && pp;
        fn += tt && !pp;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        tp += __shfl_xor_sync(0xffffffffu, tp, o);
        fp += __shfl_xor_sync(0xffffffffu, fp, o);
        fn += __shfl_xor_sync(0xffffffffu, fn, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (tp) atomicAdd(counts + 0, (unsigned long long)tp);
        if (fp) atomicAdd(counts + 1, (unsigned long long)fp);
        if (fn) atomicAdd(counts + 2, (unsigned long long)fn);
    }
}
}  // namespace

extern "C" int k2y_pr_counts(const float *y_true_dev, const float *y_pred_dev, long long n_boxes, int entry_floats, float threshold,
                             int apply_sigmoid, unsigned long long *counts_dev, void *stream) {
    if (!y_true_dev || !y_pred_dev || !counts_dev || n_boxes < 0 || entry_floats < 5) {
        k2y::set_error("k2y_pr_counts: bad arguments (entry_floats = 5 + classes >= 5)");
        return K2Y_ERR_INVALID;
    }
    if (n_boxes == 0) return K2Y_OK;
    long long blocks = (n_boxes + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    pr_counts_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(y_true_dev, y_pred_dev, n_boxes, entry_floats, threshold,
                                                                       apply_sigmoid, counts_dev);
    K2Y_CUDA_CHECK(cudaGetLastError());
    return K2Y_OK;
}
