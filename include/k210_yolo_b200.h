/* k210_yolo_b200 — C-ABI of the B200-native YOLOv3 inference hot path.
 *
 * Drop-in boundary for the path  keras_inference.py:main -> network(...) -> predict ->
 * decode -> per-class NMS  of zhen8838/K210_Yolo_framework.  Every entry point names the
 * reference interface it replaces (paths under /root/reference).  Plain pointers and sizes
 * only; device pointers come from the caller (e.g. torch.Tensor.data_ptr()), streams are
 * cudaStream_t passed as void*.  All functions return K2Y_OK (0) or a negative error code;
 * k2y_last_error() gives the message.  There is NO CPU fallback: calls that need a GPU fail
 * with K2Y_ERR_CUDA when none is present.
 */
#ifndef K210_YOLO_B200_H
#define K210_YOLO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define K2Y_OK 0
#define K2Y_ERR_INVALID (-1)   /* bad argument / unknown name / shape mismatch */
#define K2Y_ERR_CUDA (-2)      /* CUDA runtime or driver error (incl. no device) */
#define K2Y_ERR_STATE (-3)     /* call order violated (e.g. run before finalize/bind) */
#define K2Y_ERR_NOMEM (-4)

/* Arithmetic of the dense (1x1 / 3x3) convolutions. */
#define K2Y_MATH_FP32_SIMT 0   /* fp32 FFMA on CUDA cores (exact-order reference path on the GPU) */
#define K2Y_MATH_TC_3XTF32 1   /* tcgen05 kind::tf32, hi/lo split x3 (fp32-class accuracy) */
#define K2Y_MATH_TC_TF32 2     /* tcgen05 kind::tf32, single pass */
#define K2Y_MATH_TC_BF16X3 3   /* tcgen05 kind::f16: fp32 operands split into (hi, mid) bf16 planes, 3 MMAs (~16-bit mantissa) */

const char *k2y_last_error(void);
int k2y_version(void);
/* 1 if a CUDA device is usable, 0 otherwise (never fails). */
int k2y_cuda_available(void);

/* ------------------------------------------------------------------------------------------
 * Network = the reference's model-builder "plugin" API
 *   network(input_shape=[H,W,3], anchor_num, class_num, alpha=) -> (yolo_model, yolo_model_warpper)
 *   models/yolonet.py:12 (yolo_mobilev1), :49 (yolo_mobilev2), :107 (tiny_yolo), :161 (yolo)
 * and the two methods keras_inference.py uses on it: load_weights (:80) and predict (:88).
 * ---------------------------------------------------------------------------------------- */
typedef struct k2y_net k2y_net;

/* Builds the layer graph (native graph builder).  model_def is one of the four reference
 * builder names.  Output grids are derived from the input size (H/32, H/16[, H/8]) rather than
 * the reference's hard-coded Reshape targets (yolonet.py:40-41,98-99,140-141,175-177). */
int k2y_net_create(const char *model_def, int in_h, int in_w, float alpha, int anchor_num, int class_num,
                   int max_batch, int device, k2y_net **out);
int k2y_net_destroy(k2y_net *net);

/* Weighted layers, in creation order (== Keras auto-naming order). */
typedef struct {
    char name[64];     /* Keras layer name of the conv ("conv_pw_3", "conv2d_1", ...) */
    char bn_name[64];  /* Keras name of the BatchNormalization folded behind it, "" if none */
    int32_t kind;      /* 0 dense conv, 1 depthwise conv */
    int32_t kh, kw, cin, cout, stride;
    int32_t has_bias;
} k2y_layer_info;
int k2y_net_num_layers(const k2y_net *net, int *n);
int k2y_net_layer_info(const k2y_net *net, int i, k2y_layer_info *info);

/* load_weights (keras_inference.py:80): one call per Keras variable.  layer is the Keras layer
 * name as stored in the HDF5 file, var one of kernel | depthwise_kernel | bias | gamma | beta |
 * moving_mean | moving_variance; data is host float32 in the Keras layout (HWIO kernels,
 * depthwise (3,3,C,1)); dims/ndim are checked against the graph.  Unknown layer -> K2Y_ERR_INVALID. */
int k2y_net_set_weight(k2y_net *net, const char *layer, const char *var, const float *data,
                       const int64_t *dims, int ndim);
/* Folds BatchNormalization (eps 1e-3) into per-channel scale/shift, packs and uploads.  Fails with
 * K2Y_ERR_STATE naming the first variable that was never set. */
int k2y_net_finalize(k2y_net *net);

int k2y_net_set_math(k2y_net *net, int math_mode);
int k2y_net_get_math(const k2y_net *net, int *math_mode);
/* 1 (default) = replay the layer schedule as one CUDA graph per batch size; 0 = plain launches. */
int k2y_net_set_use_graph(k2y_net *net, int use_graph);

int k2y_net_num_outputs(const k2y_net *net, int *n);
/* Per-image shape of head l of the plain yolo_model: [h, w, anchor_num*(5+class_num)]. */
int k2y_net_output_shape(const k2y_net *net, int l, int *h, int *w, int *c);
/* Bytes of activation arena needed for max_batch images. */
int k2y_net_workspace_bytes(const k2y_net *net, size_t *bytes);
/* Binds caller-owned device storage: workspace (>= workspace_bytes, 256-B aligned), the input
 * x [max_batch,H,W,3] f32 NHWC and one buffer per head [max_batch,h,w,c] f32 NHWC. */
int k2y_net_bind(k2y_net *net, void *workspace, size_t workspace_bytes, const float *x_dev,
                 float *const *heads_dev, int n_heads);
/* Optional uint8 front end (the reference's `img / np.max(img)`, tools/utils.py:405, done on the GPU): x_u8_dev is
 * [max_batch,H,W,3] uint8 letterboxed RGB, img_max_dev a scratch int32[max_batch].  Each run first reduces the per-image
 * maximum, and the first convolution reads x = u8 / max through a 256-entry table (bit-identical to the float32 input the
 * reference feeds).  Pass NULL, NULL to return to the float32 input bound with k2y_net_bind. */
int k2y_net_bind_u8(k2y_net *net, const unsigned char *x_u8_dev, int32_t *img_max_dev);
/* Re-points the float32 input at another device buffer [max_batch,H,W,3] (workspace and heads stay as bound).  One CUDA
 * graph is kept per (batch, input buffer), so alternating between two input buffers (H2D of batch i+1 while batch i runs)
 * costs nothing after the first use of each.  k2y_net_bind_u8 behaves the same way for the uint8 input. */
int k2y_net_bind_input(k2y_net *net, const float *x_dev);
/* SM budget of this net's persistent tensor-core kernels (tile planner + grid sizes); 0 (default) = the whole device.  Nets
 * that run concurrently on one GPU (several batches in flight on separate streams) should each take their share — 148 / 2 on
 * a B200 with two in flight — so that their one-CTA-per-SM grids run side by side instead of queueing behind each other.
 * Results do not depend on it bit for bit only where the tile shapes stay the same; they stay within the conv tolerance. */
int k2y_net_set_sm_limit(k2y_net *net, int sms);
/* Re-points the head outputs at another set of device buffers (same shapes as in k2y_net_bind).  One CUDA graph is kept per
 * (batch, input buffer, first head buffer): alternating between two head sets lets the decode/NMS of batch i (reading set A on
 * another stream) overlap the convolutions of batch i+1 (writing set B) at no re-capture cost. */
int k2y_net_bind_heads(k2y_net *net, float *const *heads_dev, int n_heads);
/* predict on bound device buffers (keras_inference.py:88), asynchronous on `stream`. */
int k2y_net_run(k2y_net *net, int batch, void *stream);
/* predict with HOST buffers: H2D of x, run, D2H of every head, stream-synchronised on return.
 * x_host [batch,H,W,3] f32; heads_host[l] [batch,h_l,w_l,c] f32. */
int k2y_net_predict_host(k2y_net *net, const float *x_host, int batch, float *const *heads_host, void *stream);
/* Kernel launches (graph nodes) one k2y_net_run(batch) issues. */
int k2y_net_launches_per_run(const k2y_net *net, int *n);
/* Measurement hooks (bench.py): one plain (non-graph) pass with a CUDA event between consecutive launches on
 * `stream`; ms_per_launch[i] = device time of schedule entry i (n >= schedule_len).  launch_info names launch i
 * (Keras layer name) and gives its algorithmic work per image: 2*MACs and fp32 activation bytes in+out. */
int k2y_net_schedule_len(const k2y_net *net, int *n); /* layers in the schedule (a split-K layer is 2 kernels) */
int k2y_net_profile(k2y_net *net, int batch, void *stream, float *ms_per_launch, int n);
int k2y_net_launch_info(const k2y_net *net, int i, char *name, int name_len, double *flops_per_image,
                        double *bytes_per_image);
/* Debug/parity hook: copies the (post-activation) output of the conv layer `name` for the last run
 * into host memory as [batch,h,w,c] f32.  Only valid with use_graph == 0 and keep_all (below). */
int k2y_net_set_keep_all(k2y_net *net, int keep_all);
int k2y_net_read_layer(k2y_net *net, const char *name, int batch, float *host, size_t host_floats, int *h,
                       int *w, int *c);

/* Kernel-parity hook: ONE dense convolution (the op DarknetConv2D / Conv2D lower to) on device NHWC f32 tensors,
 * synchronous.  kernel_host is Keras HWIO [k,k,c0+c1,cout]; y = act(conv(x) * scale + shift) (+ residual).
 * The input is concat(src0 [optionally nearest-upsampled x2, stored at h/2 x w/2], src1).  pad_mode: 0 = SAME
 * (stride 1), 1 = ZeroPadding2D((1,1),(1,1)) + VALID, 2 = ZeroPadding2D((1,0),(1,0)) + VALID.  act: 0 none,
 * 1 leaky(alpha), 2 relu, 3 relu6. */
int k2y_conv2d(const float *src0_dev, const float *src1_dev, const float *residual_dev, float *dst_dev,
               const float *kernel_host, const float *scale_host, const float *shift_host, int batch, int h, int w,
               int c0, int c1, int up0, int cout, int ksize, int stride, int pad_mode, int act, float alpha,
               int math_mode, void *stream);

/* ------------------------------------------------------------------------------------------
 * KERAS-dialect decode + per-class NMS  (keras_inference.py:94-135; tools/utils.py:524-547
 * tf_xywh_to_all; keras_inference.py:32-72 correct_box; tf.image.non_max_suppression).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    float ymin, xmin, ymax, xmax; /* pixels of the ORIGINAL image, as correct_box returns */
    float score;                  /* sigmoid(cls) * sigmoid(conf) */
    int32_t index;                /* flat box index: off_l + (row*W_l + col)*A + a, layer 0 first */
} k2y_det;

typedef struct {
    int32_t n_layers;          /* <= 3 */
    int32_t layer_h[3], layer_w[3];
    int32_t anchor_num, class_num;
    float anchors[3 * 8 * 2];  /* [layer][anchor][(w,h)] fraction of the net input (Helper.anchors) */
    int32_t in_h, in_w;        /* network input size (image_size) */
    float obj_thresh;          /* candidate iff score >= obj_thresh */
    float iou_thresh;          /* suppressed iff IoU > iou_thresh */
    int32_t max_per_class;     /* tf.image.non_max_suppression max_output_size (30) */
} k2y_detect_cfg;

/* Scratch bytes k2y_detect_keras needs for `batch` images. */
int k2y_detect_workspace_bytes(const k2y_detect_cfg *cfg, int batch, size_t *bytes);
/* heads_dev[l]: [batch,h_l,w_l,A*(5+C)] f32 NHWC device tensors.  image_hw_dev: [batch,2] f32
 * (orig_h, orig_w) on device.  Outputs (device): dets [batch][C][max_per_class], counts [batch][C].
 * Records of one class are in descending score (ties: ascending index). */
int k2y_detect_keras(const k2y_detect_cfg *cfg, const float *const *heads_dev, int batch,
                     const float *image_hw_dev, k2y_det *dets_dev, int32_t *counts_dev, void *workspace,
                     size_t workspace_bytes, void *stream);

/* Same, with explicit per-image strides (in 32-bit words) of the two outputs, so that one image's records and counts can
 * sit next to each other in a gather block: dets of image b start at (int32*)dets_dev + b*det_image_stride_words, counts at
 * counts_dev + b*count_image_stride_words.  k2y_detect_keras == strides (C*max_per_class*6, C). */
int k2y_detect_keras_strided(const k2y_detect_cfg *cfg, const float *const *heads_dev, int batch,
                             const float *image_hw_dev, k2y_det *dets_dev, int32_t *counts_dev,
                             long long det_image_stride_words, long long count_image_stride_words, void *workspace,
                             size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------------------------
 * Multi-GPU collation: images shard across ranks (one process per GPU), and the one exchange on the path
 * is ONE ncclAllGather of the fixed-size record blocks.  The reference is single-process (keras_inference.py:12-17), so
 * this replaces nothing there — it is what makes `main()` scale over the 8 GPUs of a box.
 *   k2y_comm_unique_id   rank 0: 128 bytes to hand to every rank (any transport)
 *   k2y_comm_create      every rank: joins the communicator on `device`
 *   k2y_allgather_detections  gather_dev = [world][bytes_per_rank]; this rank's block (filled by k2y_detect_keras_strided)
 *                        lives at rank*bytes_per_rank; in place, asynchronous on `stream`; no-op for world == 1.
 * NCCL is bound at run time (libnccl.so.2, or $K2Y_NCCL_LIB); K2Y_ERR_STATE if it cannot be found.
 * ---------------------------------------------------------------------------------------- */
typedef struct k2y_comm k2y_comm;
int k2y_comm_unique_id(unsigned char *id128);
int k2y_comm_create(const unsigned char *id128, int world, int rank, int device, k2y_comm **out);
int k2y_comm_destroy(k2y_comm *comm);
int k2y_comm_info(const k2y_comm *comm, int *world, int *rank, int *nccl_version);
int k2y_allgather_detections(k2y_comm *comm, void *gather_dev, size_t bytes_per_rank, void *stream);

/* ------------------------------------------------------------------------------------------
 * REGION_C-dialect decode + NMS, batched on device  (region_layer.c:121-283 per layer).
 * in_dev: [batch][A][5+C][H][W] f32 planar.  Outputs (device): probs [batch][N][C+1] after NMS,
 * boxes [batch][N][4] (x,y,w,h centre form, letterbox-corrected), N = A*H*W, index = a*H*W+row*W+col.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int32_t layer_w, layer_h, anchor_num, classes;
    int32_t net_w, net_h, image_w, image_h;
    float anchors[8 * 2];
    float threshold, nms_value;
} k2y_region_cfg;
int k2y_region_workspace_bytes(const k2y_region_cfg *cfg, int batch, size_t *bytes);
int k2y_region_run(const k2y_region_cfg *cfg, const float *in_dev, int batch, float *out_dev /* activations, may be NULL */,
                   float *probs_dev, float *boxes_dev, void *workspace, size_t workspace_bytes, void *stream);

/* Diagnostics: the tile planner of the tensor-core conv kernel for a GEMM of M x K x N (ksize 1 or 3; K = ksize^2 * Cin).
 * Host-only (assumes a B200 when no device is initialised): tile width, split-K factor, pipeline stages, CTA-pair mode
 * and the arithmetic mode the layer really runs in. */
int k2y_tc_plan(int M, int N, int K, int ksize, int math_mode, int *bn, int *k_splits, int *stages, int *cluster,
                int *effective_math);
/* The same under an SM budget (k2y_net_set_sm_limit; 0 = whole device); also reports the persistent grid size (CTAs). */
int k2y_tc_plan_budget(int M, int N, int K, int ksize, int math_mode, int sm_limit, int *bn, int *k_splits, int *stages,
                       int *cluster, int *effective_math, int *grid);

/* ------------------------------------------------------------------------------------------
 * Operator-level pieces of the reference's decode loop, for callers that keep keras_inference.py:94-135 and swap one
 * operator at a time (device pointers; same arithmetic as k2y_detect_keras):
 *   k2y_xywh_to_all  = tf_xywh_to_all (tools/utils.py:524-547): pred_xy / pred_wh [..][h][w][A][2] contiguous ->
 *                      xy = (sigmoid(t) + (col,row)) / (w,h), wh = exp(t) * anchor[a]   (anchors_wh_host: A (w,h) pairs)
 *   k2y_correct_box  = correct_box (keras_inference.py:32-72): -> boxes [n][4] (ymin,xmin,ymax,xmax) in image pixels
 *   k2y_nms_boxes    = tf.image.non_max_suppression(boxes [n][4] yxyx, scores [n], max_output_size, iou_threshold)
 *                      (keras_inference.py:125-126): greedy, score-descending (ties: lower index first), suppress iff
 *                      IoU > threshold; writes <= max_output_size int32 indices in selection order and their count.
 *                      boxes_dev must be 16-byte aligned; workspace from k2y_nms_workspace_bytes.
 * ---------------------------------------------------------------------------------------- */
int k2y_xywh_to_all(const float *pred_xy_dev, const float *pred_wh_dev, long long n_boxes, int layer_h, int layer_w,
                    int anchor_num, const float *anchors_wh_host, float *xy_dev, float *wh_dev, void *stream);
int k2y_correct_box(const float *xy_dev, const float *wh_dev, long long n_boxes, float in_h, float in_w, float image_h,
                    float image_w, float *boxes_dev, void *stream);
int k2y_nms_workspace_bytes(int n_boxes, int max_output_size, size_t *bytes);
int k2y_nms_boxes(const float *boxes_dev, const float *scores_dev, int n_boxes, int max_output_size, float iou_threshold,
                  int32_t *indices_dev, int32_t *count_dev, void *workspace, size_t workspace_bytes, void *stream);

/* Parity hook: the float32 exponential each decode dialect is pinned to, element-wise on device arrays.
 *   mode 0  correctly rounded exp (KERAS dialect: tools/utils.py:545-546 / keras_inference.py:101 evaluate exp and sigmoid inside
 *           TensorFlow, whose last-bit behaviour is not reproducible; this build and oracle/decode_ref.py define it as the
 *           correctly rounded value so that score order and NMS survivor sets are a function of the head tensors alone)
 *   mode 1  glibc's expf algorithm bit for bit (REGION_C dialect: region_layer.c:75,104 call libm's expf) */
int k2y_expf_eval(int mode, const float *x_dev, float *y_dev, long long n, void *stream);

/* ------------------------------------------------------------------------------------------
 * Pre-processing: aspect-preserving letterbox of ONE uint8 HWC RGB image on the device
 * (tools/utils.py:372-400: AffineTransform(scale, translation) + skimage warp order 1, zero fill, .astype('uint8')).
 * inv_matrix_host: the first two rows of inv([[s,0,tx],[0,s,ty],[0,0,1]]) as 6 doubles (host memory) - what
 * `aff.inverse` hands to warp.  dst_dev: [dst_h][dst_w][3] uint8, ready for k2y_net_bind_u8.  minmax_dev: 2 ints of
 * device scratch (receives the source image's min and max, used by warp's output clipping).
 * ---------------------------------------------------------------------------------------- */
int k2y_letterbox_u8(const unsigned char *src_dev, int src_h, int src_w, const double *inv_matrix_host,
                     unsigned char *dst_dev, int dst_h, int dst_w, int *minmax_dev, void *stream);

/* ------------------------------------------------------------------------------------------
 * Evaluation: the counters behind the reference's Yolo_Precision / Yolo_Recall metrics (tools/custom.py:13-75) on device
 * tensors.  y_true_dev / y_pred_dev: n_boxes records of entry_floats = 5 + classes floats (any [..., A, 5+C] head or label
 * tensor, flattened); element 4 is the confidence.  counts_dev[3] (uint64, caller-zeroed, accumulated into): tp, fp, fn with
 *   tp: true > thr && pred > thr;  fp: !(true > thr) && pred > thr;  fn: true > thr && !(pred > thr)
 * precision = tp / (tp + fp), recall = tp / (tp + fn) (0 when the denominator is 0).  apply_sigmoid = 0 reproduces the
 * reference (raw logit against the threshold), 1 compares sigmoid(pred).
 * ---------------------------------------------------------------------------------------- */
int k2y_pr_counts(const float *y_true_dev, const float *y_pred_dev, long long n_boxes, int entry_floats, float threshold,
                  int apply_sigmoid, unsigned long long *counts_dev, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* K210_YOLO_B200_H */
