/* ABI-compatible replacement for the firmware's region-layer API, executed on a B200.
 *
 * Replaces /root/reference/yolo3_frame_test_public/region_layer.h:44-48 (same four symbols,
 * same argument meaning, same struct layout as region_layer.h:19-39 so code written against the
 * reference — e.g. main.c:278-324 — links against libk210yolo_b200.so unchanged).
 *
 * Behaviour kept from region_layer.c: init derives classes = channels/anchor_number - 5 (:27),
 * hard-sets image_width/height to 320x224 (:24-25), allocates output/boxes/probs_buf/probs with
 * malloc and returns 0 or -1..-4 (:37-65); run = forward (sigmoid x,y,conf + class softmax),
 * box decode, letterbox correction, per-class sort + greedy NMS (:378-383); draw_boxes walks the
 * boxes, takes the arg-max class and calls back for prob > threshold (:385-404); deinit frees.
 * Difference: run() stages rl->input to the GPU, executes the CUDA region kernel and copies
 * output/boxes/probs back into the struct's host buffers; it reports CUDA failures on stderr
 * and leaves probs zeroed (the reference API has no error channel).
 */
#ifndef K210_YOLO_B200_REGION_LAYER_H
#define K210_YOLO_B200_REGION_LAYER_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    uint32_t obj_number;
    struct {
        uint32_t x1, y1, x2, y2;
        uint32_t class_id;
        float prob;
    } obj[10];
} obj_info_t;

typedef struct {
    float threshold;        /* caller-set: keep prob iff > threshold */
    float nms_value;        /* caller-set: suppress iff IoU > nms_value */
    uint32_t coords;        /* = 4 */
    uint32_t anchor_number; /* caller-set */
    float *anchor;          /* caller-owned (w,h) pairs, fraction of the net input */
    uint32_t image_width, image_height;
    uint32_t classes;
    uint32_t net_width, net_height;
    uint32_t layer_width, layer_height;
    uint32_t boxes_number;  /* layer_width*layer_height*anchor_number */
    uint32_t output_number; /* boxes_number*(classes+5) */
    void *boxes;            /* boxes_number x {x,y,w,h} float */
    float *input;           /* caller-owned [A][5+C][H][W] */
    float *output;
    float *probs_buf;       /* boxes_number x (classes+1) */
    float **probs;
} region_layer_t;

typedef void (*callback_draw_box)(uint32_t x1, uint32_t y1, uint32_t x2, uint32_t y2, uint32_t class_id,
                                  float prob);

int region_layer_init(region_layer_t *rl, int width, int height, int channels, int origin_width,
                      int origin_height);
void region_layer_deinit(region_layer_t *rl);
void region_layer_run(region_layer_t *rl, obj_info_t *obj_info);
void region_layer_draw_boxes(region_layer_t *rl, callback_draw_box callback);

#ifdef __cplusplus
}
#endif
#endif
