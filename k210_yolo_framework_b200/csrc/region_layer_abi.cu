// region_layer_init / _run / _draw_boxes / _deinit with the reference's ABI (include/region_layer.h),
// the decode + NMS executed by the CUDA region kernel (detect.cu).  Mirrors the host-visible
// behaviour of /root/reference/yolo3_frame_test_public/region_layer.c: init :19-66, deinit :68-73,
// run :378-383, draw_boxes :385-404 (+ max_index :285-296).
#include <stdlib.h>
#include <string.h>

#include <map>
#include <memory>
#include <mutex>

#include "../../include/region_layer.h"
#include "common.h"

namespace {

// Device mirror of one region_layer_t.  Entries are keyed by the struct's address, remember the sizes they were allocated for
// (a struct re-initialised with another geometry gets fresh buffers) and carry their own lock: the global mutex only
// guards the map, so different layers run their GPU round trips concurrently.
struct DeviceSide {
    float *in = nullptr, *out = nullptr, *probs = nullptr, *boxes = nullptr;
    void *ws = nullptr;
    size_t ws_bytes = 0, n_out = 0, n_box = 0, n_probs = 0;
    cudaStream_t st = nullptr;
    std::mutex mu;
};
std::mutex g_mu;
std::map<region_layer_t *, std::shared_ptr<DeviceSide>> g_dev;

void free_device(DeviceSide &d) {
    cudaFree(d.in);
    cudaFree(d.out);
    cudaFree(d.probs);
    cudaFree(d.boxes);
    cudaFree(d.ws);
    if (d.st) cudaStreamDestroy(d.st);
    d.in = d.out = d.probs = d.boxes = nullptr;
    d.ws = nullptr;
    d.st = nullptr;
    d.ws_bytes = d.n_out = d.n_box = d.n_probs = 0;
}

void drop_device(region_layer_t *rl) {
    std::shared_ptr<DeviceSide> d;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_dev.find(rl);
        if (it == g_dev.end()) return;
        d = it->second;
        g_dev.erase(it);
    }
    std::lock_guard<std::mutex> lk(d->mu);
    free_device(*d);
}

k2y_region_cfg make_cfg(const region_layer_t *rl) {
    k2y_region_cfg c;
    memset(&c, 0, sizeof(c));
    c.layer_w = (int32_t)rl->layer_width;
    c.layer_h = (int32_t)rl->layer_height;
    c.anchor_num = (int32_t)rl->anchor_number;
    c.classes = (int32_t)rl->classes;
    c.net_w = (int32_t)rl->net_width;
    c.net_h = (int32_t)rl->net_height;
    c.image_w = (int32_t)rl->image_width;
    c.image_h = (int32_t)rl->image_height;
    for (uint32_t i = 0; i < 2 * rl->anchor_number && i < 16; ++i) c.anchors[i] = rl->anchor[i];
    c.threshold = rl->threshold;
    c.nms_value = rl->nms_value;
    return c;
}

}  // namespace

extern "C" int region_layer_init(region_layer_t *rl, int width, int height, int channels, int origin_width,
                                 int origin_height) {
    drop_device(rl);  // a re-initialised struct (or a reused address) must not inherit buffers sized for another layer
    rl->coords = 4;
    rl->image_width = 320;   // region_layer.c:24-25 — the firmware's display size, callers may overwrite
    rl->image_height = 224;
    rl->classes = channels / rl->anchor_number - 5;
    rl->net_width = origin_width;
    rl->net_height = origin_height;
    rl->layer_width = width;
    rl->layer_height = height;
    rl->boxes_number = rl->layer_width * rl->layer_height * rl->anchor_number;
    rl->output_number = rl->boxes_number * (rl->classes + rl->coords + 1);
    rl->output = nullptr;
    rl->boxes = nullptr;
    rl->probs_buf = nullptr;
    rl->probs = nullptr;
    int flag = 0;
    if (!(rl->output = (float *)malloc(rl->output_number * sizeof(float)))) flag = -1;
    else if (!(rl->boxes = malloc(rl->boxes_number * 4 * sizeof(float)))) flag = -2;
    else if (!(rl->probs_buf = (float *)malloc((size_t)rl->boxes_number * (rl->classes + 1) * sizeof(float)))) flag = -3;
    else if (!(rl->probs = (float **)malloc(rl->boxes_number * sizeof(float *)))) flag = -4;
    if (flag) {
        free(rl->output);
        free(rl->boxes);
        free(rl->probs_buf);
        free(rl->probs);
        rl->output = nullptr;
        rl->boxes = nullptr;
        rl->probs_buf = nullptr;
        rl->probs = nullptr;
        return flag;
    }
    for (uint32_t i = 0; i < rl->boxes_number; i++) rl->probs[i] = &(rl->probs_buf[(size_t)i * (rl->classes + 1)]);
    return 0;
}

extern "C" void region_layer_deinit(region_layer_t *rl) {
    drop_device(rl);
    free(rl->output);
    free(rl->boxes);
    free(rl->probs_buf);
    free(rl->probs);
    rl->output = nullptr;
    rl->boxes = nullptr;
    rl->probs_buf = nullptr;
    rl->probs = nullptr;
}

extern "C" void region_layer_run(region_layer_t *rl, obj_info_t *obj_info) {
    (void)obj_info;  // ignored by the reference too (region_layer_output is commented out, :382)
    const size_t n_out = rl->output_number, n_box = rl->boxes_number, n_probs = (size_t)rl->boxes_number * (rl->classes + 1);
    memset(rl->probs_buf, 0, n_probs * sizeof(float));
    if (rl->anchor_number > 8) {
        fprintf(stderr, "region_layer_run: anchor_number %u > 8 is not supported\n", rl->anchor_number);
        return;
    }
    k2y_region_cfg cfg = make_cfg(rl);
    std::shared_ptr<DeviceSide> dp;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        std::shared_ptr<DeviceSide> &slot = g_dev[rl];
        if (!slot) slot = std::make_shared<DeviceSide>();
        dp = slot;
    }
    DeviceSide &d = *dp;
    std::lock_guard<std::mutex> lk(d.mu);  // one struct is not re-entrant (as in the reference); distinct structs do not wait
    if (d.in && (d.n_out != n_out || d.n_box != n_box || d.n_probs != n_probs)) free_device(d);
    auto fail = [&](const char *what, cudaError_t e) {
        fprintf(stderr, "region_layer_run: %s: %s\n", what, cudaGetErrorString(e));
        free_device(d);  // never keep a half-built mirror
    };
    cudaError_t e;
    if (!d.in) {
        size_t ws = 0;
        if (k2y_region_workspace_bytes(&cfg, 1, &ws) != K2Y_OK) {
            fprintf(stderr, "region_layer_run: %s\n", k2y_last_error());
            return;
        }
        if ((e = cudaStreamCreateWithFlags(&d.st, cudaStreamNonBlocking)) != cudaSuccess) return fail("stream", e);
        if ((e = cudaMalloc(&d.in, n_out * sizeof(float))) != cudaSuccess) return fail("cudaMalloc", e);
        if ((e = cudaMalloc(&d.out, n_out * sizeof(float))) != cudaSuccess) return fail("cudaMalloc", e);
        if ((e = cudaMalloc(&d.probs, n_probs * sizeof(float))) != cudaSuccess) return fail("cudaMalloc", e);
        if ((e = cudaMalloc(&d.boxes, n_box * 4 * sizeof(float))) != cudaSuccess) return fail("cudaMalloc", e);
        if ((e = cudaMalloc(&d.ws, ws)) != cudaSuccess) return fail("cudaMalloc", e);
        d.ws_bytes = ws;
        d.n_out = n_out;
        d.n_box = n_box;
        d.n_probs = n_probs;
    }
    if ((e = cudaMemcpyAsync(d.in, rl->input, n_out * sizeof(float), cudaMemcpyHostToDevice, d.st)) != cudaSuccess)
        return fail("H2D", e);
    if (k2y_region_run(&cfg, d.in, 1, d.out, d.probs, d.boxes, d.ws, d.ws_bytes, d.st) != K2Y_OK) {
        fprintf(stderr, "region_layer_run: %s\n", k2y_last_error());
        return;
    }
    cudaMemcpyAsync(rl->output, d.out, n_out * sizeof(float), cudaMemcpyDeviceToHost, d.st);
    cudaMemcpyAsync(rl->boxes, d.boxes, n_box * 4 * sizeof(float), cudaMemcpyDeviceToHost, d.st);
    cudaMemcpyAsync(rl->probs_buf, d.probs, n_probs * sizeof(float), cudaMemcpyDeviceToHost, d.st);
    if ((e = cudaStreamSynchronize(d.st)) != cudaSuccess) {
        memset(rl->probs_buf, 0, n_probs * sizeof(float));
        return fail("kernel", e);
    }
}

extern "C" void region_layer_draw_boxes(region_layer_t *rl, callback_draw_box callback) {
    const float iw = (float)rl->image_width, ih = (float)rl->image_height;
    const float *boxes = (const float *)rl->boxes;
    for (uint32_t i = 0; i < rl->boxes_number; ++i) {
        const float *pr = rl->probs[i];
        int best = 0;
        float mx = pr[0];
        for (uint32_t j = 1; j < rl->classes; ++j)
            if (pr[j] > mx) {
                mx = pr[j];
                best = (int)j;
            }
        if (mx > rl->threshold) {
            const float x = boxes[4 * i], y = boxes[4 * i + 1], w = boxes[4 * i + 2], h = boxes[4 * i + 3];
            // float -> uint32 via a 64-bit truncation: what x86-64 gcc emits for the reference's
            // `uint32_t x1 = b->x * image_width - ...` (negative values wrap instead of being UB here)
            const uint32_t x1 = (uint32_t)(int64_t)(x * iw - (w * iw / 2));
            const uint32_t y1 = (uint32_t)(int64_t)(y * ih - (h * ih / 2));
            const uint32_t x2 = (uint32_t)(int64_t)(x * iw + (w * iw / 2));
            const uint32_t y2 = (uint32_t)(int64_t)(y * ih + (h * ih / 2));
            callback(x1, y1, x2, y2, (uint32_t)best, mx);
        }
    }
}
