// Fused decode + per-class greedy NMS, both dialects of the reference:
//   KERAS    keras_inference.py:94-135 (+ tools/utils.py:524-547, keras_inference.py:32-72)
//   REGION_C yolo3_frame_test_public/region_layer.c:121-283
// KERAS: one warp per (image, class): scan of the head tensors (score = sigmoid(cls) * sigmoid(conf)) with
// ballot-compaction of the candidates in index order, register-resident bitonic sort on (score desc, index asc)
// keys, then greedy IoU suppression resolved 32 candidates at a time (kept boxes and the chunk's candidates cached
// (min,max)-normalised in shared memory, intra-chunk suppression as bitmasks).  REGION_C: one CTA per image (softmax +
// decode by all threads, then one warp per class).  Final records are written contiguously per (image, class).
//
// Arithmetic is float32 in the reference's operation order with explicit round-to-nearest
// intrinsics where FMA contraction would change a rounding, so survivor sets are bit-identical to
// the oracle's whenever the transcendental results (expf) agree.
#include <cstdlib>
#include <vector>

#include "common.h"

namespace k2y {

namespace {

constexpr int DET_THREADS = 1024;
constexpr int DET_WARPS = 4;          // classes per CTA in the KERAS kernel
constexpr int DET_SMEM_KEYS = 4096;   // per-warp key capacity in shared memory (32 KB)
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ float sigmoidf_ref(float x) { return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x))); }

__device__ __forceinline__ unsigned long long pack_key(float score, int idx) {
    // scores are > 0 here, so the IEEE bit pattern is monotonic; larger key == earlier in the order
    return ((unsigned long long)__float_as_uint(score) << 32) | (unsigned long long)(0xffffffffu - (unsigned)idx);
}
__device__ __forceinline__ int key_index(unsigned long long k) { return (int)(0xffffffffu - (unsigned)(k & 0xffffffffull)); }
__device__ __forceinline__ float key_score(unsigned long long k) { return __uint_as_float((unsigned)(k >> 32)); }

// Sort (descending) n <= 32 keys held one per lane (lanes >= n hold 0).
__device__ __forceinline__ unsigned long long warp_sort_desc(unsigned long long key, int lane) {
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            const unsigned long long other = __shfl_xor_sync(FULL, key, j);
            const bool up = ((lane & k) == 0);      // this k-block sorts descending
            const bool lower = ((lane & j) == 0);   // this lane keeps the "first" element
            const bool take_max = (up == lower);
            key = take_max ? (key > other ? key : other) : (key < other ? key : other);
        }
    }
    return key;
}

// Bitonic sort (descending) of 32*R keys held R per lane, element e = r*32 + lane.  Exchanges at distance >= 32 are
// register-to-register inside a lane, shorter ones are warp shuffles; everything is unrolled, so the keys never leave
// the register file (the in-memory version below cost ~100 cycles per compare-exchange on the critical path).
template <int R>
__device__ __forceinline__ void warp_sort_desc_regs(unsigned long long (&key)[R], int lane) {
#pragma unroll
    for (int k = 2; k <= 32 * R; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= 32) {
                const int rj = j >> 5;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    if ((r & rj) == 0) {
                        const bool desc = ((r * 32) & k) == 0;  // k >= 64 here: decided by the register index alone
                        const unsigned long long a = key[r], b = key[r ^ rj];
                        const unsigned long long mx = a > b ? a : b, mn = a > b ? b : a;
                        key[r] = desc ? mx : mn;
                        key[r ^ rj] = desc ? mn : mx;
                    }
                }
            } else {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const unsigned long long other = __shfl_xor_sync(FULL, key[r], j);
                    const bool desc = (((r * 32 + lane) & k) == 0);
                    const bool lower = ((lane & j) == 0);
                    const bool take_max = (desc == lower);
                    key[r] = take_max ? (key[r] > other ? key[r] : other) : (key[r] < other ? key[r] : other);
                }
            }
        }
    }
}

template <int R>
__device__ __forceinline__ void warp_sort_desc_via_regs(unsigned long long *keys, int n, int lane) {
    unsigned long long k[R];
#pragma unroll
    for (int r = 0; r < R; ++r) k[r] = (r * 32 + lane < n) ? keys[r * 32 + lane] : 0ull;
    warp_sort_desc_regs<R>(k, lane);
    __syncwarp();
#pragma unroll
    for (int r = 0; r < R; ++r) keys[r * 32 + lane] = k[r];
    __syncwarp();
}

// Bitonic sort (descending) of P (power of two) keys in memory by one warp.
__device__ __forceinline__ void warp_sort_desc_mem(unsigned long long *keys, int P, int lane) {
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = lane; i < P; i += 32) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long a = keys[i], b = keys[ixj];
                    const bool desc = ((i & k) == 0);
                    if (desc ? (a < b) : (a > b)) {
                        keys[i] = b;
                        keys[ixj] = a;
                    }
                }
            }
            __syncwarp();
        }
    }
}

// tf.image.non_max_suppression's test `IoU(a, b) > thr` on (ymin,xmin,ymax,xmax) boxes (zero-area boxes -> IoU 0), without
// the division in the common case: inter/uni > thr  <=>  inter > thr*uni (uni > 0).
// The product form is only trusted outside a 1e-6 relative margin (>> the 3 roundings involved); inside it the
// exact IEEE division decides, so the boolean is bit-identical to the reference's `iou > iou_threshold`.
__device__ __forceinline__ bool iou_yxyx_gt(const float4 a, const float4 b, float thr) {
    const float ymin_a = fminf(a.x, a.z), ymax_a = fmaxf(a.x, a.z);
    const float xmin_a = fminf(a.y, a.w), xmax_a = fmaxf(a.y, a.w);
    const float ymin_b = fminf(b.x, b.z), ymax_b = fmaxf(b.x, b.z);
    const float xmin_b = fminf(b.y, b.w), xmax_b = fmaxf(b.y, b.w);
    const float area_a = __fmul_rn(__fsub_rn(ymax_a, ymin_a), __fsub_rn(xmax_a, xmin_a));
    const float area_b = __fmul_rn(__fsub_rn(ymax_b, ymin_b), __fsub_rn(xmax_b, xmin_b));
    if (area_a <= 0.f || area_b <= 0.f) return 0.f > thr;
    const float iy = fmaxf(__fsub_rn(fminf(ymax_a, ymax_b), fmaxf(ymin_a, ymin_b)), 0.f);
    const float ix = fmaxf(__fsub_rn(fminf(xmax_a, xmax_b), fmaxf(xmin_a, xmin_b)), 0.f);
    const float inter = __fmul_rn(iy, ix);
    const float uni = __fsub_rn(__fadd_rn(area_a, area_b), inter);
    if (thr >= 0.f && uni > 0.f) {
        const float t = __fmul_rn(thr, uni);
        if (inter > __fmul_rn(t, 1.000001f)) return true;
        if (inter < __fmul_rn(t, 0.999999f)) return false;
    }
    return __fdiv_rn(inter, uni) > thr;
}

// Same predicate on boxes already (min,max)-normalised with their areas cached (computed exactly as above, once per
// box instead of once per pair).
__device__ __forceinline__ float4 norm_box(const float4 a, float &area) {
    float4 n;
    n.x = fminf(a.x, a.z);
    n.y = fminf(a.y, a.w);
    n.z = fmaxf(a.x, a.z);
    n.w = fmaxf(a.y, a.w);
    area = __fmul_rn(__fsub_rn(n.z, n.x), __fsub_rn(n.w, n.y));
    return n;
}
__device__ __forceinline__ bool iou_norm_gt(const float4 a, float area_a, const float4 b, float area_b, float thr) {
    if (area_a <= 0.f || area_b <= 0.f) return 0.f > thr;
    const float iy = fmaxf(__fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)), 0.f);
    const float ix = fmaxf(__fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)), 0.f);
    const float inter = __fmul_rn(iy, ix);
    const float uni = __fsub_rn(__fadd_rn(area_a, area_b), inter);
    if (thr >= 0.f && uni > 0.f) {
        const float t = __fmul_rn(thr, uni);
        if (inter > __fmul_rn(t, 1.000001f)) return true;
        if (inter < __fmul_rn(t, 0.999999f)) return false;
    }
    return __fdiv_rn(inter, uni) > thr;
}

// region_layer.c box_iou on centre-form (x,y,w,h) boxes (:228-254).
__device__ __forceinline__ float overlap_c(float x1, float w1, float x2, float w2) {
    const float l1 = __fsub_rn(x1, __fmul_rn(w1, 0.5f));
    const float l2 = __fsub_rn(x2, __fmul_rn(w2, 0.5f));
    const float left = l1 > l2 ? l1 : l2;
    const float r1 = __fadd_rn(x1, __fmul_rn(w1, 0.5f));
    const float r2 = __fadd_rn(x2, __fmul_rn(w2, 0.5f));
    const float right = r1 < r2 ? r1 : r2;
    return __fsub_rn(right, left);
}
__device__ __forceinline__ float iou_center(const float4 a, const float4 b) {
    const float w = overlap_c(a.x, a.z, b.x, b.z);
    const float h = overlap_c(a.y, a.w, b.y, b.w);
    const float inter = (w < 0.f || h < 0.f) ? 0.f : __fmul_rn(w, h);
    const float uni = __fsub_rn(__fadd_rn(__fmul_rn(a.z, a.w), __fmul_rn(b.z, b.w)), inter);
    return __fdiv_rn(inter, uni);
}

struct KerasParams {
    const float *heads[3];
    int lh[3], lw[3], loff[4];
    int n_layers, A, C, nbox, P;
    float anchors[48];
    float in_h, in_w;
    float obj, iou;
    int maxk;
    const float *image_hw;
    k2y_det *dets;
    int *counts;
    unsigned long long *keys_global;  // [B][C][P] — only used when P does not fit shared memory
    int keys_in_smem;
    long long *trace;                 // optional [B][C][4]: cycles of scan / sort / nms and the candidate count (K2Y_DET_TRACE=1)
};

// Collects the candidates of one class (score passes `pred`) in index order, sorts them.
// Returns n; on return either `rkey` holds the sorted keys (n <= 32) or keys[0..n) does.
template <bool GE>
__device__ __forceinline__ int gather_sorted(const float *sc, int nbox, float thr, unsigned long long *keys, int lane,
                                             unsigned long long &rkey) {
    int n = 0;
    for (int base = 0; base < nbox; base += 32) {
        const int i = base + lane;
        const float s = (i < nbox) ? sc[i] : -1.f;
        const bool pass = (i < nbox) && (GE ? (s >= thr) : (s > thr));
        const unsigned m = __ballot_sync(FULL, pass);
        if (pass) keys[n + __popc(m & ((1u << lane) - 1u))] = pack_key(s, i);
        n += __popc(m);
    }
    __syncwarp();
    rkey = 0ull;
    if (n <= 32) {
        if (lane < n) rkey = keys[lane];
        rkey = warp_sort_desc(rkey, lane);
    } else {
        int P = 64;
        while (P < n) P <<= 1;
        for (int i = n + lane; i < P; i += 32) keys[i] = 0ull;
        __syncwarp();
        warp_sort_desc_mem(keys, P, lane);
    }
    return n;
}

struct BoxXform {  // correct_box constants of one image (keras_inference.py:53-58)
    float img_h, img_w, off_y, off_x, sc_y, sc_x;
};

__device__ __forceinline__ const float *box_entry(const KerasParams &p, int b, int box, int &l, int &a, int &col, int &row) {
    l = 0;
    if (p.n_layers > 1 && box >= p.loff[1]) l = 1;
    if (p.n_layers > 2 && box >= p.loff[2]) l = 2;
    const int local = box - p.loff[l];
    a = local % p.A;
    const int cell = local / p.A;
    const int W = p.lw[l];
    col = cell % W;
    row = cell / W;
    return p.heads[l] + ((size_t)((size_t)b * p.lh[l] * W + cell) * p.A + a) * (5 + p.C);
}

// tf_xywh_to_all (tools/utils.py:545-546) + correct_box (keras_inference.py:59-71) for one box.
__device__ __forceinline__ float4 decode_box(const KerasParams &p, const BoxXform &x, int b, int box) {
    int l, a, col, row;
    const float *e = box_entry(p, b, box, l, a, col, row);
    const float tx = __ldg(e), ty = __ldg(e + 1), tw = __ldg(e + 2), th = __ldg(e + 3);
    const float bx = __fdiv_rn(__fadd_rn(sigmoidf_ref(tx), (float)col), (float)p.lw[l]);
    const float by = __fdiv_rn(__fadd_rn(sigmoidf_ref(ty), (float)row), (float)p.lh[l]);
    const float bw = __fmul_rn(expf(tw), p.anchors[(l * p.A + a) * 2]);
    const float bh = __fmul_rn(expf(th), p.anchors[(l * p.A + a) * 2 + 1]);
    const float cy = __fmul_rn(__fsub_rn(by, x.off_y), x.sc_y), cx = __fmul_rn(__fsub_rn(bx, x.off_x), x.sc_x);
    const float hh2 = __fdiv_rn(__fmul_rn(bh, x.sc_y), 2.0f), ww2 = __fdiv_rn(__fmul_rn(bw, x.sc_x), 2.0f);
    float4 r;
    r.x = __fmul_rn(__fsub_rn(cy, hh2), x.img_h);
    r.y = __fmul_rn(__fsub_rn(cx, ww2), x.img_w);
    r.z = __fmul_rn(__fadd_rn(cy, hh2), x.img_h);
    r.w = __fmul_rn(__fadd_rn(cx, ww2), x.img_w);
    return r;
}

// grid = (ceil(C / warps), B): one warp owns one (image, class).  It scans the head tensors for its class
// (score = sigmoid(cls) * sigmoid(conf)), compacts the candidates in index order into sort keys, sorts them, decodes
// the candidate boxes 32 at a time (one per lane) and runs the greedy suppression with shuffles.
__global__ void __launch_bounds__(DET_WARPS * 32) detect_keras_kernel(const KerasParams p) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ __align__(16) unsigned long long smem_keys[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.y;
    const int c = blockIdx.x * DET_WARPS + warp;
    if (c >= p.C) return;
    unsigned long long *keys = p.keys_in_smem ? smem_keys + (size_t)warp * p.P : p.keys_global + ((size_t)b * p.C + c) * p.P;

    BoxXform x;
    x.img_h = p.image_hw[2 * b];
    x.img_w = p.image_hw[2 * b + 1];
    const float r = fminf(__fdiv_rn(p.in_h, x.img_h), __fdiv_rn(p.in_w, x.img_w));
    const float new_h = rintf(__fmul_rn(x.img_h, r)), new_w = rintf(__fmul_rn(x.img_w, r));
    x.off_y = __fdiv_rn(__fdiv_rn(__fsub_rn(p.in_h, new_h), 2.0f), p.in_h);
    x.off_x = __fdiv_rn(__fdiv_rn(__fsub_rn(p.in_w, new_w), 2.0f), p.in_w);
    x.sc_y = __fdiv_rn(p.in_h, new_h);
    x.sc_x = __fdiv_rn(p.in_w, new_w);

    // ---- scan: candidates of class c in index order (4 x 32 boxes per step so that the loads overlap) ----
    const long long tc0 = p.trace ? clock64() : 0;
    int n = 0;
    for (int base = 0; base < p.nbox; base += 128) {
        float lc[4], lk[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int box = base + u * 32 + lane;
            lc[u] = 0.f;
            lk[u] = 0.f;
            if (box < p.nbox) {
                // entry pointer needs no (row, col, anchor) split: boxes of a layer are contiguous [cell][anchor] records
                int l = 0;
                if (p.n_layers > 1 && box >= p.loff[1]) l = 1;
                if (p.n_layers > 2 && box >= p.loff[2]) l = 2;
                const float *e = p.heads[l] + ((size_t)b * (p.loff[l + 1] - p.loff[l]) + (box - p.loff[l])) * (5 + p.C);
                lc[u] = __ldg(e + 4);
                lk[u] = __ldg(e + 5 + c);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int box = base + u * 32 + lane;
            bool pass = false;
            float s = 0.f;
            if (box < p.nbox) {
                s = __fmul_rn(sigmoidf_ref(lk[u]), sigmoidf_ref(lc[u]));
                pass = s >= p.obj;
            }
            const unsigned m = __ballot_sync(FULL, pass);
            if (pass) keys[n + __popc(m & ((1u << lane) - 1u))] = pack_key(s, box);
            n += __popc(m);
        }
    }
    __syncwarp();
    const long long tc1 = p.trace ? clock64() : 0;
    unsigned long long rkey = 0ull;
    if (n <= 32) {
        if (lane < n) rkey = keys[lane];
        rkey = warp_sort_desc(rkey, lane);
    } else if (n <= 64) {
        warp_sort_desc_via_regs<2>(keys, n, lane);
    } else if (n <= 128) {
        warp_sort_desc_via_regs<4>(keys, n, lane);
    } else if (n <= 256) {
        warp_sort_desc_via_regs<8>(keys, n, lane);
    } else if (n <= 512) {
        warp_sort_desc_via_regs<16>(keys, n, lane);
    } else if (n <= 1024 && p.P >= 1024) {
        warp_sort_desc_via_regs<32>(keys, n, lane);
    } else {
        int P = 64;
        while (P < n) P <<= 1;
        for (int i = n + lane; i < P; i += 32) keys[i] = 0ull;
        __syncwarp();
        warp_sort_desc_mem(keys, P, lane);
    }

    const long long tc2 = p.trace ? clock64() : 0;
    // ---- greedy NMS over the sorted candidates, 32 at a time ----
    // Per chunk: (a) every lane tests ITS candidate against the boxes kept so far, (b) every lane computes the bitmask
    // of later candidates of the chunk its box would suppress, (c) a short warp-uniform scan over the chunk resolves
    // the greedy order from the masks.  The IoU work is thereby done 32-wide instead of one candidate at a time;
    // the result is exactly the sequential algorithm's (a candidate is kept iff no earlier KEPT box overlaps it).
    k2y_det *out = p.dets + ((size_t)b * p.C + c) * p.maxk;
    __shared__ float4 s_kbox[DET_WARPS][32], s_cbox[DET_WARPS][32];   // kept / candidate boxes, (min,max)-normalised
    __shared__ float s_karea[DET_WARPS][32], s_carea[DET_WARPS][32];
    __shared__ float4 s_orig[DET_WARPS][32];            // candidate boxes as decoded (the records keep these)
    __shared__ unsigned long long s_key[DET_WARPS][32];
    __shared__ unsigned s_mask[DET_WARPS][32];
    float4 *kbox = s_kbox[warp], *cbox = s_cbox[warp];
    float *karea = s_karea[warp], *carea = s_carea[warp];
    int nsel = 0;
    long long t_dec = 0, t_a = 0, t_b = 0, t_c = 0;
    for (int base = 0; base < n && nsel < p.maxk; base += 32) {
        const long long q0 = p.trace ? clock64() : 0;
        const int i = base + lane;
        const unsigned long long mykey = (n <= 32) ? rkey : (i < n ? keys[i] : 0ull);
        const int cnt = min(32, n - base);
        float4 cand = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < n) cand = decode_box(p, x, b, key_index(mykey));
        float my_area;
        const float4 my = norm_box(cand, my_area);
        cbox[lane] = my;
        carea[lane] = my_area;
        __syncwarp();
        const long long q1 = p.trace ? clock64() : 0;
        // (a) against the kept set: two kept boxes per step (independent chains), stop as soon as the chunk is dead
        bool dead = lane >= cnt;
        const int nreg = min(nsel, 32);
        for (int s2 = 0; s2 < nreg; s2 += 4) {  // four independent IoU chains per step
            if (__ballot_sync(FULL, !dead) == 0u) break;
            bool hit = false;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int si = s2 + u < nreg ? s2 + u : s2;
                hit |= iou_norm_gt(my, my_area, kbox[si], karea[si], p.iou);
            }
            dead = dead || hit;
        }
        for (int s2 = 32; s2 < nsel; ++s2) {  // only when max_per_class > 32
            const k2y_det d = out[s2];
            if (!dead && iou_yxyx_gt(cand, make_float4(d.ymin, d.xmin, d.ymax, d.xmax), p.iou)) dead = true;
        }
        const long long q2 = p.trace ? clock64() : 0;
        // (b) suppression masks inside the chunk — only candidates that survived (a) can suppress anything
        const unsigned alive = ~__ballot_sync(FULL, dead);
        unsigned mask = 0u;
        for (unsigned mm = alive; mm;) {
            const int j0 = __ffs(mm) - 1;
            mm &= mm - 1u;
            const int j1 = mm ? __ffs(mm) - 1 : j0;
            mm &= mm - 1u;
            const bool h0 = iou_norm_gt(cbox[j0], carea[j0], my, my_area, p.iou);
            const bool h1 = iou_norm_gt(cbox[j1], carea[j1], my, my_area, p.iou);
            if (!dead) {
                if (j0 > lane && h0) mask |= 1u << j0;
                if (j1 > lane && h1) mask |= 1u << j1;
            }
        }
        s_mask[warp][lane] = mask;
        s_key[warp][lane] = mykey;
        s_orig[warp][lane] = cand;
        __syncwarp();
        const long long q3 = p.trace ? clock64() : 0;
        // (c) resolve in score order — a short warp-uniform scan; lane 0 records the survivors
        unsigned remv = ~alive;
        for (unsigned mm = alive; mm && nsel < p.maxk; mm &= mm - 1u) {
            const int j = __ffs(mm) - 1;
            if ((remv >> j) & 1u) continue;
            remv |= s_mask[warp][j];
            if (lane == 0) {
                if (nsel < 32) {
                    kbox[nsel] = cbox[j];
                    karea[nsel] = carea[j];
                }
                const float4 kb = s_orig[warp][j];
                const unsigned long long key = s_key[warp][j];
                k2y_det d;
                d.ymin = kb.x;
                d.xmin = kb.y;
                d.ymax = kb.z;
                d.xmax = kb.w;
                d.score = key_score(key);
                d.index = key_index(key);
                out[nsel] = d;
            }
            ++nsel;
        }
        __syncwarp();
        if (p.trace) {
            const long long q4 = clock64();
            t_dec += q1 - q0;
            t_a += q2 - q1;
            t_b += q3 - q2;
            t_c += q4 - q3;
        }
    }
    if (lane == 0) p.counts[b * p.C + c] = nsel;
    if (p.trace && lane == 0) {
        long long *o = p.trace + ((size_t)b * p.C + c) * 4;
        o[0] = tc1 - tc0;
        o[1] = tc2 - tc1;
        o[2] = clock64() - tc2;
        o[3] = n;
    }
}

struct RegionParams {
    const float *in;
    float *out;                 // may be null
    float *probs;               // [B][N][C+1]
    float4 *boxes;              // [B][N]
    float *scores;              // ws [B][C][N]
    unsigned long long *keys;   // ws [B][C][P]
    int *kept;                  // ws [B][C][N]
    int W, H, A, C, N, P;
    float anchors[16];
    float thr, nms;
    double dx, dy, sxw, syh;    // correct_region_boxes constants (region_layer.c:158-159)
    float wscale, hscale;       // (:160-161)
};

__global__ void __launch_bounds__(DET_THREADS) region_kernel(const RegionParams p) {
    const int b = blockIdx.x, tid = threadIdx.x;
    const int wh = p.W * p.H, E = 5 + p.C;
    const float *in = p.in + (size_t)b * p.A * E * wh;
    float *outp = p.out ? p.out + (size_t)b * p.A * E * wh : nullptr;
    float *probs = p.probs + (size_t)b * p.N * (p.C + 1);
    float4 *boxes = p.boxes + (size_t)b * p.N;
    float *scores = p.scores + (size_t)b * p.C * p.N;

    // ---- phase 1: forward_region_layer + get_region_boxes + correct_region_boxes ----
    for (int index = tid; index < p.N; index += DET_THREADS) {
        const int a = index / wh, loc = index - a * wh;
        const int row = loc / p.W, col = loc - row * p.W;
        const float *e = in + (size_t)a * E * wh + loc;
        const float sx = sigmoidf_ref(__ldg(e)), sy = sigmoidf_ref(__ldg(e + wh));
        const float tw = __ldg(e + 2 * wh), th = __ldg(e + 3 * wh);
        const float conf = sigmoidf_ref(__ldg(e + 4 * wh));
        float largest = __ldg(e + 5 * wh);
        for (int j = 1; j < p.C; ++j) largest = fmaxf(largest, __ldg(e + (5 + j) * wh));
        float sum = 0.f;
        for (int j = 0; j < p.C; ++j) sum = __fadd_rn(sum, expf(__fsub_rn(__ldg(e + (5 + j) * wh), largest)));
        float mx = 0.f;
        for (int j = 0; j < p.C; ++j) {
            const float sm = __fdiv_rn(expf(__fsub_rn(__ldg(e + (5 + j) * wh), largest)), sum);
            const float prob = __fmul_rn(conf, sm);
            const float kept = (prob > p.thr) ? prob : 0.f;
            probs[(size_t)index * (p.C + 1) + j] = kept;
            scores[(size_t)j * p.N + index] = kept;
            if (prob > mx) mx = prob;
            if (outp) outp[(size_t)a * E * wh + (5 + j) * wh + loc] = sm;
        }
        probs[(size_t)index * (p.C + 1) + p.C] = mx;
        if (outp) {
            float *o = outp + (size_t)a * E * wh + loc;
            o[0] = sx;
            o[wh] = sy;
            o[2 * wh] = tw;
            o[3 * wh] = th;
            o[4 * wh] = conf;
        }
        const float bx = __fdiv_rn(__fadd_rn((float)col, sx), (float)p.W);
        const float by = __fdiv_rn(__fadd_rn((float)row, sy), (float)p.H);
        const float bw = __fmul_rn(expf(tw), p.anchors[2 * a]);
        const float bh = __fmul_rn(expf(th), p.anchors[2 * a + 1]);
        float4 bb;
        bb.x = (float)__ddiv_rn(__dsub_rn((double)bx, p.dx), p.sxw);
        bb.y = (float)__ddiv_rn(__dsub_rn((double)by, p.dy), p.syh);
        bb.z = __fmul_rn(bw, p.wscale);
        bb.w = __fmul_rn(bh, p.hscale);
        boxes[index] = bb;
    }
    __syncthreads();

    // ---- phase 2: do_nms_sort, one warp per class, no output cap ----
    const int warp = tid >> 5, lane = tid & 31, nwarps = DET_THREADS >> 5;
    for (int k = warp; k < p.C; k += nwarps) {
        unsigned long long *keys = p.keys + ((size_t)b * p.C + k) * p.P;
        int *kept = p.kept + ((size_t)b * p.C + k) * p.N;
        unsigned long long rkey;
        const int n = gather_sorted<false>(scores + (size_t)k * p.N, p.N, 0.f, keys, lane, rkey);
        float4 mybox = make_float4(0.f, 0.f, 0.f, 0.f);
        int nsel = 0;
        for (int i = 0; i < n; ++i) {
            const unsigned long long key = (n <= 32) ? __shfl_sync(FULL, rkey, i) : keys[i];
            const int idx = key_index(key);
            const float4 cb = boxes[idx];
            // box_iou(a = kept, b = candidate) — argument order as in the reference loop
            bool sup = (lane < nsel) && (iou_center(mybox, cb) > p.nms);
            bool any = __ballot_sync(FULL, sup) != 0u;
            for (int base = 32; !any && base < nsel; base += 32) {
                const int j = base + lane;
                sup = (j < nsel) && (iou_center(boxes[kept[j]], cb) > p.nms);
                any = __ballot_sync(FULL, sup) != 0u;
            }
            if (any) {
                if (lane == 0) probs[(size_t)idx * (p.C + 1) + k] = 0.f;
            } else {
                if (lane == nsel) mybox = cb;
                if (lane == 0) kept[nsel] = idx;
                ++nsel;
                __syncwarp();
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Operator-level entry points of the reference's decode loop (keras_inference.py:94-135), for callers that keep that loop
// and replace one operator at a time.  Same arithmetic (intrinsics, operation order) as the fused kernel above.
// ---------------------------------------------------------------------------------------------------------------------
struct OpsAnchors {
    float wh[16];  // (w, h) pairs of one layer, A <= 8
};

// tf_xywh_to_all (tools/utils.py:524-547): raw [.., h, w, A, 2] xy / wh head slices -> xy in [0, 1], wh as input fraction.
__global__ void __launch_bounds__(256) xywh_to_all_kernel(const float2 *__restrict__ pred_xy, const float2 *__restrict__ pred_wh,
                                                          float2 *__restrict__ xy, float2 *__restrict__ wh, long long total, int H,
                                                          int W, int A, const OpsAnchors anc) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int a = (int)(i % A);
    const int cell = (int)((i / A) % ((long long)H * W));
    const int col = cell % W, row = cell / W;
    const float2 t = pred_xy[i], u = pred_wh[i];
    float2 o;
    o.x = __fdiv_rn(__fadd_rn(sigmoidf_ref(t.x), (float)col), (float)W);
    o.y = __fdiv_rn(__fadd_rn(sigmoidf_ref(t.y), (float)row), (float)H);
    xy[i] = o;
    o.x = __fmul_rn(expf(u.x), anc.wh[2 * a]);
    o.y = __fmul_rn(expf(u.y), anc.wh[2 * a + 1]);
    wh[i] = o;
}

// correct_box (keras_inference.py:32-72): (xy, wh) relative to the letterboxed input -> (ymin, xmin, ymax, xmax) in pixels
// of the original image.
__global__ void __launch_bounds__(256) correct_box_kernel(const float2 *__restrict__ xy, const float2 *__restrict__ wh,
                                                          float4 *__restrict__ boxes, long long total, float in_h, float in_w,
                                                          float img_h, float img_w) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const float r = fminf(__fdiv_rn(in_h, img_h), __fdiv_rn(in_w, img_w));
    const float new_h = rintf(__fmul_rn(img_h, r)), new_w = rintf(__fmul_rn(img_w, r));
    const float off_y = __fdiv_rn(__fdiv_rn(__fsub_rn(in_h, new_h), 2.0f), in_h);
    const float off_x = __fdiv_rn(__fdiv_rn(__fsub_rn(in_w, new_w), 2.0f), in_w);
    const float sc_y = __fdiv_rn(in_h, new_h), sc_x = __fdiv_rn(in_w, new_w);
    const float2 c = xy[i], s2 = wh[i];
    const float cy = __fmul_rn(__fsub_rn(c.y, off_y), sc_y), cx = __fmul_rn(__fsub_rn(c.x, off_x), sc_x);
    const float hh2 = __fdiv_rn(__fmul_rn(s2.y, sc_y), 2.0f), ww2 = __fdiv_rn(__fmul_rn(s2.x, sc_x), 2.0f);
    float4 o;
    o.x = __fmul_rn(__fsub_rn(cy, hh2), img_h);
    o.y = __fmul_rn(__fsub_rn(cx, ww2), img_w);
    o.z = __fmul_rn(__fadd_rn(cy, hh2), img_h);
    o.w = __fmul_rn(__fadd_rn(cx, ww2), img_w);
    boxes[i] = o;
}

// tf.image.non_max_suppression(boxes, scores, max_output_size, iou_threshold) (score_threshold = -inf): one warp sorts
// every box by (score desc, index asc) and runs the greedy selection, lanes testing the kept boxes in parallel.
__device__ __forceinline__ unsigned long long pack_key_any(float score, int idx) {
    unsigned u = __float_as_uint(score);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // total order of finite floats and infinities as unsigned
    return ((unsigned long long)u << 32) | (unsigned long long)(0xffffffffu - (unsigned)idx);
}

__global__ void __launch_bounds__(32) nms_boxes_kernel(const float4 *__restrict__ boxes, const float *__restrict__ scores, int n, int P,
                                                       int max_out, float thr, unsigned long long *keys, float4 *kept, float *kept_area,
                                                       int *__restrict__ out_idx, int *__restrict__ out_count) {
    const int lane = threadIdx.x;
    for (int i = lane; i < P; i += 32) keys[i] = i < n ? pack_key_any(scores[i], i) : 0ull;
    __syncwarp();
    warp_sort_desc_mem(keys, P, lane);
    int nsel = 0;
    for (int t = 0; t < n && nsel < max_out; ++t) {
        const int idx = key_index(keys[t]);
        float area;
        const float4 c = norm_box(boxes[idx], area);
        bool hit = false;
        for (int j = lane; j < nsel; j += 32) hit = hit || iou_norm_gt(c, area, kept[j], kept_area[j], thr);
        if (!__any_sync(FULL, hit)) {
            if (lane == 0) {
                kept[nsel] = c;
                kept_area[nsel] = area;
                out_idx[nsel] = idx;
            }
            ++nsel;
            __syncwarp();
        }
    }
    if (lane == 0) *out_count = nsel;
}

inline int next_pow2(int v) {
    int p = 64;
    while (p < v) p <<= 1;
    return p;
}
inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

}  // namespace

}  // namespace k2y

using namespace k2y;

extern "C" int k2y_detect_workspace_bytes(const k2y_detect_cfg *cfg, int batch, size_t *bytes) {
    if (!cfg || !bytes || batch <= 0 || cfg->n_layers < 1 || cfg->n_layers > 3) {
        set_error("k2y_detect_workspace_bytes: bad arguments");
        return K2Y_ERR_INVALID;
    }
    size_t nbox = 0;
    for (int l = 0; l < cfg->n_layers; ++l) nbox += (size_t)cfg->layer_h[l] * cfg->layer_w[l] * cfg->anchor_num;
    const size_t P = next_pow2((int)nbox);
    // sort keys live in shared memory up to DET_SMEM_KEYS boxes per image; larger grids spill them to this workspace
    *bytes = 256 + (P > (size_t)DET_SMEM_KEYS ? align256((size_t)batch * cfg->class_num * P * sizeof(unsigned long long)) : 0);
    return K2Y_OK;
}

extern "C" int k2y_detect_keras(const k2y_detect_cfg *cfg, const float *const *heads_dev, int batch,
                                const float *image_hw_dev, k2y_det *dets_dev, int32_t *counts_dev, void *workspace,
                                size_t workspace_bytes, void *stream) {
    size_t need = 0;
    int rc = k2y_detect_workspace_bytes(cfg, batch, &need);
    if (rc != K2Y_OK) return rc;
    if (!heads_dev || !image_hw_dev || !dets_dev || !counts_dev || !workspace || workspace_bytes < need) {
        set_error("k2y_detect_keras: null pointer or workspace too small (%zu < %zu)", workspace_bytes, need);
        return K2Y_ERR_INVALID;
    }
    if (cfg->anchor_num < 1 || cfg->anchor_num > 8 || cfg->class_num < 1 || cfg->max_per_class < 1) {
        set_error("k2y_detect_keras: anchor_num must be 1..8, class_num and max_per_class >= 1");
        return K2Y_ERR_INVALID;
    }
    KerasParams p;
    p.n_layers = cfg->n_layers;
    p.A = cfg->anchor_num;
    p.C = cfg->class_num;
    int off = 0;
    for (int l = 0; l < 3; ++l) {
        p.heads[l] = l < cfg->n_layers ? heads_dev[l] : nullptr;
        p.lh[l] = l < cfg->n_layers ? cfg->layer_h[l] : 0;
        p.lw[l] = l < cfg->n_layers ? cfg->layer_w[l] : 0;
        p.loff[l] = off;
        off += p.lh[l] * p.lw[l] * p.A;
    }
    p.loff[3] = off;
    p.nbox = off;
    p.P = next_pow2(off);
    for (int i = 0; i < 48; ++i) p.anchors[i] = 0.f;
    for (int i = 0; i < cfg->n_layers * cfg->anchor_num * 2; ++i) p.anchors[i] = cfg->anchors[i];
    p.in_h = (float)cfg->in_h;
    p.in_w = (float)cfg->in_w;
    p.obj = cfg->obj_thresh;
    p.iou = cfg->iou_thresh;
    p.maxk = cfg->max_per_class;
    p.image_hw = image_hw_dev;
    p.dets = dets_dev;
    p.counts = counts_dev;
    p.keys_in_smem = p.P <= DET_SMEM_KEYS ? 1 : 0;
    p.keys_global = (unsigned long long *)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    const size_t smem = p.keys_in_smem ? (size_t)DET_WARPS * p.P * sizeof(unsigned long long) : 0;
    static bool attr_set = false;
    if (!attr_set) {
        K2Y_CUDA_CHECK(cudaFuncSetAttribute(detect_keras_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)(DET_WARPS * DET_SMEM_KEYS * sizeof(unsigned long long))));
        attr_set = true;
    }
    dim3 grid((p.C + DET_WARPS - 1) / DET_WARPS, batch);
    p.trace = nullptr;
    const char *tr = getenv("K2Y_DET_TRACE");
    if (tr && tr[0] == '1') {
        K2Y_CUDA_CHECK(cudaMalloc(&p.trace, (size_t)batch * p.C * 4 * sizeof(long long)));
        K2Y_CUDA_CHECK(cudaMemset(p.trace, 0, (size_t)batch * p.C * 4 * sizeof(long long)));
    }
    launch_k(detect_keras_kernel, grid, dim3(DET_WARPS * 32), smem, (cudaStream_t)stream, p);
    K2Y_CUDA_CHECK(cudaGetLastError());
    if (p.trace) {
        K2Y_CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
        std::vector<long long> h((size_t)batch * p.C * 4);
        cudaMemcpy(h.data(), p.trace, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
        cudaFree(p.trace);
        long long mx[4] = {0, 0, 0, 0};
        double sum[4] = {0, 0, 0, 0};
        size_t worst = 0;
        for (size_t i = 0; i < (size_t)batch * p.C; ++i) {
            for (int j = 0; j < 4; ++j) {
                sum[j] += (double)h[i * 4 + j];
                if (h[i * 4 + j] > mx[j]) mx[j] = h[i * 4 + j];
            }
            if (h[i * 4] + h[i * 4 + 1] + h[i * 4 + 2] > h[worst * 4] + h[worst * 4 + 1] + h[worst * 4 + 2]) worst = i;
        }
        const double cnt = (double)batch * p.C;
        fprintf(stderr, "[det-trace] cycles mean (scan %.0f sort %.0f nms %.0f n %.1f) max (scan %lld sort %lld nms %lld n %lld) worst warp: scan %lld sort %lld nms %lld n %lld\n",
                sum[0] / cnt, sum[1] / cnt, sum[2] / cnt, sum[3] / cnt, mx[0], mx[1], mx[2], mx[3], h[worst * 4], h[worst * 4 + 1],
                h[worst * 4 + 2], h[worst * 4 + 3]);
    }
    return K2Y_OK;
}

extern "C" int k2y_region_workspace_bytes(const k2y_region_cfg *cfg, int batch, size_t *bytes) {
    if (!cfg || !bytes || batch <= 0) {
        set_error("k2y_region_workspace_bytes: bad arguments");
        return K2Y_ERR_INVALID;
    }
    const size_t N = (size_t)cfg->layer_w * cfg->layer_h * cfg->anchor_num;
    const size_t P = next_pow2((int)N);
    *bytes = align256((size_t)batch * cfg->classes * N * sizeof(float)) +
             align256((size_t)batch * cfg->classes * P * sizeof(unsigned long long)) +
             align256((size_t)batch * cfg->classes * N * sizeof(int));
    return K2Y_OK;
}

extern "C" int k2y_region_run(const k2y_region_cfg *cfg, const float *in_dev, int batch, float *out_dev,
                              float *probs_dev, float *boxes_dev, void *workspace, size_t workspace_bytes,
                              void *stream) {
    size_t need = 0;
    int rc = k2y_region_workspace_bytes(cfg, batch, &need);
    if (rc != K2Y_OK) return rc;
    if (!in_dev || !probs_dev || !boxes_dev || !workspace || workspace_bytes < need) {
        set_error("k2y_region_run: null pointer or workspace too small (%zu < %zu)", workspace_bytes, need);
        return K2Y_ERR_INVALID;
    }
    if (cfg->anchor_num < 1 || cfg->anchor_num > 8 || cfg->classes < 1) {
        set_error("k2y_region_run: anchor_num must be 1..8 and classes >= 1");
        return K2Y_ERR_INVALID;
    }
    RegionParams p;
    p.in = in_dev;
    p.out = out_dev;
    p.probs = probs_dev;
    p.boxes = (float4 *)boxes_dev;
    p.W = cfg->layer_w;
    p.H = cfg->layer_h;
    p.A = cfg->anchor_num;
    p.C = cfg->classes;
    p.N = p.W * p.H * p.A;
    p.P = next_pow2(p.N);
    for (int i = 0; i < 16; ++i) p.anchors[i] = i < 2 * p.A ? cfg->anchors[i] : 0.f;
    p.thr = cfg->threshold;
    p.nms = cfg->nms_value;
    // correct_region_boxes (region_layer.c:139-164) — same C types and evaluation order.
    {
        const uint32_t net_width = (uint32_t)cfg->net_w, net_height = (uint32_t)cfg->net_h;
        const uint32_t image_width = (uint32_t)cfg->image_w, image_height = (uint32_t)cfg->image_h;
        int new_w = 0, new_h = 0;
        if (((float)net_width / image_width) < ((float)net_height / image_height)) {
            new_w = (int)net_width;
            new_h = (int)((image_height * net_width) / image_width);
        } else {
            new_h = (int)net_height;
            new_w = (int)((image_width * net_height) / image_height);
        }
        p.dx = (net_width - new_w) / 2. / net_width;
        p.dy = (net_height - new_h) / 2. / net_height;
        p.sxw = (double)((float)new_w / net_width);
        p.syh = (double)((float)new_h / net_height);
        p.wscale = (float)net_width / new_w;
        p.hscale = (float)net_height / new_h;
    }
    char *ws = (char *)workspace;
    p.scores = (float *)ws;
    ws += align256((size_t)batch * p.C * p.N * sizeof(float));
    p.keys = (unsigned long long *)ws;
    ws += align256((size_t)batch * p.C * p.P * sizeof(unsigned long long));
    p.kept = (int *)ws;
    region_kernel<<<batch, DET_THREADS, 0, (cudaStream_t)stream>>>(p);
    K2Y_CUDA_CHECK(cudaGetLastError());
    return K2Y_OK;
}

extern "C" int k2y_xywh_to_all(const float *pred_xy_dev, const float *pred_wh_dev, long long n_boxes, int layer_h, int layer_w,
                               int anchor_num, const float *anchors_wh_host, float *xy_dev, float *wh_dev, void *stream) {
    if (!pred_xy_dev || !pred_wh_dev || !xy_dev || !wh_dev || !anchors_wh_host || n_boxes <= 0 || layer_h <= 0 || layer_w <= 0 ||
        anchor_num < 1 || anchor_num > 8 || (n_boxes % ((long long)layer_h * layer_w * anchor_num)) != 0) {
        set_error("k2y_xywh_to_all: bad arguments (n_boxes must be a multiple of layer_h*layer_w*anchor_num, anchor_num 1..8)");
        return K2Y_ERR_INVALID;
    }
    OpsAnchors anc;
    memset(&anc, 0, sizeof(anc));
    for (int i = 0; i < 2 * anchor_num; ++i) anc.wh[i] = anchors_wh_host[i];
    xywh_to_all_kernel<<<(unsigned)((n_boxes + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float2 *>(pred_xy_dev), reinterpret_cast<const float2 *>(pred_wh_dev), reinterpret_cast<float2 *>(xy_dev),
        reinterpret_cast<float2 *>(wh_dev), n_boxes, layer_h, layer_w, anchor_num, anc);
    K2Y_CUDA_CHECK(cudaGetLastError());
    return K2Y_OK;
}

extern "C" int k2y_correct_box(const float *xy_dev, const float *wh_dev, long long n_boxes, float in_h, float in_w, float image_h,
                               float image_w, float *boxes_dev, void *stream) {
    if (!xy_dev || !wh_dev || !boxes_dev || n_boxes <= 0 || !(in_h > 0.f) || !(in_w > 0.f) || !(image_h > 0.f) || !(image_w > 0.f)) {
        set_error("k2y_correct_box: bad arguments");
        return K2Y_ERR_INVALID;
    }
    correct_box_kernel<<<(unsigned)((n_boxes + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float2 *>(xy_dev), reinterpret_cast<const float2 *>(wh_dev), reinterpret_cast<float4 *>(boxes_dev), n_boxes,
        in_h, in_w, image_h, image_w);
    K2Y_CUDA_CHECK(cudaGetLastError());
    return K2Y_OK;
}

extern "C" int k2y_nms_workspace_bytes(int n_boxes, int max_output_size, size_t *bytes) {
    if (!bytes || n_boxes < 0 || max_output_size < 0) {
        set_error("k2y_nms_workspace_bytes: bad arguments");
        return K2Y_ERR_INVALID;
    }
    const size_t P = (size_t)next_pow2(n_boxes);
    const size_t keep = (size_t)(max_output_size < n_boxes ? max_output_size : n_boxes);
    *bytes = align256(P * sizeof(unsigned long long)) + align256(keep * sizeof(float4)) + align256(keep * sizeof(float)) + 256;
    return K2Y_OK;
}

extern "C" int k2y_nms_boxes(const float *boxes_dev, const float *scores_dev, int n_boxes, int max_output_size, float iou_threshold,
                             int32_t *indices_dev, int32_t *count_dev, void *workspace, size_t workspace_bytes, void *stream) {
    size_t need = 0;
    int rc = k2y_nms_workspace_bytes(n_boxes, max_output_size, &need);
    if (rc != K2Y_OK) return rc;
    if (!count_dev || !workspace || workspace_bytes < need || (n_boxes > 0 && (!boxes_dev || !scores_dev)) ||
        (n_boxes > 0 && max_output_size > 0 && !indices_dev) || (((uintptr_t)boxes_dev) & 15) != 0) {
        set_error("k2y_nms_boxes: null / misaligned pointer or workspace too small (%zu < %zu)", workspace_bytes, need);
        return K2Y_ERR_INVALID;
    }
    const int P = next_pow2(n_boxes);
    const size_t keep = (size_t)(max_output_size < n_boxes ? max_output_size : n_boxes);
    char *w = reinterpret_cast<char *>(workspace);
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(w);
    float4 *kept = reinterpret_cast<float4 *>(w + align256((size_t)P * sizeof(unsigned long long)));
    float *kept_area = reinterpret_cast<float *>(reinterpret_cast<char *>(kept) + align256(keep * sizeof(float4)));
    nms_boxes_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4 *>(boxes_dev), scores_dev, n_boxes, P,
                                                         (int)keep, iou_threshold, keys, kept, kept_area, indices_dev, count_dev);
    K2Y_CUDA_CHECK(cudaGetLastError());
    return K2Y_OK;
}
