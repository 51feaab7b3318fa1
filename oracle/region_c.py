"""Oracle for the REGION_C dialect (TEST INFRASTRUCTURE).

Two checkers for the firmware decode path
(/root/reference/yolo3_frame_test_public/region_layer.c):

``RegionLayerRef``   ctypes binding of the *unmodified* reference source compiled into
                     ``oracle/_ref/libregion_layer_ref.so`` (oracle/Makefile) — the real thing.
``region_layer_np``  numpy/float32 restatement of ``region_layer_run`` + ``region_layer_draw_boxes``
                     (region_layer.c:121-137 forward, :139-214 boxes, :216-283 NMS, :385-404 draw),
                     usable where the .so is unavailable, validated against it in
                     tests/test_oracle_region.py.

Input layout for both: float32 CHW planar ``[A][5+C][H][W]`` (entry_index, region_layer.c:84-89).
"""
from __future__ import annotations

import ctypes
import functools
import math
import os
from typing import List, Tuple

import numpy as np

f32 = np.float32
_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(_HERE, "_ref", "libregion_layer_ref.so")


class region_layer_t(ctypes.Structure):
    """Mirror of the struct at region_layer.h:19-39 (field order and types)."""
    _fields_ = [
        ("threshold", ctypes.c_float), ("nms_value", ctypes.c_float),
        ("coords", ctypes.c_uint32), ("anchor_number", ctypes.c_uint32),
        ("anchor", ctypes.POINTER(ctypes.c_float)),
        ("image_width", ctypes.c_uint32), ("image_height", ctypes.c_uint32),
        ("classes", ctypes.c_uint32), ("net_width", ctypes.c_uint32), ("net_height", ctypes.c_uint32),
        ("layer_width", ctypes.c_uint32), ("layer_height", ctypes.c_uint32),
        ("boxes_number", ctypes.c_uint32), ("output_number", ctypes.c_uint32),
        ("boxes", ctypes.c_void_p), ("input", ctypes.POINTER(ctypes.c_float)),
        ("output", ctypes.POINTER(ctypes.c_float)), ("probs_buf", ctypes.POINTER(ctypes.c_float)),
        ("probs", ctypes.POINTER(ctypes.POINTER(ctypes.c_float))),
    ]


DRAW_CB = ctypes.CFUNCTYPE(None, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32,
                           ctypes.c_uint32, ctypes.c_float)


def ref_available() -> bool:
    return os.path.exists(REF_SO)


class RegionLayerRef:
    """Drives the compiled reference exactly as main.c:278-324 does."""

    def __init__(self, width, height, channels, origin_width, origin_height, anchors, threshold, nms_value,
                 image_width=None, image_height=None, lib_path: str = REF_SO):
        self.lib = ctypes.CDLL(lib_path)
        self.lib.region_layer_init.restype = ctypes.c_int
        self.lib.region_layer_init.argtypes = [ctypes.POINTER(region_layer_t)] + [ctypes.c_int] * 5
        self.lib.region_layer_run.argtypes = [ctypes.POINTER(region_layer_t), ctypes.c_void_p]
        self.lib.region_layer_draw_boxes.argtypes = [ctypes.POINTER(region_layer_t), DRAW_CB]
        self.lib.region_layer_deinit.argtypes = [ctypes.POINTER(region_layer_t)]
        self.rl = region_layer_t()
        self._anchors = np.ascontiguousarray(np.asarray(anchors, f32).reshape(-1))
        self.rl.anchor_number = len(self._anchors) // 2
        self.rl.anchor = self._anchors.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
        self.rl.threshold = threshold
        self.rl.nms_value = nms_value
        rc = self.lib.region_layer_init(ctypes.byref(self.rl), width, height, channels, origin_width, origin_height)
        if rc != 0:
            raise MemoryError(f"region_layer_init -> {rc}")
        # region_layer.c:24-25 hard-sets 320x224; callers that want another image size patch the struct.
        if image_width is not None:
            self.rl.image_width = image_width
        if image_height is not None:
            self.rl.image_height = image_height

    def run(self, chw: np.ndarray) -> List[Tuple[int, int, int, int, int, float]]:
        inp = np.ascontiguousarray(chw, f32).reshape(-1)
        assert inp.size == self.rl.output_number
        self.rl.input = inp.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
        self.lib.region_layer_run(ctypes.byref(self.rl), None)
        out = []
        cb = DRAW_CB(lambda x1, y1, x2, y2, c, p: out.append((x1, y1, x2, y2, c, p)))
        self.lib.region_layer_draw_boxes(ctypes.byref(self.rl), cb)
        return out

    def probs(self) -> np.ndarray:
        n, c = self.rl.boxes_number, self.rl.classes
        return np.ctypeslib.as_array(self.rl.probs_buf, shape=(n, c + 1)).copy()

    def boxes(self) -> np.ndarray:
        n = self.rl.boxes_number
        p = ctypes.cast(self.rl.boxes, ctypes.POINTER(ctypes.c_float))
        return np.ctypeslib.as_array(p, shape=(n, 4)).copy()

    def close(self):
        if self.rl.output:
            self.lib.region_layer_deinit(ctypes.byref(self.rl))
            self.rl.output = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---------------------------------------------------------------------------
# numpy restatement
# ---------------------------------------------------------------------------
def _exp2f_table():
    # T[i] = bits(2^(i/32)) - (i << 47), 2^(i/32) correctly rounded to double (glibc's __exp2f_data.tab)
    from decimal import Decimal, getcontext
    getcontext().prec = 60
    t = np.zeros(32, np.uint64)
    for i in range(32):
        v = float(Decimal(2) ** (Decimal(i) / Decimal(32)))
        t[i] = (np.float64(v).view(np.uint64) - np.uint64(i << 47)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    return t


_EXP2F_TAB = _exp2f_table()


def expf_glibc(x) -> np.ndarray:
    """glibc >= 2.28 ``expf`` (sysdeps/ieee754/flt-32/e_expf.c, the algorithm the compiled reference links against),
    restated in float64 numpy: exp(x) = 2^(k/32) * 2^(r/32), k = round(x*32/ln2), cubic in r, one rounding to float32.
    Checked bit for bit against this host's libm in tests/test_oracle_region.py."""
    xf = np.asarray(x, f32)
    xd = xf.astype(np.float64)
    shift = np.float64(6755399441055744.0)  # 0x1.8p52
    z = np.float64(float.fromhex("0x1.71547652b82fep+5")) * xd
    kd = z + shift
    ki = kd.view(np.uint64) if kd.ndim else np.asarray(kd).reshape(1).view(np.uint64).reshape(())
    kd = kd - shift
    r = z - kd
    with np.errstate(over="ignore"):
        t = _EXP2F_TAB[(ki & np.uint64(31)).astype(np.int64)] + (ki << np.uint64(47))
    s = t.view(np.float64) if isinstance(t, np.ndarray) and t.ndim else np.asarray(t, np.uint64).reshape(1).view(np.float64).reshape(())
    c0, c1, c2 = (float.fromhex("0x1.c6af84b912394p-20"), float.fromhex("0x1.ebfce50fac4f3p-13"),
                  float.fromhex("0x1.62e42ff0c52d6p-6"))
    zz = c0 * r + c1
    y = c2 * r + 1.0
    y = zz * (r * r) + y
    with np.errstate(over="ignore", under="ignore"):
        out = (y * s).astype(f32)
    out = np.where(xf > f32(88.72283172607421875), f32(np.inf), out)
    out = np.where(xf < f32(-103.972076416015625), f32(0), out)
    return np.where(np.isnan(xf), xf, out).astype(f32)


def _sigmoid(x):
    return (f32(1) / (f32(1) + expf_glibc(-np.asarray(x, f32)))).astype(f32)


def _iou_center(a, b) -> f32:
    """box_iou on centre-form boxes, float32 op order (region_layer.c:228-254)."""
    def overlap(x1, w1, x2, w2):
        l1 = f32(x1 - f32(w1 / f32(2)))
        l2 = f32(x2 - f32(w2 / f32(2)))
        left = l1 if l1 > l2 else l2
        r1 = f32(x1 + f32(w1 / f32(2)))
        r2 = f32(x2 + f32(w2 / f32(2)))
        right = r1 if r1 < r2 else r2
        return f32(right - left)
    w = overlap(a[0], a[2], b[0], b[2])
    h = overlap(a[1], a[3], b[1], b[3])
    inter = f32(0) if (w < 0 or h < 0) else f32(w * h)
    union = f32(f32(f32(a[2] * a[3]) + f32(b[2] * b[3])) - inter)
    with np.errstate(divide="ignore", invalid="ignore"):
        return f32(inter / union)


def region_layer_np(chw: np.ndarray, width: int, height: int, anchors, threshold: float, nms_value: float,
                    net_width: int, net_height: int, image_width: int = 320, image_height: int = 224):
    """Returns (draw_list, probs[N,C+1], boxes[N,4]) like the compiled reference."""
    anchors = np.asarray(anchors, f32).reshape(-1, 2)
    A = anchors.shape[0]
    wh = width * height
    x = np.asarray(chw, f32).reshape(A, -1, wh)
    C = x.shape[1] - 5
    N = A * wh
    thr = f32(threshold)
    # forward_region_layer (:121-137)
    sx, sy, conf = _sigmoid(x[:, 0]), _sigmoid(x[:, 1]), _sigmoid(x[:, 4])
    cls = x[:, 5:]
    e = expf_glibc((cls - cls.max(axis=1, keepdims=True)).astype(f32))
    ssum = np.zeros((A, wh), f32)
    for j in range(C):  # sequential float32 accumulation, as the C loop does
        ssum = (ssum + e[:, j]).astype(f32)
    soft = (e / ssum[:, None, :]).astype(f32)
    # get_region_boxes (:177-214); index = n*wh + loc
    col = (np.arange(wh) % width).astype(f32)
    row = (np.arange(wh) // width).astype(f32)
    bx = ((col[None] + sx) / f32(width)).astype(f32)
    by = ((row[None] + sy) / f32(height)).astype(f32)
    bw = (expf_glibc(x[:, 2]) * anchors[:, 0:1]).astype(f32)
    bh = (expf_glibc(x[:, 3]) * anchors[:, 1:2]).astype(f32)
    prob = (conf[:, None, :] * soft).astype(f32)  # [A, C, wh]
    probs = np.zeros((N, C + 1), f32)
    pm = np.transpose(prob, (0, 2, 1)).reshape(N, C)
    probs[:, :C] = np.where(pm > thr, pm, f32(0))
    probs[:, C] = np.maximum(pm.max(axis=1), f32(0))
    # correct_region_boxes (:139-164): integer new_w/new_h, double arithmetic for x/y
    if f32(net_width) / f32(image_width) < f32(net_height) / f32(image_height):
        new_w = net_width
        new_h = (image_height * net_width) // image_width
    else:
        new_h = net_height
        new_w = (image_width * net_height) // image_height
    dx = np.float64(np.uint32((net_width - new_w) & 0xFFFFFFFF)) / 2.0 / net_width
    dy = np.float64(np.uint32((net_height - new_h) & 0xFFFFFFFF)) / 2.0 / net_height
    sxw = np.float64(f32(f32(new_w) / f32(net_width)))
    syh = np.float64(f32(f32(new_h) / f32(net_height)))
    boxes = np.empty((N, 4), f32)
    boxes[:, 0] = ((bx.reshape(-1).astype(np.float64) - dx) / sxw).astype(f32)
    boxes[:, 1] = ((by.reshape(-1).astype(np.float64) - dy) / syh).astype(f32)
    boxes[:, 2] = (bw.reshape(-1) * f32(f32(net_width) / f32(new_w))).astype(f32)
    boxes[:, 3] = (bh.reshape(-1) * f32(f32(net_height) / f32(new_h))).astype(f32)
    # do_nms_sort (:256-283); qsort is unstable — ties broken here by ascending index
    for k in range(C):
        order = sorted(range(N), key=lambda i: (-float(probs[i, k]), i))
        for ii, i in enumerate(order):
            if probs[i, k] == 0:
                continue
            a = boxes[i]
            for j in order[ii + 1:]:
                if probs[j, k] == 0:
                    continue  # (the C code re-zeroes; same result)
                if _iou_center(a, boxes[j]) > f32(nms_value):
                    probs[j, k] = 0
    # region_layer_draw_boxes (:385-404)
    out = []
    iw, ih = f32(image_width), f32(image_height)
    for i in range(N):
        c = int(np.argmax(probs[i, :C]))  # first maximum, like max_index (:285-296)
        p = probs[i, c]
        if p > thr:
            b = boxes[i]
            half_w = f32(f32(b[2] * iw) / f32(2))
            half_h = f32(f32(b[3] * ih) / f32(2))
            vals = (f32(f32(b[0] * iw) - half_w), f32(f32(b[1] * ih) - half_h),
                    f32(f32(b[0] * iw) + half_w), f32(f32(b[1] * ih) + half_h))
            out.append(tuple(_f32_to_u32(v) for v in vals) + (c, float(p)))
    return out, probs, boxes


def _f32_to_u32(v) -> int:
    """x86-64 gcc float->uint32: cvttss2si to 64-bit then truncation (negative wraps)."""
    v = float(v)
    if math.isnan(v) or abs(v) >= 2.0 ** 63:
        return 0
    return int(v) & 0xFFFFFFFF


def nhwc_to_chw(head_nhwc: np.ndarray, anchor_num: int) -> np.ndarray:
    """[h, w, A*(5+C)] Keras head -> the firmware's [A][5+C][h][w] planar layout."""
    h, w, ch = head_nhwc.shape
    return np.ascontiguousarray(head_nhwc.reshape(h, w, anchor_num, ch // anchor_num).transpose(2, 3, 0, 1))
