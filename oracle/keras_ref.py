"""Oracle restatement of the reference's four graph builders (TEST INFRASTRUCTURE).

Follows, layer by layer:
  * ``yolo_mobilev1``  /root/reference/models/yolonet.py:12-46  +  MobileNet
    /root/reference/models/keras_mobilenet.py:215-229, _conv_block :291-356,
    _depthwise_conv_block :359-436
  * ``yolo_mobilev2``  models/yolonet.py:49-104 + MobileNetV2
    models/keras_mobilenet_v2.py:311-382, _inverted_res_block :426-485, _make_divisible :118-126
  * ``tiny_yolo``      models/yolonet.py:107-158
  * ``yolo``           models/yolonet.py:161-229 (darknet_body, resblock_body, make_last_layers)
  * DarknetConv2D / DarknetConv2D_BN_Leaky  models/yolonet.py:244-260

The arithmetic the reference delegates to TensorFlow 1.14 (Conv2D, DepthwiseConv2dNative,
FusedBatchNorm(inference), LeakyRelu, Relu/Relu6, ResizeNearestNeighbor, ConcatV2, MaxPool,
Add) is restated with torch CPU ops in fp32 (the "TF-CPU stand-in") or fp64 (ground truth
for error measurement).  Weights arrive as ``{keras_layer_name: {var: ndarray}}`` — exactly
what the Keras HDF5 file holds — and layers are auto-named the way Keras names them when
one model is built per process (conv2d, conv2d_1, batch_normalization, ...).

Deviation from the reference, on purpose: the wrapper ``Reshape`` targets are derived from
the actual grid (H/32, H/16, H/8) instead of the hard-coded (7,10)/(14,20)/(13,13)...
(models/yolonet.py:40-41,98-99,140-141,175-177), which the reference cannot build for
any other input size.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Weights = Dict[str, Dict[str, np.ndarray]]

BN_EPS = 1e-3  # keras BatchNormalization default epsilon (h5 model_config: 0.001)


class _Tape:
    """Executes Keras-layer semantics eagerly and hands out Keras auto-names."""

    def __init__(self, weights: Weights, dtype: torch.dtype, record: bool = False, cache: dict = None):
        self.w = weights
        self.dtype = dtype
        self.cache = cache  # optional {key: converted tensor} reused across calls (CPU-baseline timing)
        self._count: Dict[str, int] = {}
        self.record = record
        self.acts: Dict[str, torch.Tensor] = {}

    def auto(self, base: str) -> str:
        n = self._count.get(base, 0)
        self._count[base] = n + 1
        return base if n == 0 else f"{base}_{n}"

    def _t(self, a: np.ndarray) -> torch.Tensor:
        return torch.from_numpy(np.ascontiguousarray(a)).to(self.dtype)

    def _rec(self, name: str, x: torch.Tensor) -> torch.Tensor:
        if self.record:
            self.acts[name] = x
        return x

    # -- layers (NCHW tensors) --------------------------------------------
    def zero_pad(self, x, top, bottom, left, right):
        return F.pad(x, (left, right, top, bottom))

    def _cached(self, key, make):
        if self.cache is None:
            return make()
        if key not in self.cache:
            self.cache[key] = make()
        return self.cache[key]

    def conv(self, x, name, stride=1, padding="same", use_bias=False):
        kh = self.w[name]["kernel"].shape[0]  # HWIO
        wt = self._cached((name, "w"), lambda: self._t(self.w[name]["kernel"]).permute(3, 2, 0, 1).contiguous())
        b = self._cached((name, "b"), lambda: self._t(self.w[name]["bias"])) if use_bias else None
        if padding == "same":
            assert stride == 1, "the reference never uses SAME with stride 2 for convs"
            p = kh // 2
        else:
            p = 0
        return self._rec(name, F.conv2d(x, wt, b, stride=stride, padding=p))

    def dwconv(self, x, name, stride=1, padding="same"):
        c = self.w[name]["depthwise_kernel"].shape[2]  # (3,3,C,1)
        wt = self._cached((name, "w"), lambda: self._t(self.w[name]["depthwise_kernel"]).permute(2, 3, 0, 1).contiguous())  # (C,1,3,3)
        p = 1 if padding == "same" else 0
        if padding == "same":
            assert stride == 1
        return self._rec(name, F.conv2d(x, wt, None, stride=stride, padding=p, groups=c))

    def bn(self, x, name):
        def make():
            p = self.w[name]
            g, b = self._t(p["gamma"]), self._t(p["beta"])
            m, v = self._t(p["moving_mean"]), self._t(p["moving_variance"])
            return m[None, :, None, None], (torch.rsqrt(v + BN_EPS) * g)[None, :, None, None], b[None, :, None, None]
        m, inv, b = self._cached((name, "bn"), make)
        return self._rec(name, (x - m) * inv + b)

    def leaky(self, x, alpha, name=None):
        a = float(np.float32(alpha))
        y = torch.where(x >= 0, x, x * a)
        return self._rec(name, y) if name else y

    def relu(self, x, max_value=None, name=None):
        y = torch.clamp(x, min=0.0) if max_value is None else torch.clamp(x, min=0.0, max=float(max_value))
        return self._rec(name, y) if name else y

    def upsample2(self, x):
        return x.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)

    def maxpool_same(self, x, stride):
        # 2x2 window, TF 'SAME': pads at bottom/right only, with -inf.
        h, w = x.shape[2], x.shape[3]
        oh, ow = -(-h // stride), -(-w // stride)
        ph = max((oh - 1) * stride + 2 - h, 0)
        pw = max((ow - 1) * stride + 2 - w, 0)
        if ph or pw:
            x = F.pad(x, (0, pw, 0, ph), value=float("-inf"))
        return F.max_pool2d(x, 2, stride)

    # -- composites from models/yolonet.py:244-260 ---------------------------
    def darknet_conv_bn_leaky(self, x, ksize, stride=1):
        cname = self.auto("conv2d")
        bname = self.auto("batch_normalization")
        self.auto("leaky_re_lu")
        x = self.conv(x, cname, stride=stride, padding="valid" if stride == 2 else "same", use_bias=False)
        assert self.w[cname]["kernel"].shape[0] == ksize
        x = self.bn(x, bname)
        return self.leaky(x, 0.1, name=bname + "/leaky")

    def darknet_conv(self, x):
        cname = self.auto("conv2d")
        return self.conv(x, cname, stride=1, padding="same", use_bias=True)


def _to_nchw(x: np.ndarray, dtype) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(x)).to(dtype).permute(0, 3, 1, 2).contiguous()


def _to_nhwc(y: torch.Tensor) -> np.ndarray:
    return y.permute(0, 2, 3, 1).contiguous().numpy()


# ---------------------------------------------------------------------------
# backbones
# ---------------------------------------------------------------------------
def _mobilenet_v1(t: _Tape, x, alpha):
    """keras_mobilenet.py:215-229 — returns (conv_pw_11_relu, conv_pw_13_relu)."""
    x = t.zero_pad(x, 1, 1, 1, 1)
    x = t.conv(x, "conv1", stride=2, padding="valid")
    x = t.bn(x, "conv1_bn")
    x = t.leaky(x, 0.3, name="conv1_relu")
    plan = [(40 if alpha == 1.0 else 64, 1), (128, 2), (128, 1), (256, 2), (256, 1), (512, 2),
            (512, 1), (512, 1), (512, 1), (512, 1), (512, 1), (1024, 2), (1024, 1)]
    x1 = None
    for i, (f, s) in enumerate(plan, start=1):
        if s == 2:
            x = t.zero_pad(x, 1, 1, 1, 1)
        x = t.dwconv(x, f"conv_dw_{i}", stride=s, padding="same" if s == 1 else "valid")
        x = t.bn(x, f"conv_dw_{i}_bn")
        x = t.relu(x, name=f"conv_dw_{i}_relu")
        x = t.conv(x, f"conv_pw_{i}", stride=1, padding="same")
        assert t.w[f"conv_pw_{i}"]["kernel"].shape[3] == int(f * alpha)
        x = t.bn(x, f"conv_pw_{i}_bn")
        x = t.leaky(x, 0.3, name=f"conv_pw_{i}_relu")
        if i == 11:
            x1 = x
    return x1, x


def make_divisible(v, divisor, min_value=None):
    """keras_mobilenet_v2.py:118-126."""
    if min_value is None:
        min_value = divisor
    new_v = max(min_value, int(v + divisor / 2) // divisor * divisor)
    if new_v < 0.9 * v:
        new_v += divisor
    return new_v


def _mobilenet_v2(t: _Tape, x, alpha):
    """keras_mobilenet_v2.py:311-382 — returns (block_13_expand_relu, out_relu)."""
    x = t.zero_pad(x, 1, 1, 1, 1)
    x = t.conv(x, "Conv1", stride=2, padding="valid")  # hard-wired 32 filters (:313)
    x = t.bn(x, "bn_Conv1")
    x = t.relu(x, 6.0, name="Conv1_relu")
    blocks = [  # (filters, stride, expansion, block_id, expand_channel)
        (16, 1, 1, 0, None),
        (24, 2, 6, 1, 48 if alpha > 0.6 else None), (24, 1, 6, 2, 124 if alpha > 0.6 else None),
        (32, 2, 6, 3, None), (32, 1, 6, 4, None), (32, 1, 6, 5, None),
        (64, 2, 6, 6, None), (64, 1, 6, 7, None), (64, 1, 6, 8, None), (64, 1, 6, 9, None),
        (96, 1, 6, 10, None), (96, 1, 6, 11, None), (96, 1, 6, 12, None),
        (160, 2, 6, 13, None), (160, 1, 6, 14, None), (160, 1, 6, 15, None),
        (320, 1, 6, 16, None)]
    x1 = None
    for filters, stride, expansion, bid, expand_channel in blocks:
        inputs = x
        in_ch = x.shape[1]
        pw_filters = make_divisible(int(filters * alpha), 8)
        prefix = f"block_{bid}_"
        if bid:
            x = t.conv(x, prefix + "expand", 1, "same")
            assert x.shape[1] == (expand_channel if expand_channel else expansion * in_ch)
            x = t.bn(x, prefix + "expand_BN")
            x = t.relu(x, 6.0, name=prefix + "expand_relu")
            if bid == 13:
                x1 = x
        else:
            prefix = "expanded_conv_"
        if stride == 2:
            x = t.zero_pad(x, 1, 1, 1, 1)
        x = t.dwconv(x, prefix + "depthwise", stride, "same" if stride == 1 else "valid")
        x = t.bn(x, prefix + "depthwise_BN")
        x = t.relu(x, 6.0, name=prefix + "depthwise_relu")
        x = t.conv(x, prefix + "project", 1, "same")
        assert x.shape[1] == pw_filters
        x = t.bn(x, prefix + "project_BN")
        if in_ch == pw_filters and stride == 1:
            x = t._rec(prefix + "add", inputs + x)
    x = t.conv(x, "Conv_1", 1, "same")
    x = t.bn(x, "Conv_1_bn")
    x = t.relu(x, 6.0, name="out_relu")
    return x1, x


def _two_scale_heads(t: _Tape, x1, x2):
    """The head wiring shared by yolo_mobilev1/v2/tiny_yolo (yolonet.py:27-38, 87-96, 126-138)."""
    y1 = t.darknet_conv_bn_leaky(x2, 3)
    y1 = t.darknet_conv(y1)
    x2 = t.darknet_conv_bn_leaky(x2, 1)
    x2 = t.upsample2(x2)
    y2 = torch.cat([x2, x1], dim=1)
    y2 = t.darknet_conv_bn_leaky(y2, 3)
    y2 = t.darknet_conv(y2)
    return [y1, y2]


def _tiny_yolo_body(t: _Tape, x):
    """yolonet.py:110-124."""
    for i in range(4):
        x = t.darknet_conv_bn_leaky(x, 3)
        x = t.maxpool_same(x, 2)
    x1 = t.darknet_conv_bn_leaky(x, 3)
    x = t.maxpool_same(x1, 2)
    x = t.darknet_conv_bn_leaky(x, 3)
    x = t.maxpool_same(x, 1)
    x = t.darknet_conv_bn_leaky(x, 3)
    x2 = t.darknet_conv_bn_leaky(x, 1)
    return x1, x2


def _darknet53(t: _Tape, x):
    """yolonet.py:194-215 — returns (skip92, skip152, out) = 256-, 512-, 1024-channel stage outputs."""
    x = t.darknet_conv_bn_leaky(x, 3)
    outs = []
    for nf, nb in ((64, 1), (128, 2), (256, 8), (512, 8), (1024, 4)):
        x = t.zero_pad(x, 1, 0, 1, 0)  # ((1,0),(1,0)): top/left only (:197)
        x = t.darknet_conv_bn_leaky(x, 3, stride=2)
        for _ in range(nb):
            y = t.darknet_conv_bn_leaky(x, 1)
            y = t.darknet_conv_bn_leaky(y, 3)
            x = x + y
        outs.append(x)
    return outs[2], outs[3], outs[4]


def _make_last_layers(t: _Tape, x):
    """yolonet.py:218-229."""
    for k in (1, 3, 1, 3, 1):
        x = t.darknet_conv_bn_leaky(x, k)
    y = t.darknet_conv_bn_leaky(x, 3)
    y = t.darknet_conv(y)
    return x, y


# ---------------------------------------------------------------------------
# public entry: forward(model_def, weights, x_nhwc) -> list of [N,h,w,A*(5+C)]
# ---------------------------------------------------------------------------
def forward(model_def: str, weights: Weights, x_nhwc: np.ndarray, alpha: float = 1.0,
            dtype: torch.dtype = torch.float32, record: bool = False, cache: dict = None, channels_last: bool = False):
    """Run the plain ``yolo_model`` graph; returns NHWC head tensors ``[N,h_l,w_l,A*(5+C)]``.

    With ``record=True`` also returns ``{layer_name: NCHW tensor}`` of intermediate outputs.
    """
    t = _Tape(weights, dtype, record, cache)
    x = _to_nchw(x_nhwc, dtype)
    if channels_last:
        x = x.contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        if model_def == "yolo_mobilev1":
            x1, x2 = _mobilenet_v1(t, x, alpha)
            ys = _two_scale_heads(t, x1, x2)
        elif model_def == "yolo_mobilev2":
            x1, x2 = _mobilenet_v2(t, x, alpha)
            ys = _two_scale_heads(t, x1, x2)
        elif model_def == "tiny_yolo":
            x1, x2 = _tiny_yolo_body(t, x)
            ys = _two_scale_heads(t, x1, x2)
        elif model_def == "yolo":
            s92, s152, out = _darknet53(t, x)
            xx, y1 = _make_last_layers(t, out)
            xx = t.upsample2(t.darknet_conv_bn_leaky(xx, 1))
            xx = torch.cat([xx, s152], dim=1)
            xx, y2 = _make_last_layers(t, xx)
            xx = t.upsample2(t.darknet_conv_bn_leaky(xx, 1))
            xx = torch.cat([xx, s92], dim=1)
            xx, y3 = _make_last_layers(t, xx)
            ys = [y1, y2, y3]
        else:
            raise ValueError(model_def)
    heads = [_to_nhwc(y.to(torch.float32) if dtype == torch.float32 else y) for y in ys]
    if record:
        return heads, t.acts
    return heads


def forward_wrapper(model_def, weights, x_nhwc, anchor_num, class_num, alpha=1.0, dtype=torch.float32):
    """``yolo_model_warpper.predict``: heads reshaped to ``[N,h,w,A,5+C]`` (yolonet.py:40-44)."""
    heads = forward(model_def, weights, x_nhwc, alpha, dtype)
    return [h.reshape(h.shape[0], h.shape[1], h.shape[2], anchor_num, 5 + class_num) for h in heads]
