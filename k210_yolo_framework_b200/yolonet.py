"""Host-side mirror of the reference's model-builder API (/root/reference/models/yolonet.py).

    network = eval(model_def)                                   # keras_inference.py:77
    yolo_model, yolo_model_warpper = network([H, W, 3], anchor_num, class_num, alpha=depth_multiplier)
    yolo_model_warpper.load_weights(ckpt)                       # :80
    y_pred = yolo_model_warpper.predict(img[None])              # :88

Same four builder names and call signature (yolonet.py:12, :49, :107, :161); the returned objects
offer ``load_weights`` / ``predict`` like the two ``keras.Model`` views, but execute the layer graph
through the C-ABI (``k2y_net_*``) on hand-written sm_100a kernels.  PyTorch tensors hold the device
storage only.  Differences from the reference, on purpose:
  * no side files are read at construction (the reference loads un-shipped ``data/*_base_*.h5`` /
    ``data/*yolo_weights.h5`` inside the builders, yolonet.py:16-21,76-81,146,182);
  * wrapper outputs are reshaped to the real grid ``[N, H/32·2^l, W/32·2^l, A, 5+C]`` instead of the
    hard-coded (7,10)/(14,20)/(13,13)... targets (yolonet.py:40-41,98-99,140-141,175-177).
"""
from __future__ import annotations

import ctypes
import re
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import K2YError, check, lib
from .hdf5_min import load_keras_weights

Weights = Dict[str, Dict[str, np.ndarray]]
DEFAULT_MAX_BATCH = 32


class YoloEngine:
    """One native network (k2y_net) plus the torch storage bound to it."""

    def __init__(self, model_def: str, input_shape: Sequence[int], anchor_num: int, class_num: int, alpha: float,
                 max_batch: int = DEFAULT_MAX_BATCH, device: Optional[int] = None):
        if len(input_shape) != 3 or input_shape[2] != 3:
            raise ValueError(f"input_shape must be [H, W, 3], got {list(input_shape)}")
        self.model_def = model_def
        self.in_h, self.in_w = int(input_shape[0]), int(input_shape[1])
        self.anchor_num, self.class_num, self.alpha = int(anchor_num), int(class_num), float(alpha)
        self.max_batch = int(max_batch)
        self.device_index = torch.cuda.current_device() if (device is None and torch.cuda.is_available()) else int(device or 0)
        h = ctypes.c_void_p()
        check(lib.k2y_net_create(model_def.encode(), self.in_h, self.in_w, self.alpha, self.anchor_num, self.class_num,
                                 self.max_batch, self.device_index, ctypes.byref(h)))
        self._h = h
        n = ctypes.c_int()
        check(lib.k2y_net_num_outputs(h, ctypes.byref(n)))
        self.out_shapes = []
        for l in range(n.value):
            a, b, c = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
            check(lib.k2y_net_output_shape(h, l, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
            self.out_shapes.append((a.value, b.value, c.value))
        self._bound = False
        self._finalized = False
        self._x = None
        self._heads: List[torch.Tensor] = []
        self._ws = None
        self._keep_all = False

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib.k2y_net_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # -- graph introspection -------------------------------------------------
    def layers(self) -> List[_lib.LayerInfo]:
        n = ctypes.c_int()
        check(lib.k2y_net_num_layers(self._h, ctypes.byref(n)))
        out = []
        for i in range(n.value):
            info = _lib.LayerInfo()
            check(lib.k2y_net_layer_info(self._h, i, ctypes.byref(info)))
            out.append(info)
        return out

    def expected_variables(self) -> Dict[str, Dict[str, tuple]]:
        """{keras_layer: {var: shape}} the graph needs — the file format of keras save_model."""
        exp: Dict[str, Dict[str, tuple]] = {}
        for L in self.layers():
            name, bn = L.name.decode(), L.bn_name.decode()
            if L.kind == 1:
                exp[name] = {"depthwise_kernel": (3, 3, L.cin, 1)}
            else:
                exp[name] = {"kernel": (L.kh, L.kw, L.cin, L.cout)}
                if L.has_bias:
                    exp[name]["bias"] = (L.cout,)
            if bn:
                exp[bn] = {v: (L.cout,) for v in ("gamma", "beta", "moving_mean", "moving_variance")}
        return exp

    def bn_pairs(self) -> Dict[str, str]:
        """{conv layer: the BatchNormalization layer folded behind it ('' if none)}, in creation order."""
        return {L.name.decode(): L.bn_name.decode() for L in self.layers()}

    # -- weights -------------------------------------------------------------
    def set_weights(self, weights: Weights) -> None:
        _lib.require_cuda()
        exp = self.expected_variables()
        weights = _rename_auto_named(weights, exp)
        for layer, vars_ in exp.items():
            if layer not in weights:
                raise ValueError(f"weights for layer '{layer}' not found ({self.model_def}); file has {len(weights)} layers")
            for var, shape in vars_.items():
                if var not in weights[layer]:
                    raise ValueError(f"variable '{layer}/{var}' not found")
                arr = np.ascontiguousarray(weights[layer][var], dtype=np.float32)
                if tuple(arr.shape) != tuple(shape):
                    raise ValueError(f"shape mismatch for {layer}/{var}: file {arr.shape}, graph {shape}")
                dims = (ctypes.c_int64 * arr.ndim)(*arr.shape)
                check(lib.k2y_net_set_weight(self._h, layer.encode(), var.encode(), arr.ctypes.data, dims, arr.ndim))
        with torch.cuda.device(self.device_index):
            check(lib.k2y_net_finalize(self._h))
        self._finalized = True

    def load_weights(self, path: str) -> None:
        """Keras HDF5 (save_model / save_weights).  Extension: a ``.npz`` with ``layer/var`` keys (the
        container tests/golden uses for the reference's trained model) is accepted too."""
        path = str(path)
        if path.endswith(".npz"):
            self.set_weights(load_npz_weights(path))
        else:
            self.set_weights(load_keras_weights(path))

    # -- execution -----------------------------------------------------------
    def set_math(self, mode: int) -> None:
        check(lib.k2y_net_set_math(self._h, int(mode)))

    def get_math(self) -> int:
        m = ctypes.c_int()
        check(lib.k2y_net_get_math(self._h, ctypes.byref(m)))
        return m.value

    def set_use_graph(self, flag: bool) -> None:
        check(lib.k2y_net_set_use_graph(self._h, int(bool(flag))))

    def set_keep_all(self, flag: bool) -> None:
        check(lib.k2y_net_set_keep_all(self._h, int(bool(flag))))
        self._keep_all = bool(flag)
        self._bound = False

    def launches_per_run(self) -> int:
        n = ctypes.c_int()
        check(lib.k2y_net_launches_per_run(self._h, ctypes.byref(n)))
        return n.value

    def _bind(self) -> None:
        if self._bound:
            return
        _lib.require_cuda()
        dev = torch.device("cuda", self.device_index)
        nbytes = ctypes.c_size_t()
        check(lib.k2y_net_workspace_bytes(self._h, ctypes.byref(nbytes)))
        self._ws = torch.empty(max(nbytes.value, 256), dtype=torch.uint8, device=dev)
        self._x = torch.empty((self.max_batch, self.in_h, self.in_w, 3), dtype=torch.float32, device=dev)
        self._heads = [torch.empty((self.max_batch,) + s, dtype=torch.float32, device=dev) for s in self.out_shapes]
        ptrs = (ctypes.c_void_p * len(self._heads))(*[t.data_ptr() for t in self._heads])
        check(lib.k2y_net_bind(self._h, self._ws.data_ptr(), self._ws.numel(), self._x.data_ptr(), ptrs, len(self._heads)))
        self._bound = True

    def enable_u8_input(self) -> torch.Tensor:
        """Switch the network to the uint8 front end (k2y_net_bind_u8): returns the bound device buffer
        [max_batch,H,W,3] uint8 (letterboxed RGB).  ``img / np.max(img)`` (tools/utils.py:405) then happens on the GPU."""
        self._bind()
        if getattr(self, "_x_u8", None) is None:
            dev = torch.device("cuda", self.device_index)
            self._x_u8 = torch.empty((self.max_batch, self.in_h, self.in_w, 3), dtype=torch.uint8, device=dev)
            if getattr(self, "_img_max", None) is None:
                self._img_max = torch.zeros((self.max_batch,), dtype=torch.int32, device=dev)
        check(lib.k2y_net_bind_u8(self._h, self._x_u8.data_ptr(), self._img_max.data_ptr()))
        self._u8_on = True
        self._ext_input = None
        return self._x_u8

    def disable_u8_input(self) -> None:
        if getattr(self, "_u8_on", False):
            check(lib.k2y_net_bind_u8(self._h, None, None))
            self._u8_on = False

    def bind_input(self, buf: torch.Tensor) -> None:
        """Points the network input at `buf` (CUDA [max_batch,H,W,3], float32 or uint8) instead of the engine's own
        buffer — the zero-copy form of double-buffered ingest (one CUDA graph is kept per input buffer)."""
        self._bind()
        if (not buf.is_cuda or not buf.is_contiguous() or tuple(buf.shape) != (self.max_batch, self.in_h, self.in_w, 3)
                or buf.dtype not in (torch.float32, torch.uint8)):
            raise ValueError(f"expected contiguous CUDA [{self.max_batch},{self.in_h},{self.in_w},3] float32 or uint8")
        if buf.dtype == torch.uint8:
            if getattr(self, "_img_max", None) is None:
                self._img_max = torch.zeros((self.max_batch,), dtype=torch.int32, device=buf.device)
            check(lib.k2y_net_bind_u8(self._h, buf.data_ptr(), self._img_max.data_ptr()))
            self._u8_on = True
        else:
            self.disable_u8_input()
            check(lib.k2y_net_bind_input(self._h, buf.data_ptr()))
        self._ext_input = buf  # keep it alive

    def set_sm_limit(self, sms: int) -> None:
        """SM budget of this net's persistent tensor-core kernels (0 = whole device); see k2y_net_set_sm_limit."""
        check(lib.k2y_net_set_sm_limit(self._h, int(sms)))

    def bind_heads(self, heads: List[torch.Tensor]) -> None:
        """Points the head outputs at another set of CUDA float32 buffers ``[max_batch, h, w, c]`` (one CUDA graph is kept per
        set): with two sets, decode/NMS of batch i on another stream can overlap the convolutions of batch i+1."""
        self._bind()
        if len(heads) != len(self.out_shapes) or any(
                (not t.is_cuda) or t.dtype != torch.float32 or (not t.is_contiguous()) or tuple(t.shape) != (self.max_batch,) + tuple(sh)
                for t, sh in zip(heads, self.out_shapes)):
            raise ValueError(f"expected contiguous CUDA float32 heads of shapes {[(self.max_batch,) + tuple(sh) for sh in self.out_shapes]}")
        ptrs = (ctypes.c_void_p * len(heads))(*[t.data_ptr() for t in heads])
        check(lib.k2y_net_bind_heads(self._h, ptrs, len(heads)))
        self._heads = list(heads)

    def new_head_set(self) -> List[torch.Tensor]:
        """A second set of head buffers, shaped like the engine's own."""
        self._bind()
        return [torch.empty_like(t) for t in self._heads]

    def unbind_input(self) -> None:
        """Back to the engine's own float32 input buffer."""
        if getattr(self, "_ext_input", None) is not None:
            self.disable_u8_input()
            check(lib.k2y_net_bind_input(self._h, self._x.data_ptr()))
            self._ext_input = None

    def predict_device_u8(self, x_u8: torch.Tensor) -> List[torch.Tensor]:
        """x_u8: CUDA uint8 [N,H,W,3].  Same result as predict_device(x_u8 / max(x_u8) per image)."""
        if x_u8.dim() != 4 or tuple(x_u8.shape[1:]) != (self.in_h, self.in_w, 3) or x_u8.dtype != torch.uint8 or not x_u8.is_cuda:
            raise ValueError(f"expected CUDA uint8 [N,{self.in_h},{self.in_w},3]")
        n = x_u8.shape[0]
        if n > self.max_batch:
            raise ValueError(f"batch {n} > max_batch {self.max_batch}")
        buf = self.enable_u8_input()
        if x_u8.data_ptr() != buf.data_ptr():
            buf[:n].copy_(x_u8, non_blocking=True)
        return self.run(n)

    @property
    def input_buffer(self) -> torch.Tensor:
        """The bound device input [max_batch, H, W, 3] f32 — write into it to skip the staging copy."""
        self._bind()
        return self._x

    @property
    def head_buffers(self) -> List[torch.Tensor]:
        self._bind()
        return self._heads

    def run(self, batch: int, stream: Optional[torch.cuda.Stream] = None) -> List[torch.Tensor]:
        """Runs the bound buffers; returns views ``heads[l][:batch]`` (asynchronous)."""
        if not self._finalized:
            raise K2YError("weights not loaded: call load_weights()/set_weights() first")
        self._bind()
        st = stream if stream is not None else torch.cuda.current_stream(self.device_index)
        check(lib.k2y_net_run(self._h, int(batch), ctypes.c_void_p(st.cuda_stream)))
        return [t[:batch] for t in self._heads]

    def predict_device(self, x: torch.Tensor) -> List[torch.Tensor]:
        """x: CUDA float32 [N,H,W,3] (N <= max_batch).  Returns device head tensors [N,h,w,A*(5+C)]."""
        if x.dim() != 4 or tuple(x.shape[1:]) != (self.in_h, self.in_w, 3) or x.dtype != torch.float32 or not x.is_cuda:
            raise ValueError(f"expected CUDA float32 [N,{self.in_h},{self.in_w},3], got {tuple(x.shape)} {x.dtype} {x.device}")
        n = x.shape[0]
        if n > self.max_batch:
            raise ValueError(f"batch {n} > max_batch {self.max_batch}")
        self._bind()
        self.unbind_input()
        self.disable_u8_input()
        if x.data_ptr() != self._x.data_ptr():
            self._x[:n].copy_(x, non_blocking=True)
        return self.run(n)

    def predict_host(self, x: np.ndarray) -> List[np.ndarray]:
        """keras ``predict``: host float array [N,H,W,3] -> list of host arrays, via k2y_net_predict_host."""
        if not self._finalized:
            raise K2YError("weights not loaded: call load_weights()/set_weights() first")
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.ndim != 4 or x.shape[1:] != (self.in_h, self.in_w, 3):
            raise ValueError(f"expected [N,{self.in_h},{self.in_w},3], got {x.shape}")
        self._bind()
        self.unbind_input()
        self.disable_u8_input()  # predict() feeds float32 pixels (k2y_net_predict_host leaves the uint8 front end as well)
        outs = [np.empty((x.shape[0],) + s, np.float32) for s in self.out_shapes]
        st = torch.cuda.current_stream(self.device_index)
        for s in range(0, x.shape[0], self.max_batch):
            n = min(self.max_batch, x.shape[0] - s)
            ptrs = (ctypes.c_void_p * len(outs))(*[o[s:s + n].ctypes.data for o in outs])
            check(lib.k2y_net_predict_host(self._h, x[s:s + n].ctypes.data, n, ptrs, ctypes.c_void_p(st.cuda_stream)))
        return outs

    def profile(self, batch: int) -> List[dict]:
        """One event-instrumented pass: [{name, ms, flops, bytes}] per launch (flops/bytes for `batch` images)."""
        self._bind()
        cnt = ctypes.c_int()
        check(lib.k2y_net_schedule_len(self._h, ctypes.byref(cnt)))
        n = cnt.value
        ms = (ctypes.c_float * n)()
        st = torch.cuda.current_stream(self.device_index)
        check(lib.k2y_net_profile(self._h, int(batch), ctypes.c_void_p(st.cuda_stream), ms, n))
        out = []
        for i in range(n):
            name = ctypes.create_string_buffer(64)
            fl, by = ctypes.c_double(), ctypes.c_double()
            check(lib.k2y_net_launch_info(self._h, i, name, 64, ctypes.byref(fl), ctypes.byref(by)))
            out.append({"name": name.value.decode(), "ms": float(ms[i]), "flops": fl.value * batch, "bytes": by.value * batch})
        return out

    def read_layer(self, name: str, batch: int) -> np.ndarray:
        """Parity hook: output of conv layer `name` from the last run (needs set_keep_all(True), no graph)."""
        buf = np.empty(batch * self.in_h * self.in_w * 32, np.float32)  # >= any layer output
        h, w, c = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        check(lib.k2y_net_read_layer(self._h, name.encode(), batch, buf.ctypes.data, buf.size, ctypes.byref(h),
                                     ctypes.byref(w), ctypes.byref(c)))
        return buf[:batch * h.value * w.value * c.value].reshape(batch, h.value, w.value, c.value).copy()


def load_npz_weights(path: str) -> Weights:
    out: Weights = {}
    with np.load(path) as z:
        for key in z.files:
            layer, var = key.rsplit("/", 1)
            out.setdefault(layer, {})[var] = z[key]
    return out


def _rename_auto_named(weights: Weights, expected: Dict[str, Dict[str, tuple]]) -> Weights:
    """Keras auto-names (conv2d_7, batch_normalization_3...) depend on how many layers the saving process
    had created before; when the file's names are offset, re-map them by creation order per base name."""
    if all(k in weights for k in expected):
        return weights

    def split(name):
        m = re.fullmatch(r"(.*?)(?:_(\d+))?", name)
        return (m.group(1), int(m.group(2))) if m.group(2) is not None else (name, 0)

    out = dict(weights)
    for base in ("conv2d", "batch_normalization"):
        want = sorted((k for k in expected if split(k)[0] == base), key=lambda k: split(k)[1])
        have = sorted((k for k in weights if split(k)[0] == base), key=lambda k: split(k)[1])
        if len(want) == len(have):
            for w, h in zip(want, have):
                out[w] = weights[h]
    return out


class YoloModel:
    """A ``keras.Model``-like view: ``wrapped=False`` is ``yolo_model`` (heads [N,h,w,A*(5+C)]),
    ``wrapped=True`` is ``yolo_model_warpper`` (heads reshaped to [N,h,w,A,5+C])."""

    def __init__(self, engine: YoloEngine, wrapped: bool):
        self.engine = engine
        self.wrapped = wrapped

    def load_weights(self, filepath) -> None:
        self.engine.load_weights(str(filepath))

    def set_weights_dict(self, weights: Weights) -> None:
        self.engine.set_weights(weights)

    def _shape(self, arr):
        if not self.wrapped:
            return arr
        a, c = self.engine.anchor_num, self.engine.class_num
        return arr.reshape(arr.shape[0], arr.shape[1], arr.shape[2], a, 5 + c)

    def predict(self, x) -> List[np.ndarray]:
        if isinstance(x, torch.Tensor):
            x = x.detach().cpu().numpy()
        return [self._shape(o) for o in self.engine.predict_host(np.asarray(x))]

    def predict_device(self, x: torch.Tensor) -> List[torch.Tensor]:
        return [self._shape(o) for o in self.engine.predict_device(x)]

    def predict_device_u8(self, x_u8: torch.Tensor) -> List[torch.Tensor]:
        """CUDA uint8 [N,H,W,3] letterboxed RGB in; ``img / np.max(img)`` (tools/utils.py:405) runs on the GPU."""
        return [self._shape(o) for o in self.engine.predict_device_u8(x_u8)]

    @property
    def output_shapes(self):
        a, c = self.engine.anchor_num, self.engine.class_num
        return [((None, h, w, a, 5 + c) if self.wrapped else (None, h, w, ch)) for h, w, ch in self.engine.out_shapes]


def _network(model_def, input_shape, anchor_num, class_num, kwargs):
    alpha = float(kwargs.get("alpha", 1.0))
    eng = YoloEngine(model_def, input_shape, anchor_num, class_num, alpha,
                     max_batch=kwargs.get("max_batch", DEFAULT_MAX_BATCH), device=kwargs.get("device"))
    return YoloModel(eng, False), YoloModel(eng, True)


def yolo_mobilev1(input_shape: list, anchor_num: int, class_num: int, **kwargs):
    """models/yolonet.py:12-46."""
    return _network("yolo_mobilev1", input_shape, anchor_num, class_num, kwargs)


def yolo_mobilev2(input_shape: list, anchor_num: int, class_num: int, **kwargs):
    """models/yolonet.py:49-104."""
    return _network("yolo_mobilev2", input_shape, anchor_num, class_num, kwargs)


def tiny_yolo(input_shape, anchor_num, class_num, **kwargs):
    """models/yolonet.py:107-158."""
    return _network("tiny_yolo", input_shape, anchor_num, class_num, kwargs)


def yolo(input_shape, anchor_num, class_num, **kwargs):
    """models/yolonet.py:161-191."""
    return _network("yolo", input_shape, anchor_num, class_num, kwargs)


__all__ = ["yolo_mobilev1", "yolo_mobilev2", "tiny_yolo", "yolo", "YoloEngine", "YoloModel"]
