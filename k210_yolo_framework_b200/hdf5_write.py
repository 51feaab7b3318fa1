"""Minimal pure-Python HDF5 WRITER for Keras weight files, and the BN-folded "frozen" export.

The reference saves checkpoints with ``keras.models.save_model`` / ``model.save_weights``
(/root/reference/keras_train.py:105-109) and freezes a model for deployment in keras_freeze.py:17-19.  h5py is not
available in this image; this module emits the on-disk format directly, the mirror image of ``hdf5_min`` (the reader):

  superblock v0 -> root group (v1 object header, symbol-table message -> v1 B-tree -> SNOD leaves + local heap)
  -> ``[model_weights/]<layer>/<layer>/<var>:0`` contiguous little-endian float32 datasets, with the attributes Keras'
  ``load_weights`` walks (``layer_names``, ``weight_names``, ``backend``, ``keras_version``) as fixed-length strings.

``save_keras_weights(path, weights)``   {layer: {var: ndarray}} -> Keras-HDF5 (round-trips through ``load_keras_weights``)
``fold_batchnorm(weights, expected)``   BN folded into the conv it follows: kernel' = kernel * g/sqrt(var+eps) per output
                                        channel, bias' = beta - mean * g/sqrt(var+eps) (+ bias * ...) — the packed form
``export_frozen(path, weights, expected)``  the folded weights as Keras-HDF5 (``.h5``) or ``.npz``
"""
from __future__ import annotations

import struct
from typing import Dict, List, Tuple

import numpy as np

Weights = Dict[str, Dict[str, np.ndarray]]
_UNDEF = 0xFFFFFFFFFFFFFFFF
_LEAF_K, _INT_K = 32, 16          # group B-tree parameters written into the superblock (64 links per leaf, 32 leaves per node)
BN_EPS = 1e-3


def _pad8(b: bytes) -> bytes:
    return b + b"\x00" * (-len(b) % 8)


class _Writer:
    def __init__(self):
        self.buf = bytearray(b"\x00" * 96)   # superblock placeholder

    def alloc(self, data: bytes) -> int:
        while len(self.buf) % 8:
            self.buf.append(0)
        addr = len(self.buf)
        self.buf += data
        return addr

    # ---- messages -------------------------------------------------------
    @staticmethod
    def _msg(mtype: int, body: bytes, flags: int = 0) -> bytes:
        body = _pad8(body)
        return struct.pack("<HHB3x", mtype, len(body), flags) + body

    @staticmethod
    def _dataspace(shape: Tuple[int, ...]) -> bytes:
        return struct.pack("<BBB5x", 1, len(shape), 0) + b"".join(struct.pack("<Q", int(d)) for d in shape)

    @staticmethod
    def _dtype_f32() -> bytes:
        # class 1 (float) v1; little-endian, implied-msb mantissa normalisation, sign at bit 31; IEEE binary32 field layout
        return struct.pack("<BBBBI", 0x11, 0x20, 31, 0, 4) + struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)

    @staticmethod
    def _dtype_str(n: int) -> bytes:
        return struct.pack("<BBBBI", 0x13, 0x00, 0, 0, n)   # class 3 (string) v1, null-terminated, ASCII

    def _attr(self, name: str, dtype: bytes, shape: Tuple[int, ...], data: bytes) -> bytes:
        nm = name.encode() + b"\x00"
        space = self._dataspace(shape)
        body = struct.pack("<BBHHH", 1, 0, len(nm), len(dtype), len(space)) + _pad8(nm) + _pad8(dtype) + _pad8(space) + data
        return self._msg(0x000C, body)

    def attr_strings(self, name: str, values: List[str]) -> bytes:
        raw = [v.encode() for v in values]
        n = max([len(r) for r in raw] + [1])
        return self._attr(name, self._dtype_str(n), (len(raw),), b"".join(r.ljust(n, b"\x00") for r in raw))

    def attr_string(self, name: str, value: str) -> bytes:
        raw = value.encode()
        return self._attr(name, self._dtype_str(max(len(raw), 1)), (), raw or b"\x00")

    def object_header(self, messages: List[bytes]) -> int:
        body = b"".join(messages)
        hdr = struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(body))
        return self.alloc(hdr + body)

    # ---- objects --------------------------------------------------------
    def dataset(self, arr: np.ndarray) -> int:
        arr = np.ascontiguousarray(arr, dtype="<f4")
        data_addr = self.alloc(arr.tobytes()) if arr.size else _UNDEF
        msgs = [self._msg(0x0001, self._dataspace(arr.shape)),
                self._msg(0x0003, self._dtype_f32(), flags=1),
                self._msg(0x0005, struct.pack("<BBBB", 2, 2, 0, 0)),                      # fill value v2: late alloc, undefined
                self._msg(0x0008, struct.pack("<BBQQ", 3, 1, data_addr, arr.size * 4))]  # layout v3, contiguous
        return self.object_header(msgs)

    def group(self, links: Dict[str, int], attrs: List[bytes] = ()) -> Tuple[int, int, int]:
        """Old-style group over {name: object header address}.  Returns (object header, b-tree, heap) addresses."""
        names = sorted(links, key=lambda s: s.encode())
        if len(names) > 2 * _LEAF_K * 2 * _INT_K:
            raise ValueError(f"group with {len(names)} links exceeds the single-level B-tree this writer emits")
        # local heap: offset 0 holds the empty string (the B-tree's first key), names follow 8-byte aligned
        heap = bytearray(b"\x00" * 8)
        noff = {}
        for nm in names:
            noff[nm] = len(heap)
            heap += _pad8(nm.encode() + b"\x00")
        heap_data = self.alloc(bytes(heap))
        heap_addr = self.alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap), 1, heap_data))   # free-list head 1 = none
        per = 2 * _LEAF_K
        leaves = [names[i:i + per] for i in range(0, len(names), per)] or [[]]
        children, keys = [], [0]
        for leaf in leaves:
            body = bytearray(b"SNOD" + struct.pack("<BBH", 1, 0, len(leaf)))
            for nm in leaf:
                body += struct.pack("<QQII16x", noff[nm], links[nm], 0, 0)
            body += b"\x00" * (40 * (per - len(leaf)))
            children.append(self.alloc(bytes(body)))
            keys.append(noff[leaf[-1]] if leaf else 0)
        node = bytearray(b"TREE" + struct.pack("<BBHQQ", 0, 0, len(children), _UNDEF, _UNDEF))
        for i, ch in enumerate(children):
            node += struct.pack("<QQ", keys[i], ch)
        node += struct.pack("<Q", keys[len(children)])
        node += b"\x00" * (24 + (2 * _INT_K + 1) * 8 + 2 * _INT_K * 8 - len(node))
        btree = self.alloc(bytes(node))
        ohdr = self.object_header([self._msg(0x0011, struct.pack("<QQ", btree, heap_addr))] + list(attrs))
        return ohdr, btree, heap_addr

    def finish(self, root: Tuple[int, int, int]) -> bytes:
        ohdr, btree, heap = root
        while len(self.buf) % 8:
            self.buf.append(0)
        sb = (b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBBBBBB", 0, 0, 0, 0, 0, 8, 8, 0) + struct.pack("<HHI", _LEAF_K, _INT_K, 0) +
              struct.pack("<QQQQ", 0, _UNDEF, len(self.buf), _UNDEF) + struct.pack("<QQII", 0, ohdr, 1, 0) + struct.pack("<QQ", btree, heap))
        assert len(sb) == 96
        self.buf[:96] = sb
        return bytes(self.buf)


_VAR_ORDER = ["kernel", "depthwise_kernel", "bias", "gamma", "beta", "moving_mean", "moving_variance"]


def save_keras_weights(path: str, weights: Weights, full_model: bool = True, keras_version: str = "2.2.4-tf",
                       backend: str = "tensorflow") -> None:
    """Writes ``{layer: {var: array}}`` as a Keras HDF5 file: ``/model_weights/<layer>/<layer>/<var>:0`` (``full_model``, what
    ``save_model`` produces and keras_inference.py:80 loads) or ``/<layer>/<layer>/<var>:0`` (``save_weights``)."""
    w = _Writer()
    layer_links = {}
    for layer, vars_ in weights.items():
        order = [v for v in _VAR_ORDER if v in vars_] + [v for v in vars_ if v not in _VAR_ORDER]
        inner = {f"{v}:0": w.dataset(np.asarray(vars_[v])) for v in order}
        inner_grp = w.group(inner)[0]
        attrs = [w.attr_strings("weight_names", [f"{layer}/{v}:0" for v in order])]
        layer_links[layer] = w.group({layer: inner_grp}, attrs)[0]
    top_attrs = [w.attr_strings("layer_names", list(weights)), w.attr_string("backend", backend),
                 w.attr_string("keras_version", keras_version)]
    if full_model:
        mw = w.group(layer_links, top_attrs)[0]
        root = w.group({"model_weights": mw}, [w.attr_string("backend", backend), w.attr_string("keras_version", keras_version)])
    else:
        root = w.group(layer_links, top_attrs)
    with open(path, "wb") as fh:
        fh.write(w.finish(root))


def fold_batchnorm(weights: Weights, expected: Dict[str, Dict[str, tuple]], bn_of: Dict[str, str]) -> Weights:
    """BatchNormalization (eps 1e-3, inference statistics) folded into the conv in front of it.

    ``bn_of`` maps conv layer -> its BN layer ('' if none), ``YoloEngine.bn_pairs()``.  Returns {conv: {kernel |
    depthwise_kernel, bias}} — every layer becomes ``conv(x, kernel') + bias'``; the BN layers disappear."""
    out: Weights = {}
    for layer, bn in bn_of.items():
        src = weights[layer]
        kname = "depthwise_kernel" if "depthwise_kernel" in src else "kernel"
        k = np.asarray(src[kname], np.float32)
        cout = k.shape[2] if kname == "depthwise_kernel" else k.shape[3]
        scale = np.ones(cout, np.float32)
        shift = np.zeros(cout, np.float32)
        if bn:
            b = weights[bn]
            scale = (np.asarray(b["gamma"], np.float32) / np.sqrt(np.asarray(b["moving_variance"], np.float32) + np.float32(BN_EPS))).astype(np.float32)
            shift = (np.asarray(b["beta"], np.float32) - np.asarray(b["moving_mean"], np.float32) * scale).astype(np.float32)
        if "bias" in src:
            shift = (shift + np.asarray(src["bias"], np.float32) * scale).astype(np.float32)
        kf = (k * scale[None, None, :, None]) if kname == "depthwise_kernel" else (k * scale[None, None, None, :])
        out[layer] = {kname: kf.astype(np.float32), "bias": shift}
    return out


def export_frozen(path: str, weights: Weights, expected: Dict[str, Dict[str, tuple]], bn_of: Dict[str, str]) -> Weights:
    """keras_freeze.py's role without TensorFlow: the BN-folded, inference-only weights of the model, as ``.npz``
    (``layer/var`` keys) or Keras-HDF5.  Returns the folded dict."""
    folded = fold_batchnorm(weights, expected, bn_of)
    if str(path).endswith(".npz"):
        np.savez(path, **{f"{layer}/{var}": arr for layer, vars_ in folded.items() for var, arr in vars_.items()})
    else:
        save_keras_weights(path, folded, full_model=True)
    return folded
