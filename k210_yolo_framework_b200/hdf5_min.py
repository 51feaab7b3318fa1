"""Minimal pure-Python HDF5 reader — enough for Keras ``save_model`` / ``save_weights`` files.

The reference loads checkpoints with ``yolo_model_warpper.load_weights(ckpt)``
(/root/reference/keras_inference.py:80), i.e. Keras-HDF5 written by
``keras.models.save_model`` (/root/reference/keras_train.py:105,109).  h5py is not
available in this image, so this module walks the on-disk format directly.

Supported subset (everything the Keras writer of that era emits for weight files):
superblock v0/v1, version-1 object headers (+ continuation blocks), "old style" groups
(symbol-table message -> v1 B-tree -> SNOD nodes + local heap), contiguous or compact
datasets of little-endian IEEE floats / fixed-point ints, fixed-length string and
variable-length string attributes (global heap).  Chunked / filtered datasets raise.

Public API:  ``File(path)`` -> ``.root`` Group; ``Group.keys()``, ``Group[name]``,
``Group.attrs``; ``Dataset.read() -> np.ndarray``; ``File.visit_datasets()``.
"""
from __future__ import annotations

import struct
from typing import Dict, Iterator, List, Tuple, Union

import numpy as np

_UNDEF = 0xFFFFFFFFFFFFFFFF


class HDF5FormatError(ValueError):
    pass


class _Datatype:
    __slots__ = ("cls", "size", "np_dtype", "is_vlen_str", "str_pad")

    def __init__(self, buf: bytes, off: int):
        cv = buf[off]
        self.cls = cv & 0x0F
        bits0 = buf[off + 1]
        self.size = struct.unpack_from("<I", buf, off + 4)[0]
        self.np_dtype = None
        self.is_vlen_str = False
        self.str_pad = 0
        if self.cls == 0:  # fixed point
            big = bits0 & 1
            signed = (bits0 >> 3) & 1
            self.np_dtype = np.dtype(("i" if signed else "u") + str(self.size)).newbyteorder(">" if big else "<")
        elif self.cls == 1:  # float
            big = bits0 & 1
            self.np_dtype = np.dtype("f" + str(self.size)).newbyteorder(">" if big else "<")
        elif self.cls == 3:  # fixed-length string
            self.np_dtype = np.dtype("S" + str(self.size))
            self.str_pad = bits0 & 0x0F
        elif self.cls == 9:  # variable length
            vtype = bits0 & 0x0F
            self.is_vlen_str = vtype == 1
            if not self.is_vlen_str:
                raise HDF5FormatError("variable-length sequences are not supported")
        else:
            raise HDF5FormatError(f"unsupported datatype class {self.cls}")


def _parse_dataspace(buf: bytes, off: int) -> Tuple[int, ...]:
    ver = buf[off]
    rank = buf[off + 1]
    if ver == 1:
        p = off + 8
    elif ver == 2:
        p = off + 4
    else:
        raise HDF5FormatError(f"dataspace version {ver}")
    return tuple(struct.unpack_from("<" + "Q" * rank, buf, p)) if rank else ()


class Dataset:
    def __init__(self, f: "File", name: str, msgs):
        self._f = f
        self.name = name
        self.shape: Tuple[int, ...] = ()
        self._dt = None
        self._layout = None
        self.attrs: Dict[str, object] = {}
        for mtype, off, size in msgs:
            if mtype == 0x0001:
                self.shape = _parse_dataspace(f.buf, off)
            elif mtype == 0x0003:
                self._dt = _Datatype(f.buf, off)
            elif mtype == 0x0008:
                self._layout = (off, size)
            elif mtype == 0x000B:
                raise HDF5FormatError(f"{name}: filtered datasets are not supported")
            elif mtype == 0x000C:
                k, v = f._parse_attribute(off)
                self.attrs[k] = v

    @property
    def dtype(self):
        return self._dt.np_dtype

    def read(self) -> np.ndarray:
        buf = self._f.buf
        off, _ = self._layout
        ver = buf[off]
        if ver != 3:
            raise HDF5FormatError(f"{self.name}: data layout version {ver}")
        lclass = buf[off + 1]
        count = int(np.prod(self.shape)) if self.shape else 1
        nbytes = count * self._dt.size
        if lclass == 1:
            addr, sz = struct.unpack_from("<QQ", buf, off + 2)
            if addr == _UNDEF:
                return np.zeros(self.shape, self._dt.np_dtype)
            raw = buf[addr:addr + nbytes]
        elif lclass == 0:
            sz = struct.unpack_from("<H", buf, off + 2)[0]
            raw = buf[off + 4:off + 4 + sz][:nbytes]
        else:
            raise HDF5FormatError(f"{self.name}: chunked layout is not supported")
        if len(raw) != nbytes:
            raise HDF5FormatError(f"{self.name}: truncated data")
        return np.frombuffer(raw, self._dt.np_dtype, count).reshape(self.shape).copy()


class Group:
    def __init__(self, f: "File", name: str, msgs):
        self._f = f
        self.name = name
        self._links: Dict[str, int] = {}
        self.attrs: Dict[str, object] = {}
        for mtype, off, size in msgs:
            if mtype == 0x0011:
                btree, heap = struct.unpack_from("<QQ", f.buf, off)
                self._links.update(f._read_group_links(btree, heap))
            elif mtype == 0x000C:
                k, v = f._parse_attribute(off)
                self.attrs[k] = v
            elif mtype in (0x0002, 0x0006):
                raise HDF5FormatError("new-style (link-message) groups are not supported")

    def keys(self) -> List[str]:
        return list(self._links.keys())

    def __contains__(self, k: str) -> bool:
        return k.split("/")[0] in self._links if "/" in k else k in self._links

    def __getitem__(self, path: str) -> Union["Group", Dataset]:
        node: Union[Group, Dataset] = self
        for part in [p for p in path.split("/") if p]:
            if not isinstance(node, Group) or part not in node._links:
                raise KeyError(f"{path!r} (missing {part!r} in {node.name!r})")
            child = (node.name.rstrip("/") + "/" + part)
            node = node._f._load_object(node._links[part], child)
        return node


class File:
    def __init__(self, path: str):
        with open(path, "rb") as fh:
            self.buf = fh.read()
        self.path = str(path)
        b = self.buf
        if b[:8] != b"\x89HDF\r\n\x1a\n":
            raise HDF5FormatError(f"{path}: not an HDF5 file")
        ver = b[8]
        if ver not in (0, 1):
            raise HDF5FormatError(f"superblock version {ver} is not supported")
        if b[13] != 8 or b[14] != 8:
            raise HDF5FormatError("only 8-byte offsets/lengths are supported")
        p = 24 + (4 if ver == 1 else 0)
        base = struct.unpack_from("<Q", b, p)[0]
        if base != 0:
            raise HDF5FormatError("non-zero base address")
        root_entry = p + 32
        _, ohdr, ctype, _ = struct.unpack_from("<QQII", b, root_entry)
        self._cache: Dict[int, object] = {}
        self.root: Group = self._load_object(ohdr, "/")

    # ---- object headers -------------------------------------------------
    def _messages(self, addr: int):
        b = self.buf
        if b[addr] != 1:
            raise HDF5FormatError(f"object header version {b[addr]} at {addr:#x}")
        nmsg = struct.unpack_from("<H", b, addr + 2)[0]
        hsize = struct.unpack_from("<I", b, addr + 8)[0]
        blocks = [(addr + 16, hsize)]
        out = []
        while blocks and len(out) < nmsg:
            p, left = blocks.pop(0)
            end = p + left
            while p + 8 <= end and len(out) < nmsg:
                mtype, msize, _flags = struct.unpack_from("<HHB", b, p)
                body = p + 8
                if mtype == 0x0010:
                    coff, clen = struct.unpack_from("<QQ", b, body)
                    blocks.append((coff, clen))
                out.append((mtype, body, msize))
                p = body + msize
        return out

    def _load_object(self, addr: int, name: str):
        if addr in self._cache:
            return self._cache[addr]
        msgs = self._messages(addr)
        types = {m[0] for m in msgs}
        obj = Dataset(self, name, msgs) if 0x0008 in types else Group(self, name, msgs)
        self._cache[addr] = obj
        return obj

    # ---- old-style groups -----------------------------------------------
    def _heap_string(self, heap_addr: int, off: int) -> str:
        b = self.buf
        if b[heap_addr:heap_addr + 4] != b"HEAP":
            raise HDF5FormatError("bad local heap signature")
        data = struct.unpack_from("<Q", b, heap_addr + 24)[0]
        s = data + off
        e = b.index(b"\x00", s)
        return b[s:e].decode("utf-8")

    def _read_group_links(self, btree: int, heap: int) -> Dict[str, int]:
        b = self.buf
        out: Dict[str, int] = {}
        sig = b[btree:btree + 4]
        if sig == b"TREE":
            level = b[btree + 5]
            n = struct.unpack_from("<H", b, btree + 6)[0]
            p = btree + 24
            for i in range(n):
                child = struct.unpack_from("<Q", b, p + 8 + i * 16)[0]
                out.update(self._read_group_links(child, heap))
        elif sig == b"SNOD":
            n = struct.unpack_from("<H", b, btree + 6)[0]
            p = btree + 8
            for i in range(n):
                noff, ohdr = struct.unpack_from("<QQ", b, p + i * 40)
                out[self._heap_string(heap, noff)] = ohdr
        else:
            raise HDF5FormatError(f"bad group node signature {sig!r}")
        return out

    # ---- attributes -----------------------------------------------------
    def _global_heap_object(self, coll: int, index: int) -> bytes:
        b = self.buf
        if b[coll:coll + 4] != b"GCOL":
            raise HDF5FormatError("bad global heap signature")
        csize = struct.unpack_from("<Q", b, coll + 8)[0]
        p = coll + 16
        end = coll + csize
        while p + 16 <= end:
            idx, _ref, _r, size = struct.unpack_from("<HHIQ", b, p)
            if idx == 0:
                break
            if idx == index:
                return b[p + 16:p + 16 + size]
            p += 16 + ((size + 7) & ~7)
        raise HDF5FormatError("global heap object not found")

    def _parse_attribute(self, off: int):
        b = self.buf
        ver = b[off]
        nsize, tsize, ssize = struct.unpack_from("<HHH", b, off + 2)
        p = off + 8
        if ver == 3:
            p += 1
        pad = (lambda n: (n + 7) & ~7) if ver == 1 else (lambda n: n)
        name = b[p:p + nsize].split(b"\x00")[0].decode("utf-8")
        p += pad(nsize)
        dt = _Datatype(b, p)
        p += pad(tsize)
        shape = _parse_dataspace(b, p) if ssize >= 4 else ()
        p += pad(ssize)
        count = int(np.prod(shape)) if shape else 1
        if dt.is_vlen_str:
            vals = []
            for i in range(count):
                ln, coll, idx = struct.unpack_from("<IQI", b, p + i * 16)
                vals.append(self._global_heap_object(coll, idx)[:ln].decode("utf-8") if ln else "")
            val = vals[0] if not shape else np.array(vals, dtype=object).reshape(shape)
        else:
            arr = np.frombuffer(b[p:p + count * dt.size], dt.np_dtype, count).reshape(shape).copy()
            if dt.cls == 3:
                arr = np.char.rstrip(arr, b"\x00") if arr.shape else arr
            val = arr if shape else arr.reshape(()).item()
        return name, val

    # ---- helpers ----------------------------------------------------------
    def visit_datasets(self, group: Group = None) -> Iterator[Tuple[str, Dataset]]:
        group = group or self.root
        for k in group.keys():
            node = group[k]
            if isinstance(node, Group):
                yield from self.visit_datasets(node)
            else:
                yield node.name, node


def load_keras_weights(path: str) -> Dict[str, Dict[str, np.ndarray]]:
    """Read every layer's weights from a Keras HDF5 file.

    Handles both containers the reference produces/consumes: full-model files
    (``/model_weights/<layer>/<layer>/<var>:0``, keras_train.py:105) and weights-only
    files (``/<layer>/<layer>/<var>:0``).  Returns ``{layer: {var_name: array}}`` with
    the ``:0`` suffix stripped (e.g. ``kernel``, ``depthwise_kernel``, ``gamma``,
    ``beta``, ``moving_mean``, ``moving_variance``, ``bias``).
    """
    f = File(path)
    root = f.root["model_weights"] if "model_weights" in f.root.keys() else f.root
    out: Dict[str, Dict[str, np.ndarray]] = {}
    for layer in root.keys():
        g = root[layer]
        if not isinstance(g, Group):
            continue
        vars_: Dict[str, np.ndarray] = {}
        for name, ds in f.visit_datasets(g):
            var = name.rsplit("/", 1)[-1].split(":")[0]
            vars_[var] = ds.read()
        if vars_:
            out[layer] = vars_
    return out
