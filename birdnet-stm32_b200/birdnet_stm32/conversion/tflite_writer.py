"""Dependency-free writer for `.tflite` flatbuffers (TFLite schema v3), the inverse of `tflite_reader`.

The reference produces its deployable model with `tf.lite.TFLiteConverter` (`conversion/quantize.py:111-168`);
TensorFlow cannot run here, so graphs quantised by `conversion/ptq.py` are serialised by this module instead.
Only the tables the engine consumes are written (Model, OperatorCode, SubGraph, Tensor, QuantizationParameters,
Operator with its builtin-options table, Buffer), with the schema's own field ids, so the files are ordinary
TFLite models: `read_tflite(write_tflite(g))` reproduces `g`, and the shipped checkpoint survives a
read -> write -> read round trip with an identical exported blob (tests/test_ptq.py).

Flatbuffer mechanics: every `uoffset` points forward, so each object is written before its children and the
offsets are patched once the children are placed; vtables sit immediately in front of their tables.
"""

from __future__ import annotations

import struct

import numpy as np

from birdnet_stm32.conversion.tflite_reader import BUILTIN, FUSED_ACT, PADDING, TENSOR_DTYPES, Graph

_CODE = {v: k for k, v in BUILTIN.items()}
_TTYPE = {np.dtype(v): k for k, v in TENSOR_DTYPES.items()}
_ACT = {v: k for k, v in FUSED_ACT.items()}
_PAD = {v: k for k, v in PADDING.items()}

# BuiltinOptions union tags (schema v3)
_OPT_TAG = {
    "CONV_2D": 1, "DEPTHWISE_CONV_2D": 2, "AVERAGE_POOL_2D": 5, "FULLY_CONNECTED": 8, "SOFTMAX": 9, "CONCATENATION": 10,
    "ADD": 11, "RESHAPE": 17, "MUL": 21, "PAD": 22, "TRANSPOSE": 26, "MEAN": 27, "SUM": 27, "REDUCE_MAX": 27, "SUB": 28,
    "DIV": 29, "STRIDED_SLICE": 32, "DEQUANTIZE": 38, "SHAPE": 55, "PACK": 59, "LOGISTIC": 0, "FILL": 68, "QUANTIZE": 76,
}


class _Table:
    def __init__(self):
        self.fields: dict[int, tuple] = {}

    def scalar(self, idx: int, fmt: str, value, default=0):
        if value != default:
            self.fields[idx] = ("s", fmt, value)
        return self

    def ref(self, idx: int, obj):
        if obj is not None:
            self.fields[idx] = ("r", obj)
        return self


class _Vector:
    """kind: numpy dtype for scalar vectors, 'tables' for vectors of tables, 'str' for strings."""

    def __init__(self, kind, items):
        self.kind, self.items = kind, items


class _Builder:
    def __init__(self):
        self.b = bytearray(8)

    def _align(self, n: int, bias: int = 0):
        while (len(self.b) + bias) % n:
            self.b.append(0)

    def emit(self, obj) -> int:
        if isinstance(obj, _Table):
            return self._table(obj)
        return self._vector(obj)

    def _table(self, t: _Table) -> int:
        nf = (max(t.fields) + 1) if t.fields else 0
        vt_size = 4 + 2 * nf
        self._align(4, vt_size)
        vt_pos = len(self.b)
        self.b += bytes(vt_size)
        tpos = len(self.b)
        self.b += struct.pack("<i", tpos - vt_pos)
        offs = [0] * nf
        patches = []
        for idx in sorted(t.fields):
            f = t.fields[idx]
            if f[0] == "s":
                size = struct.calcsize(f[1])
                self._align(size)
                offs[idx] = len(self.b) - tpos
                self.b += struct.pack("<" + f[1], f[2])
            else:
                self._align(4)
                offs[idx] = len(self.b) - tpos
                patches.append((len(self.b), f[1]))
                self.b += bytes(4)
        self._align(4)
        struct.pack_into("<HH", self.b, vt_pos, vt_size, len(self.b) - tpos)
        for i, o in enumerate(offs):
            struct.pack_into("<H", self.b, vt_pos + 4 + 2 * i, o)
        for pos, child in patches:
            cpos = self.emit(child)
            struct.pack_into("<I", self.b, pos, cpos - pos)
        return tpos

    def _vector(self, v: _Vector) -> int:
        if v.kind == "str":
            raw = v.items.encode("utf-8")
            self._align(4)
            pos = len(self.b)
            self.b += struct.pack("<I", len(raw)) + raw + b"\0"
            return pos
        if v.kind == "tables":
            self._align(4)
            pos = len(self.b)
            self.b += struct.pack("<I", len(v.items))
            slots = len(self.b)
            self.b += bytes(4 * len(v.items))
            for i, child in enumerate(v.items):
                cpos = self.emit(child)
                struct.pack_into("<I", self.b, slots + 4 * i, cpos - (slots + 4 * i))
            return pos
        if v.kind == "bytes":                       # buffer payloads: 16-byte aligned like the converter writes them
            self._align(16, 4)
            pos = len(self.b)
            self.b += struct.pack("<I", len(v.items)) + bytes(v.items)
            return pos
        arr = np.ascontiguousarray(v.items, dtype=v.kind)
        self._align(max(4, arr.dtype.itemsize), 4)
        pos = len(self.b)
        self.b += struct.pack("<I", arr.size) + arr.tobytes()
        return pos

    def finish(self, root: _Table, ident: bytes) -> bytes:
        rpos = self.emit(root)
        struct.pack_into("<I", self.b, 0, rpos)
        self.b[4:8] = ident
        return bytes(self.b)


def _options_table(kind: str, o: dict):
    t = _Table()
    if kind == "CONV_2D":
        t.scalar(0, "b", _PAD[o.get("padding", "SAME")]).scalar(1, "i", o.get("stride_w", 1)).scalar(2, "i", o.get("stride_h", 1))
        t.scalar(3, "b", _ACT[o.get("act", "NONE")]).scalar(4, "i", o.get("dil_w", 1), 1).scalar(5, "i", o.get("dil_h", 1), 1)
    elif kind == "DEPTHWISE_CONV_2D":
        t.scalar(0, "b", _PAD[o.get("padding", "SAME")]).scalar(1, "i", o.get("stride_w", 1)).scalar(2, "i", o.get("stride_h", 1))
        t.scalar(3, "i", o.get("depth_multiplier", 1)).scalar(4, "b", _ACT[o.get("act", "NONE")])
        t.scalar(5, "i", o.get("dil_w", 1), 1).scalar(6, "i", o.get("dil_h", 1), 1)
    elif kind == "FULLY_CONNECTED":
        t.scalar(0, "b", _ACT[o.get("act", "NONE")]).scalar(1, "b", o.get("weights_format", 0)).scalar(2, "B", int(o.get("keep_num_dims", False)))
    elif kind in ("ADD", "MUL", "SUB", "DIV"):
        t.scalar(0, "b", _ACT[o.get("act", "NONE")])
    elif kind in ("MEAN", "SUM", "REDUCE_MAX"):
        t.scalar(0, "B", int(o.get("keep_dims", False)))
    elif kind == "AVERAGE_POOL_2D":
        t.scalar(0, "b", _PAD[o.get("padding", "SAME")]).scalar(1, "i", o.get("stride_w", 1)).scalar(2, "i", o.get("stride_h", 1))
        t.scalar(3, "i", o.get("filter_w", 1)).scalar(4, "i", o.get("filter_h", 1)).scalar(5, "b", _ACT[o.get("act", "NONE")])
    elif kind == "SOFTMAX":
        t.scalar(0, "f", float(o.get("beta", 1.0)), 0.0)
    elif kind == "CONCATENATION":
        t.scalar(0, "i", o.get("axis", 0)).scalar(1, "b", _ACT[o.get("act", "NONE")])
    elif kind == "STRIDED_SLICE":
        for i, k in enumerate(("begin_mask", "end_mask", "ellipsis_mask", "new_axis_mask", "shrink_axis_mask")):
            t.scalar(i, "i", o.get(k, 0))
    elif kind == "PACK":
        t.scalar(0, "i", o.get("values_count", 0)).scalar(1, "i", o.get("axis", 0))
    elif kind == "RESHAPE":
        if o.get("new_shape"):
            t.ref(0, _Vector(np.int32, list(o["new_shape"])))
    elif not o:
        return None
    return t


def write_tflite(g: Graph) -> bytes:
    """Serialise a :class:`Graph` (as produced by `read_tflite` or `conversion.ptq`) to `.tflite` bytes."""
    kinds = []
    for op in g.ops:
        if op.kind not in _CODE:
            raise ValueError(f"op {op.index}: cannot serialise operator {op.kind}")
        if (op.kind, op.version) not in kinds:
            kinds.append((op.kind, op.version))
    opcodes = []
    for kind, ver in kinds:
        code = _CODE[kind]
        oc = _Table().scalar(0, "b", min(code, 127)).scalar(2, "i", ver, 1).scalar(3, "i", code)
        opcodes.append(oc)

    buffers = [_Table()]                                   # buffer 0: the empty sentinel
    tensors = []
    for t in g.tensors:
        tt = _Table()
        tt.ref(0, _Vector(np.int32, list(t.shape)))
        tt.scalar(1, "b", _TTYPE[np.dtype(t.dtype)])
        if t.data is not None:
            raw = np.ascontiguousarray(t.data, dtype=t.dtype).tobytes()
            buffers.append(_Table().ref(0, _Vector("bytes", raw)))
            tt.scalar(2, "I", len(buffers) - 1)
        tt.ref(3, _Vector("str", t.name))
        if t.scale.size:
            q = _Table()
            q.ref(2, _Vector(np.float32, t.scale)).ref(3, _Vector(np.int64, t.zero_point)).scalar(6, "i", int(t.quantized_dimension))
            tt.ref(4, q)
        if tuple(t.shape_signature) != tuple(t.shape):
            tt.ref(7, _Vector(np.int32, list(t.shape_signature)))
        tensors.append(tt)

    ops = []
    for op in g.ops:
        ot = _Table().scalar(0, "I", kinds.index((op.kind, op.version)))
        ot.ref(1, _Vector(np.int32, list(op.inputs))).ref(2, _Vector(np.int32, list(op.outputs)))
        opt = _options_table(op.kind, op.options)
        if opt is not None:
            ot.scalar(3, "B", _OPT_TAG.get(op.kind, 0))
            ot.ref(4, opt)
        ops.append(ot)

    sg = _Table()
    sg.ref(0, _Vector("tables", tensors)).ref(1, _Vector(np.int32, list(g.inputs))).ref(2, _Vector(np.int32, list(g.outputs)))
    sg.ref(3, _Vector("tables", ops)).ref(4, _Vector("str", "main"))
    model = _Table().scalar(0, "I", 3)
    model.ref(1, _Vector("tables", opcodes)).ref(2, _Vector("tables", [sg]))
    model.ref(3, _Vector("str", g.description or "birdnet_stm32 B200 PTQ")).ref(4, _Vector("tables", buffers))
    return _Builder().finish(model, b"TFL3")
