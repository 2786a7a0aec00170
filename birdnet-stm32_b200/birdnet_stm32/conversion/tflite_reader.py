"""Dependency-free reader for `.tflite` flatbuffers -> a small graph IR.

The engine is driven by the `.tflite` file itself (op list, shapes, scales,
zero points) because the reference's `_model_config.json` does not describe the
architecture reliably (legacy checkpoints default `use_se=True` etc. on load,
reference `birdnet_stm32/training/config.py:57-59,117-131`).  The reference
hands the same file to `tf.lite.Interpreter` (`models/runners.py:57`); this
module is the TensorFlow-free replacement for that loading step.

Only the flatbuffer mechanics and the schema field slots of TFLite schema v3
that this path needs are implemented (Model, OperatorCode, SubGraph, Tensor,
QuantizationParameters, Operator, Buffer and the option tables of the
supported ops).
"""

from __future__ import annotations

import struct
from dataclasses import dataclass, field

import numpy as np

# -- builtin operator codes (TFLite schema v3 `BuiltinOperator`) -------------
BUILTIN = {
    0: "ADD", 1: "AVERAGE_POOL_2D", 2: "CONCATENATION", 3: "CONV_2D", 4: "DEPTHWISE_CONV_2D",
    6: "DEQUANTIZE", 9: "FULLY_CONNECTED", 14: "LOGISTIC", 18: "MUL", 19: "RELU", 21: "RELU6",
    22: "RESHAPE", 25: "SOFTMAX", 34: "PAD", 39: "TRANSPOSE", 40: "MEAN", 41: "SUB", 42: "DIV",
    45: "STRIDED_SLICE", 55: "MAXIMUM", 73: "LOG", 74: "SUM", 77: "SHAPE", 82: "REDUCE_MAX",
    83: "PACK", 94: "FILL", 114: "QUANTIZE",
}

# TensorType enum -> numpy dtype
TENSOR_DTYPES = {0: np.float32, 1: np.float16, 2: np.int32, 3: np.uint8, 4: np.int64, 6: np.bool_, 7: np.int16, 9: np.int8}

FUSED_ACT = {0: "NONE", 1: "RELU", 2: "RELU_N1_TO_1", 3: "RELU6"}
PADDING = {0: "SAME", 1: "VALID"}


class _FB:
    """Minimal flatbuffer accessor (little-endian, offsets relative as per spec)."""

    def __init__(self, buf: bytes):
        self.b = buf

    def u8(self, o): return self.b[o]
    def i8(self, o): return struct.unpack_from("<b", self.b, o)[0]
    def u16(self, o): return struct.unpack_from("<H", self.b, o)[0]
    def i32(self, o): return struct.unpack_from("<i", self.b, o)[0]
    def u32(self, o): return struct.unpack_from("<I", self.b, o)[0]
    def f32(self, o): return struct.unpack_from("<f", self.b, o)[0]

    def root(self):
        return self.u32(0)

    def field(self, tbl: int, idx: int) -> int:
        """Absolute offset of field `idx` in table `tbl`, or 0 when absent."""
        vt = tbl - self.i32(tbl)
        vt_len = self.u16(vt)
        slot = 4 + 2 * idx
        if slot >= vt_len:
            return 0
        off = self.u16(vt + slot)
        return tbl + off if off else 0

    def indirect(self, o: int) -> int:
        return o + self.u32(o)

    def table(self, tbl, idx):
        f = self.field(tbl, idx)
        return self.indirect(f) if f else 0

    def scalar(self, tbl, idx, kind, default=0):
        f = self.field(tbl, idx)
        if not f:
            return default
        return getattr(self, kind)(f)

    def vector(self, tbl, idx):
        """Return (start offset of elements, length) or (0, 0)."""
        f = self.field(tbl, idx)
        if not f:
            return 0, 0
        v = self.indirect(f)
        return v + 4, self.u32(v)

    def np_vector(self, tbl, idx, dtype):
        start, n = self.vector(tbl, idx)
        if not n:
            return np.zeros((0,), dtype=dtype)
        return np.frombuffer(self.b, dtype=dtype, count=n, offset=start).copy()

    def string(self, tbl, idx):
        start, n = self.vector(tbl, idx)
        return self.b[start:start + n].decode("utf-8", "replace") if n else ""

    def table_vector(self, tbl, idx):
        start, n = self.vector(tbl, idx)
        return [self.indirect(start + 4 * i) for i in range(n)]


@dataclass
class TensorInfo:
    """One tensor of the subgraph (activation or constant)."""

    index: int
    name: str
    shape: tuple
    shape_signature: tuple
    dtype: type
    scale: np.ndarray          # float32 [n] (empty if not quantised)
    zero_point: np.ndarray     # int64 [n]
    quantized_dimension: int
    data: np.ndarray | None    # constant payload or None

    @property
    def is_const(self) -> bool:
        return self.data is not None

    @property
    def quantized(self) -> bool:
        return self.scale.size > 0

    def s(self) -> float:
        """Per-tensor scale as python float (float32 value)."""
        if self.scale.size != 1:
            raise ValueError(f"tensor {self.index} ({self.name}) is not per-tensor quantised")
        return float(self.scale[0])

    def zp(self) -> int:
        return int(self.zero_point[0]) if self.zero_point.size else 0


@dataclass
class OpInfo:
    """One operator: builtin name, tensor indices and decoded options."""

    index: int
    kind: str
    version: int
    inputs: list
    outputs: list
    options: dict = field(default_factory=dict)


@dataclass
class Graph:
    """Flat IR of one `.tflite` subgraph."""

    tensors: list
    ops: list
    inputs: list
    outputs: list
    description: str = ""

    def tensor(self, i: int) -> TensorInfo:
        return self.tensors[i]


def _decode_options(fb: _FB, kind: str, tbl: int) -> dict:
    if not tbl:
        return {}
    if kind == "CONV_2D":
        return dict(padding=PADDING[fb.scalar(tbl, 0, "i8")], stride_w=fb.scalar(tbl, 1, "i32"),
                    stride_h=fb.scalar(tbl, 2, "i32"), act=FUSED_ACT[fb.scalar(tbl, 3, "i8")],
                    dil_w=fb.scalar(tbl, 4, "i32", 1), dil_h=fb.scalar(tbl, 5, "i32", 1))
    if kind == "DEPTHWISE_CONV_2D":
        return dict(padding=PADDING[fb.scalar(tbl, 0, "i8")], stride_w=fb.scalar(tbl, 1, "i32"),
                    stride_h=fb.scalar(tbl, 2, "i32"), depth_multiplier=fb.scalar(tbl, 3, "i32"),
                    act=FUSED_ACT[fb.scalar(tbl, 4, "i8")],
                    dil_w=fb.scalar(tbl, 5, "i32", 1), dil_h=fb.scalar(tbl, 6, "i32", 1))
    if kind == "FULLY_CONNECTED":
        return dict(act=FUSED_ACT[fb.scalar(tbl, 0, "i8")], weights_format=fb.scalar(tbl, 1, "i8"),
                    keep_num_dims=bool(fb.scalar(tbl, 2, "u8")))
    if kind in ("ADD", "MUL", "SUB", "DIV"):
        return dict(act=FUSED_ACT[fb.scalar(tbl, 0, "i8")])
    if kind in ("MEAN", "SUM", "REDUCE_MAX"):
        return dict(keep_dims=bool(fb.scalar(tbl, 0, "u8")))
    if kind == "AVERAGE_POOL_2D":
        return dict(padding=PADDING[fb.scalar(tbl, 0, "i8")], stride_w=fb.scalar(tbl, 1, "i32"),
                    stride_h=fb.scalar(tbl, 2, "i32"), filter_w=fb.scalar(tbl, 3, "i32"),
                    filter_h=fb.scalar(tbl, 4, "i32"), act=FUSED_ACT[fb.scalar(tbl, 5, "i8")])
    if kind == "SOFTMAX":
        return dict(beta=fb.scalar(tbl, 0, "f32", 0.0))
    if kind == "CONCATENATION":
        return dict(axis=fb.scalar(tbl, 0, "i32"), act=FUSED_ACT[fb.scalar(tbl, 1, "i8")])
    if kind == "STRIDED_SLICE":
        return dict(begin_mask=fb.scalar(tbl, 0, "i32"), end_mask=fb.scalar(tbl, 1, "i32"),
                    ellipsis_mask=fb.scalar(tbl, 2, "i32"), new_axis_mask=fb.scalar(tbl, 3, "i32"),
                    shrink_axis_mask=fb.scalar(tbl, 4, "i32"))
    if kind == "PACK":
        return dict(values_count=fb.scalar(tbl, 0, "i32"), axis=fb.scalar(tbl, 1, "i32"))
    if kind == "RESHAPE":
        return dict(new_shape=tuple(int(v) for v in fb.np_vector(tbl, 0, np.int32)))
    return {}


def read_tflite(path_or_bytes) -> Graph:
    """Parse a `.tflite` file (path or bytes) into a :class:`Graph`.

    Raises ValueError on unsupported structure (more than one subgraph,
    custom ops, unknown builtin codes).
    """
    if isinstance(path_or_bytes, (bytes, bytearray, memoryview)):
        buf = bytes(path_or_bytes)
    else:
        with open(path_or_bytes, "rb") as fh:
            buf = fh.read()
    if len(buf) < 8 or buf[4:8] != b"TFL3":
        raise ValueError("not a TFLite schema-v3 flatbuffer (missing 'TFL3' identifier)")
    fb = _FB(buf)
    model = fb.root()

    # operator codes: builtin = max(deprecated_builtin_code, builtin_code)
    opcodes = []
    for oc in fb.table_vector(model, 1):
        dep = fb.scalar(oc, 0, "i8")
        new = fb.scalar(oc, 3, "i32")
        code = max(dep, new)
        if fb.field(oc, 1):
            raise ValueError("custom operators are not supported")
        opcodes.append((code, fb.scalar(oc, 2, "i32", 1)))

    subgraphs = fb.table_vector(model, 2)
    if len(subgraphs) != 1:
        raise ValueError(f"expected exactly 1 subgraph, found {len(subgraphs)}")
    sg = subgraphs[0]

    buffers = []
    for bt in fb.table_vector(model, 4):
        start, n = fb.vector(bt, 0)
        buffers.append(buf[start:start + n] if n else b"")

    tensors = []
    for ti, tt in enumerate(fb.table_vector(sg, 0)):
        shape = tuple(int(v) for v in fb.np_vector(tt, 0, np.int32))
        ttype = fb.scalar(tt, 1, "i8")
        if ttype not in TENSOR_DTYPES:
            raise ValueError(f"tensor {ti}: unsupported TensorType {ttype}")
        dtype = TENSOR_DTYPES[ttype]
        bidx = fb.scalar(tt, 2, "u32")
        name = fb.string(tt, 3)
        q = fb.table(tt, 4)
        if q:
            scale = fb.np_vector(q, 2, np.float32)
            zp = fb.np_vector(q, 3, np.int64)
            qdim = fb.scalar(q, 6, "i32")
        else:
            scale, zp, qdim = np.zeros((0,), np.float32), np.zeros((0,), np.int64), 0
        sig = tuple(int(v) for v in fb.np_vector(tt, 7, np.int32)) or shape
        raw = buffers[bidx] if bidx < len(buffers) else b""
        data = None
        if len(raw):
            data = np.frombuffer(raw, dtype=dtype).copy()
            if int(np.prod(shape, dtype=np.int64)) == data.size:
                data = data.reshape(shape)
        tensors.append(TensorInfo(ti, name, shape, sig, dtype, scale, zp, qdim, data))

    ops = []
    for oi, ot in enumerate(fb.table_vector(sg, 3)):
        code, ver = opcodes[fb.scalar(ot, 0, "u32")]
        if code not in BUILTIN:
            raise ValueError(f"op {oi}: unsupported builtin operator code {code}")
        kind = BUILTIN[code]
        ins = [int(v) for v in fb.np_vector(ot, 1, np.int32)]
        outs = [int(v) for v in fb.np_vector(ot, 2, np.int32)]
        opts = _decode_options(fb, kind, fb.table(ot, 4))
        ops.append(OpInfo(oi, kind, ver, ins, outs, opts))

    g = Graph(tensors=tensors, ops=ops,
              inputs=[int(v) for v in fb.np_vector(sg, 1, np.int32)],
              outputs=[int(v) for v in fb.np_vector(sg, 2, np.int32)],
              description=fb.string(model, 3))
    return g
