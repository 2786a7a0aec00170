"""Exporter: parsed `.tflite` graph -> flat blob (`include/bn_blob.h`).

This is the new step the B200 engine adds to the reference's conversion
subsystem (`birdnet_stm32/conversion/quantize.py:111-168` writes the `.tflite`;
this module flattens it).  It lowers the TFLite op list to the engine's op set
and resolves, once and in double precision, every integer the int8 kernels
need: per-channel requantisation multipliers and shifts, activation clamp
ranges, SAME-padding offsets, the ADD three-multiplier scheme, the MEAN
multipliers and the LOGISTIC lookup table.  The arithmetic restates TFLite's
`kernel_util.cc` / `quantization_util.cc` preparation code (third-party, not
vendored in the reference; TF 2.19 is the pinned version, requirements.txt:2).
"""

from __future__ import annotations

import math
import struct

import numpy as np

from birdnet_stm32.conversion.tflite_reader import Graph, OpInfo, TensorInfo, read_tflite

BLOB_MAGIC = b"BNB200\0\0"
BLOB_VERSION = 2
HEADER_BYTES = 128
TENSOR_BYTES = 56
OP_BYTES = 176

DT_F32, DT_I8, DT_I32 = 0, 1, 2
FE_NONE, FE_HYBRID, FE_MEL, FE_RAW = 0, 1, 2, 3
MAG = {"none": 0, "pwl": 1, "pcen": 2, "db": 3}

OP = dict(QUANTIZE=1, DEQUANTIZE=2, TRANSPOSE=3, SLICE=4, FILL=5, CONCAT=6, CONV2D=7, DWCONV2D=8,
          FC=9, ADD=10, MUL=11, MEAN=12, LOGISTIC=13, RESHAPE=14, SOFTMAX=15, PAD=16, SUM=17,
          REDUCE_MAX=18, REQUANT=19)


# ---------------------------------------------------------------------------
# TFLite preparation arithmetic
# ---------------------------------------------------------------------------
def quantize_multiplier(m: float) -> tuple[int, int]:
    """TFLite `QuantizeMultiplier(double)` -> (int32 significand, shift)."""
    if m == 0.0:
        return 0, 0
    q, shift = math.frexp(m)
    # TfLiteRound == std::round: half away from zero (q > 0 here)
    q_fixed = int(math.floor(q * float(1 << 31) + 0.5))
    assert q_fixed <= (1 << 31)
    if q_fixed == (1 << 31):
        q_fixed //= 2
        shift += 1
    if shift < -31:
        shift, q_fixed = 0, 0
    return q_fixed, shift


def _round_half_away(x: float) -> int:
    return int(math.floor(abs(x) + 0.5)) * (1 if x >= 0 else -1)


def activation_range(act: str, scale: float, zp: int) -> tuple[int, int]:
    """TFLite `CalculateActivationRangeQuantized` for int8."""
    qmin, qmax = -128, 127
    f32 = np.float32

    def quant(v):
        return zp + _round_half_away(float(f32(v) / f32(scale)))

    if act == "NONE":
        return qmin, qmax
    if act == "RELU":
        return max(qmin, quant(0.0)), qmax
    if act == "RELU6":
        return max(qmin, quant(0.0)), min(qmax, quant(6.0))
    if act == "RELU_N1_TO_1":
        return max(qmin, quant(-1.0)), min(qmax, quant(1.0))
    raise ValueError(f"unsupported fused activation {act}")


def same_padding(in_size: int, k: int, stride: int, dilation: int = 1) -> tuple[int, int]:
    """Return (out_size, pad_before) for SAME padding (TFLite ComputePaddingHeightWidth)."""
    out = (in_size + stride - 1) // stride
    keff = (k - 1) * dilation + 1
    total = max((out - 1) * stride + keff - in_size, 0)
    return out, total // 2


def logistic_lut(in_scale: float, in_zp: int, out_scale: float, out_zp: int) -> np.ndarray:
    """256-entry int8 LUT, index (uint8)(q+128); TFLite `LUTPopulate<int8_t>` with float32 maths."""
    f32 = np.float32
    lut = np.zeros(256, dtype=np.int8)
    inv = f32(1.0) / f32(out_scale)
    for q in range(-128, 128):
        x = f32(in_scale) * f32(q - in_zp)
        e = f32(math.exp(-float(x)))              # std::exp(float) -> correctly rounded float
        y = f32(1.0) / (f32(1.0) + e)
        r = _round_half_away(float(y * inv))
        lut[q + 128] = max(-128, min(127, r + out_zp))
    return lut


# ---------------------------------------------------------------------------
# blob assembly
# ---------------------------------------------------------------------------
class _Data:
    def __init__(self):
        self.chunks: list[bytes] = []
        self.size = 0

    def add(self, arr: np.ndarray) -> int:
        """Append bytes 16-byte aligned; return offset relative to the data section."""
        pad = (-self.size) % 16
        if pad:
            self.chunks.append(b"\0" * pad)
            self.size += pad
        off = self.size
        raw = np.ascontiguousarray(arr).tobytes()
        self.chunks.append(raw)
        self.size += len(raw)
        return off


def _dims3(shape: tuple) -> tuple[int, tuple]:
    """Drop the batch dim, return (rank, 3 dims leading-padded with 1)."""
    nb = tuple(int(d) for d in shape[1:])
    if not 1 <= len(nb) <= 3:
        raise ValueError(f"unsupported tensor rank for shape {shape}")
    return len(nb), (1,) * (3 - len(nb)) + nb


class Lowering:
    """Lower a TFLite :class:`Graph` to the blob's tensor and op tables."""

    def __init__(self, g: Graph):
        self.g = g
        self.data = _Data()
        self.tensors: list[dict] = []
        self.slot: dict[int, int] = {}
        self.ops: list[dict] = []

    # -- tensors -------------------------------------------------------------
    def tslot(self, tid: int) -> int:
        if tid in self.slot:
            return self.slot[tid]
        t: TensorInfo = self.g.tensor(tid)
        if t.dtype == np.float32:
            dt, esz = DT_F32, 4
        elif t.dtype == np.int8:
            dt, esz = DT_I8, 1
        elif t.dtype == np.int32:
            dt, esz = DT_I32, 4
        else:
            raise ValueError(f"tensor {tid}: unsupported dtype {t.dtype}")
        if t.is_const:
            # constants have no batch dim
            nb = tuple(int(d) for d in t.shape) or (1,)
            if len(nb) > 3:
                nb = (int(np.prod(nb[:-2])),) + nb[-2:]
            rank, dims = len(nb), (1,) * (3 - len(nb)) + nb
            off = self.data.add(t.data)
        else:
            if t.shape_signature and t.shape_signature[0] not in (-1, 1):
                raise ValueError(f"tensor {tid}: leading dim must be the batch dim")
            if any(d < 0 for d in t.shape_signature[1:]):
                raise ValueError(f"tensor {tid}: only the batch dim may be dynamic")
            rank, dims = _dims3(t.shape)
            off = None
        scale = float(t.scale[0]) if t.scale.size == 1 else 0.0
        zp = int(t.zero_point[0]) if t.zero_point.size == 1 else 0
        rec = dict(id=tid, dtype=dt, rank=rank, dims=dims, scale=scale, zp=zp,
                   is_const=int(t.is_const), data_rel=off, nbytes=int(np.prod(dims)) * esz)
        self.tensors.append(rec)
        self.slot[tid] = len(self.tensors) - 1
        return self.slot[tid]

    def emit(self, kind: str, op: OpInfo, ins: list[int], out: int, p=(), f=(), arrays=()):
        rec = dict(kind=OP[kind], tfl_index=op.index, ins=[self.tslot(i) for i in ins], out=self.tslot(out),
                   p=list(int(v) for v in p), f=list(float(v) for v in f),
                   off_rel=[self.data.add(a) for a in arrays])
        assert len(rec["p"]) <= 24 and len(rec["f"]) <= 4 and len(rec["off_rel"]) <= 4 and len(ins) <= 3
        self.ops.append(rec)

    # -- per-op lowering -----------------------------------------------------
    def _conv_common(self, op: OpInfo, depthwise: bool, fc: bool = False):
        g = self.g
        x, w = g.tensor(op.inputs[0]), g.tensor(op.inputs[1])
        b = g.tensor(op.inputs[2]) if len(op.inputs) > 2 and op.inputs[2] >= 0 else None
        y = g.tensor(op.outputs[0])
        if x.dtype != np.int8 or w.dtype != np.int8 or y.dtype != np.int8:
            raise ValueError(f"op {op.index} {op.kind}: only int8 conv/fc is supported")
        if np.any(w.zero_point != 0):
            raise ValueError(f"op {op.index}: weight zero points must be 0")
        if fc:
            cout, cin = w.shape
            kh = kw = sh = sw = 1
            pt = pl = 0
            act = op.options.get("act", "NONE")
        else:
            _, ih, iw, ic = (1,) + _dims3(x.shape)[1]
            if depthwise:
                _, kh, kw, cout = w.shape
                cin = ic
                if op.options.get("depth_multiplier", 1) != 1 or cout != ic:
                    raise ValueError(f"op {op.index}: depth_multiplier != 1 is not supported")
            else:
                cout, kh, kw, cin = w.shape
                if cin != ic:
                    raise ValueError(f"op {op.index}: channel mismatch")
            sh, sw = op.options["stride_h"], op.options["stride_w"]
            if op.options.get("dil_h", 1) != 1 or op.options.get("dil_w", 1) != 1:
                raise ValueError(f"op {op.index}: dilation is not supported")
            if op.options["padding"] == "SAME":
                oh, pt = same_padding(ih, kh, sh)
                ow, pl = same_padding(iw, kw, sw)
            else:
                oh, pt = (ih - kh) // sh + 1, 0
                ow, pl = (iw - kw) // sw + 1, 0
            _, (yh, yw, yc) = _dims3(y.shape)
            if (oh, ow, cout) != (yh, yw, yc):
                raise ValueError(f"op {op.index}: computed out shape {(oh, ow, cout)} != {(yh, yw, yc)}")
            act = op.options["act"]
        ws = w.scale.astype(np.float64)
        if ws.size not in (1, cout):
            raise ValueError(f"op {op.index}: {ws.size} weight scales for {cout} channels")
        mult = np.zeros(cout, np.int32)
        shift = np.zeros(cout, np.int32)
        for c in range(cout):
            s = float(ws[c] if ws.size > 1 else ws[0])
            m, e = quantize_multiplier(float(np.float64(x.s()) * s / np.float64(y.s())))
            mult[c], shift[c] = m, e
        bias = b.data.astype(np.int32).reshape(-1) if b is not None else np.zeros(cout, np.int32)
        if bias.size != cout:
            raise ValueError(f"op {op.index}: bias size")
        amin, amax = activation_range(act, y.s(), y.zp())
        weights = w.data.astype(np.int8)
        if depthwise:
            weights = weights.reshape(kh, kw, cout)
        kind = "FC" if fc else ("DWCONV2D" if depthwise else "CONV2D")
        if fc and int(np.prod(_dims3(x.shape)[1])) != cin:
            # Dense applied to every position of [B, ..., K] (keep_num_dims, e.g. the attention-pooling score of
            # models/blocks.py:151-156): the same arithmetic as a 1x1 convolution over the leading dims
            if _dims3(x.shape)[1][2] != cin:
                raise ValueError(f"op {op.index}: FULLY_CONNECTED input {x.shape} does not end in K={cin}")
            kind = "CONV2D"
        self.emit(kind, op, [op.inputs[0]], op.outputs[0],
                  p=[kh, kw, sh, sw, pt, pl, x.zp(), y.zp(), amin, amax, cin, cout],
                  arrays=[weights, bias, mult, shift])

    def _add_mul(self, op: OpInfo):
        g = self.g
        a, b, y = g.tensor(op.inputs[0]), g.tensor(op.inputs[1]), g.tensor(op.outputs[0])
        ia, ib = op.inputs[0], op.inputs[1]
        if a.is_const and not b.is_const:
            a, b, ia, ib = b, a, ib, ia
        if a.is_const:
            raise ValueError(f"op {op.index}: both {op.kind} inputs constant")
        bcast = 0
        if b.is_const:
            _, da = _dims3(a.shape)
            if b.data.size != da[2]:
                raise ValueError(f"op {op.index}: only [C] broadcast constants are supported")
            bcast = 1
        elif _dims3(a.shape) != _dims3(b.shape):
            _, da = _dims3(a.shape)
            _, db = _dims3(b.shape)
            if db[0] == 1 and db[1] == 1 and db[2] == da[2]:
                bcast = 2   # per-item [1,1,C] activation broadcast over H,W (SE gate)
            elif db[2] == 1 and db[:2] == da[:2] and op.kind == "MUL":
                bcast = 3   # per-position [H,W,1] activation broadcast over C (attention weights, blocks.py:157)
            else:
                raise ValueError(f"op {op.index}: unsupported broadcast {a.shape} vs {b.shape}")
        amin, amax = activation_range(op.options.get("act", "NONE"), y.s(), y.zp())
        if op.kind == "ADD":
            left_shift = 20
            s1, s2, so = np.float32(a.s()), np.float32(b.s()), np.float32(y.s())
            twice_max = np.float64(np.float32(2.0) * max(s1, s2))
            m1, e1 = quantize_multiplier(float(np.float64(s1) / twice_max))
            m2, e2 = quantize_multiplier(float(np.float64(s2) / twice_max))
            denom = np.float64(np.float32(1 << left_shift) * so)
            mo, eo = quantize_multiplier(float(twice_max / denom))
            self.emit("ADD", op, [ia, ib], op.outputs[0],
                      p=[a.zp(), b.zp(), y.zp(), left_shift, m1, e1, m2, e2, mo, eo, amin, amax, bcast])
        else:
            real = float(np.float64(np.float32(a.s()) * np.float32(b.s())) / np.float64(y.s()))
            m, e = quantize_multiplier(real)
            self.emit("MUL", op, [ia, ib], op.outputs[0], p=[a.zp(), b.zp(), y.zp(), m, e, amin, amax, bcast])

    def _mean(self, op: OpInfo):
        g = self.g
        x, ax, y = g.tensor(op.inputs[0]), g.tensor(op.inputs[1]), g.tensor(op.outputs[0])
        axes = sorted(int(v) % len(x.shape) for v in ax.data.reshape(-1))
        if len(x.shape) != 4 or axes != [1, 2]:
            raise ValueError(f"op {op.index}: MEAN only over H,W of a 4-D tensor")
        n = int(x.shape[1] * x.shape[2])
        m, e = quantize_multiplier(float(np.float64(x.s()) / np.float64(y.s())))
        # variant (ii): fold 1/N into the multiplier (TFLite reduce.cc, >= 2.10)
        sh = min(n.bit_length() - 1, 32, 31 + e)
        mn = int((m << sh) // n) if sh >= 0 else 0
        en = e - sh
        # keep int32 range
        assert -(1 << 31) <= mn < (1 << 31)
        self.emit("MEAN", op, [op.inputs[0]], op.outputs[0],
                  p=[n, x.zp(), y.zp(), m, e, mn, en, int(op.options.get("keep_dims", False))])

    def lower(self):
        g = self.g
        shape_tensors: set[int] = set()   # int32 shape-computation tensors, folded away
        for op in g.ops:
            k = op.kind
            outs = op.outputs
            if k == "SHAPE":
                shape_tensors.add(outs[0])
                continue
            if k in ("STRIDED_SLICE", "PACK") and any(i in shape_tensors for i in op.inputs):
                shape_tensors.add(outs[0])
                continue
            if k == "QUANTIZE":
                x, y = g.tensor(op.inputs[0]), g.tensor(outs[0])
                if x.dtype == np.float32:
                    self.emit("QUANTIZE", op, [op.inputs[0]], outs[0], p=[y.zp()], f=[y.s()])
                else:
                    m, e = quantize_multiplier(float(np.float64(x.s()) / np.float64(y.s())))
                    self.emit("REQUANT", op, [op.inputs[0]], outs[0], p=[x.zp(), y.zp(), m, e])
            elif k == "DEQUANTIZE":
                x = g.tensor(op.inputs[0])
                self.emit("DEQUANTIZE", op, [op.inputs[0]], outs[0], p=[x.zp()], f=[x.s()])
            elif k == "TRANSPOSE":
                perm = [int(v) for v in g.tensor(op.inputs[1]).data.reshape(-1)]
                x = g.tensor(op.inputs[0])
                if len(perm) != 4 or perm[0] != 0 or len(x.shape) != 4:
                    raise ValueError(f"op {op.index}: TRANSPOSE must be 4-D and keep the batch dim")
                self.emit("TRANSPOSE", op, [op.inputs[0]], outs[0], p=[v - 1 for v in perm[1:]])
            elif k == "STRIDED_SLICE":
                x, y = g.tensor(op.inputs[0]), g.tensor(outs[0])
                begin = g.tensor(op.inputs[1]).data.reshape(-1)
                strides = g.tensor(op.inputs[3]).data.reshape(-1)
                o = op.options
                if np.any(strides != 1) or o["ellipsis_mask"] or o["new_axis_mask"] or o["shrink_axis_mask"]:
                    raise ValueError(f"op {op.index}: only unit-stride plain STRIDED_SLICE is supported")
                rank = len(x.shape)
                b = [0 if (o["begin_mask"] >> i) & 1 else int(begin[i]) % max(int(x.shape[i]), 1) for i in range(rank)]
                if b[0] != 0:
                    raise ValueError(f"op {op.index}: slicing the batch dim is not supported")
                b3 = [0] * (4 - rank) + b[1:]
                self.emit("SLICE", op, [op.inputs[0]], outs[0], p=b3[-3:])
            elif k == "FILL":
                val = g.tensor(op.inputs[1])
                if not val.is_const or op.inputs[0] not in shape_tensors:
                    raise ValueError(f"op {op.index}: FILL needs a constant value and a folded shape")
                self.emit("FILL", op, [], outs[0], p=[int(val.data.reshape(-1)[0])])
            elif k == "CONCATENATION":
                if len(op.inputs) != 2:
                    raise ValueError(f"op {op.index}: CONCATENATION with {len(op.inputs)} inputs")
                y = g.tensor(outs[0])
                for i in op.inputs:
                    t = g.tensor(i)
                    if t.s() != y.s() or t.zp() != y.zp():
                        raise ValueError(f"op {op.index}: CONCATENATION inputs must share quant params")
                rank = len(y.shape)
                axis = op.options["axis"] % rank
                if axis == 0:
                    raise ValueError(f"op {op.index}: cannot concatenate along batch")
                self.emit("CONCAT", op, list(op.inputs), outs[0], p=[axis - 1 + (3 - (rank - 1))])
            elif k == "CONV_2D":
                self._conv_common(op, depthwise=False)
            elif k == "DEPTHWISE_CONV_2D":
                self._conv_common(op, depthwise=True)
            elif k == "FULLY_CONNECTED":
                self._conv_common(op, depthwise=False, fc=True)
            elif k in ("ADD", "MUL"):
                self._add_mul(op)
            elif k == "MEAN":
                self._mean(op)
            elif k == "LOGISTIC":
                x, y = g.tensor(op.inputs[0]), g.tensor(outs[0])
                self.emit("LOGISTIC", op, [op.inputs[0]], outs[0], arrays=[logistic_lut(x.s(), x.zp(), y.s(), y.zp())])
            elif k == "RESHAPE":
                self.emit("RESHAPE", op, [op.inputs[0]], outs[0])
            elif k == "PAD":
                x, y = g.tensor(op.inputs[0]), g.tensor(outs[0])
                pads = g.tensor(op.inputs[1]).data.reshape(-1, 2).astype(int)
                if x.s() != y.s() or x.zp() != y.zp():
                    raise ValueError(f"op {op.index}: PAD must keep the quantisation parameters")
                if pads.shape[0] != len(x.shape) or pads[0].any():
                    raise ValueError(f"op {op.index}: PAD must not touch the batch dim")
                rows = [(0, 0)] * (4 - len(x.shape)) + [tuple(r) for r in pads[1:]]
                self.emit("PAD", op, [op.inputs[0]], outs[0], p=[v for r in rows[-3:] for v in r] + [x.zp()])
            elif k == "SOFTMAX":
                x, y = g.tensor(op.inputs[0]), g.tensor(outs[0])
                beta = float(op.options.get("beta", 1.0)) or 1.0
                # optimized_ops::PopulateSoftmaxLookupTable: table[255 - v] = expf(-input_scale * beta * v), float32
                sc = np.float32(-np.float32(x.s()) * np.float32(beta))
                table = np.exp((sc * np.arange(256, dtype=np.float32)).astype(np.float32)).astype(np.float32)[::-1].copy()
                self.emit("SOFTMAX", op, [op.inputs[0]], outs[0], p=[y.zp()], f=[float(np.float32(x.s()) * np.float32(beta)), y.s()],
                          arrays=[table])
            elif k == "SUM":
                x, ax, y = g.tensor(op.inputs[0]), g.tensor(op.inputs[1]), g.tensor(outs[0])
                axes = [int(v) % len(x.shape) for v in ax.data.reshape(-1)]
                if len(axes) != 1 or axes[0] == 0:
                    raise ValueError(f"op {op.index}: SUM over exactly one non-batch axis is supported")
                axis3 = axes[0] - 1 + (3 - (len(x.shape) - 1))
                count = int(x.shape[axes[0]])
                scale = np.float32(np.float32(x.s()) / np.float32(y.s()))
                bias = np.float32(np.float32(np.float32(-x.zp()) * scale) * np.float32(count))
                self.emit("SUM", op, [op.inputs[0]], outs[0], p=[axis3, count, x.zp(), y.zp()], f=[float(scale), float(bias)])
            else:
                raise ValueError(f"op {op.index}: unsupported operator {k}")
        return self


def export_blob(graph_or_path, cfg: dict | None = None) -> bytes:
    """Flatten a `.tflite` (path, bytes or parsed Graph) into the engine blob.

    Args:
        graph_or_path: `.tflite` path / bytes or a parsed :class:`Graph`.
        cfg: the `_model_config.json` dict (audio parameters only are used).
    """
    g = graph_or_path if isinstance(graph_or_path, Graph) else read_tflite(graph_or_path)
    if len(g.inputs) != 1 or len(g.outputs) != 1:
        raise ValueError("exactly one graph input and one output are required")
    low = Lowering(g)
    in_slot = low.tslot(g.inputs[0])
    low.lower()
    out_slot = low.tslot(g.outputs[0])

    cfg = dict(cfg or {})
    fe_name = cfg.get("audio_frontend", "")
    fe = {"hybrid": FE_HYBRID, "librosa": FE_MEL, "precomputed": FE_MEL, "raw": FE_RAW, "tf": FE_RAW}.get(fe_name, FE_NONE)
    sr = int(cfg.get("sample_rate", 0))
    chunk_len = int(sr * float(cfg.get("chunk_duration", 0)))
    spec_width = int(cfg.get("spec_width", 0))
    n_fft = int(cfg.get("fft_length", 0))
    hop = chunk_len // spec_width if spec_width > 0 else 0
    out_t = low.tensors[out_slot]
    num_classes = int(out_t["dims"][2])

    n_t, n_o = len(low.tensors), len(low.ops)
    tensors_off = HEADER_BYTES
    ops_off = tensors_off + n_t * TENSOR_BYTES
    data_off = (ops_off + n_o * OP_BYTES + 15) // 16 * 16
    total = data_off + low.data.size

    hdr = struct.pack("<8sIIIIQQQQ9I2i7I", BLOB_MAGIC, BLOB_VERSION, HEADER_BYTES, n_t, n_o,
                      tensors_off, ops_off, data_off, total,
                      fe, MAG.get(cfg.get("mag_scale", "none"), 0), sr, chunk_len, n_fft, hop, spec_width,
                      int(cfg.get("num_mels", 0)), num_classes,
                      in_slot, out_slot, *([0] * 7))
    assert len(hdr) == HEADER_BYTES
    out = bytearray(hdr)
    for t in low.tensors:
        out += struct.pack("<3i3ifiiiQQ", t["id"], t["dtype"], t["rank"], *t["dims"], t["scale"], t["zp"],
                           t["is_const"], 0, (data_off + t["data_rel"]) if t["data_rel"] is not None else 0,
                           t["nbytes"])
    for o in low.ops:
        ins = o["ins"] + [-1] * (3 - len(o["ins"]))
        p = o["p"] + [0] * (24 - len(o["p"]))
        f = o["f"] + [0.0] * (4 - len(o["f"]))
        offs = [data_off + r for r in o["off_rel"]] + [0] * (4 - len(o["off_rel"]))
        out += struct.pack("<3i3iii24i4f4Q", o["kind"], o["tfl_index"], len(o["ins"]), *ins, o["out"], 0, *p, *f, *offs)
    out = out.ljust(data_off, b"\0")
    out += b"".join(low.data.chunks)
    assert len(out) == total, (len(out), total)
    return bytes(out)


def export_blob_file(tflite_path: str, config_path: str | None, out_path: str) -> str:
    """CLI helper: write `<name>.b200blob` next to a converted model."""
    import json

    cfg = json.loads(open(config_path).read()) if config_path else {}
    blob = export_blob(tflite_path, cfg)
    with open(out_path, "wb") as fh:
        fh.write(blob)
    return out_path
