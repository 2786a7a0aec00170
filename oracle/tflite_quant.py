"""Oracle-side derivation of every requantisation constant, straight from the `.tflite` file.

TEST INFRASTRUCTURE ONLY (like everything under `oracle/`).

Why this exists: the product's exporter (`birdnet_stm32/conversion/export_blob.py`) resolves
`QuantizeMultiplier`, the ADD three-multiplier scheme, the MEAN multipliers, activation clamp ranges
and the LOGISTIC table once and writes them into the blob that BOTH the CUDA engine and the C oracle
load.  A wrong tie-break there would be invisible to a parity test.  This module shares no code with
the product: it has its own flatbuffer walker (schema slots of SURVEY.md Appendix D) and derives the
same integers a second way -- exact rational arithmetic (`fractions.Fraction`) instead of `math.frexp`
on doubles, glibc `expf` through the C oracle instead of `math.exp` -- following TFLite's preparation
code (third-party, TF 2.19: `kernel_util.cc::PopulateConvolutionQuantizationParams`,
`quantization_util.cc::QuantizeMultiplier`, `add.cc::Prepare`, `reduce.cc`, `activations.cc`; the
reference reaches them through `tf.lite.Interpreter.allocate_tensors`, `models/runners.py:57-61`).

`derive(path)` -> {tflite operator index: dict of constants};  `compare_with_blob(blob, derived)` ->
list of mismatches (empty = the product's blob carries exactly these integers);  `patch_blob` writes
the oracle's own values over the blob's, so that `OracleModel(patch_blob(...))` runs on constants the
product never touched.
"""

from __future__ import annotations

import struct
from fractions import Fraction

import numpy as np

_OPNAME = {0: "ADD", 3: "CONV_2D", 4: "DEPTHWISE_CONV_2D", 6: "DEQUANTIZE", 9: "FULLY_CONNECTED", 14: "LOGISTIC",
           18: "MUL", 40: "MEAN", 114: "QUANTIZE"}
_ACT = {0: "NONE", 1: "RELU", 2: "RELU_N1_TO_1", 3: "RELU6"}
_NP = {0: np.float32, 2: np.int32, 3: np.uint8, 4: np.int64, 7: np.int16, 9: np.int8}


# ------------------------------------------------------------------------------------------------
# flatbuffer walking (tables are vtable-indexed; vectors are u32 length + payload)
# ------------------------------------------------------------------------------------------------
class _Buf:
    def __init__(self, raw: bytes):
        self.raw = raw

    def rd(self, fmt: str, at: int):
        return struct.unpack_from("<" + fmt, self.raw, at)[0]

    def slot(self, table: int, k: int) -> int:
        vt = table - self.rd("i", table)
        if 4 + 2 * k >= self.rd("H", vt):
            return 0
        rel = self.rd("H", vt + 4 + 2 * k)
        return table + rel if rel else 0

    def deref(self, at: int) -> int:
        return at + self.rd("I", at)

    def num(self, table: int, k: int, fmt: str, default=0):
        at = self.slot(table, k)
        return self.rd(fmt, at) if at else default

    def vec(self, table: int, k: int):
        at = self.slot(table, k)
        if not at:
            return 0, 0
        v = self.deref(at)
        return v + 4, self.rd("I", v)

    def sub(self, table: int, k: int) -> int:
        at = self.slot(table, k)
        return self.deref(at) if at else 0

    def array(self, table: int, k: int, dtype) -> np.ndarray:
        start, n = self.vec(table, k)
        return np.frombuffer(self.raw, dtype=dtype, count=n, offset=start).copy() if n else np.zeros(0, dtype)


class _T:
    __slots__ = ("shape", "dtype", "scale", "zp", "data")


def _load(path_or_bytes):
    raw = path_or_bytes if isinstance(path_or_bytes, (bytes, bytearray)) else open(path_or_bytes, "rb").read()
    fb = _Buf(bytes(raw))
    model = fb.deref(0)
    codes_at, n_codes = fb.vec(model, 1)
    codes = []
    for i in range(n_codes):
        oc = fb.deref(codes_at + 4 * i)
        codes.append(max(fb.num(oc, 0, "b"), fb.num(oc, 3, "i")))
    sg_at, _ = fb.vec(model, 2)
    sg = fb.deref(sg_at)
    bufs_at, _ = fb.vec(model, 4)
    tens_at, n_t = fb.vec(sg, 0)
    tensors = []
    for i in range(n_t):
        tt = fb.deref(tens_at + 4 * i)
        t = _T()
        t.shape = tuple(int(v) for v in fb.array(tt, 0, np.int32))
        t.dtype = _NP.get(fb.num(tt, 1, "b"), None)
        q = fb.sub(tt, 4)
        t.scale = fb.array(q, 2, np.float32) if q else np.zeros(0, np.float32)
        t.zp = fb.array(q, 3, np.int64) if q else np.zeros(0, np.int64)
        b = fb.deref(bufs_at + 4 * fb.num(tt, 2, "I"))
        start, n = fb.vec(b, 0)
        t.data = np.frombuffer(fb.raw, dtype=t.dtype, count=n // np.dtype(t.dtype).itemsize, offset=start).copy() if (n and t.dtype) else None
        tensors.append(t)
    ops_at, n_ops = fb.vec(sg, 3)
    ops = []
    for i in range(n_ops):
        ot = fb.deref(ops_at + 4 * i)
        code = codes[fb.num(ot, 0, "I")]
        ins = [int(v) for v in fb.array(ot, 1, np.int32)]
        outs = [int(v) for v in fb.array(ot, 2, np.int32)]
        opt = fb.sub(ot, 4)
        act = "NONE"
        keep = 0
        if opt:
            if code == 3:
                act = _ACT[fb.num(opt, 3, "b")]
            elif code == 4:
                act = _ACT[fb.num(opt, 4, "b")]
            elif code in (9, 0, 18):
                act = _ACT[fb.num(opt, 0, "b")]
            elif code == 40:
                keep = int(fb.num(opt, 0, "b"))
        ops.append((i, code, ins, outs, act, keep))
    return tensors, ops


# ------------------------------------------------------------------------------------------------
# TFLite preparation arithmetic, restated with exact rationals
# ------------------------------------------------------------------------------------------------
def quantize_multiplier(real: float) -> tuple[int, int]:
    """`QuantizeMultiplier(double)`: real = q * 2^shift with q in [0.5, 1); q_fixed = round(q * 2^31) (half away from zero;
    q > 0), renormalised when it hits 2^31; multipliers below 2^-31 flush to zero."""
    x = Fraction(real)                    # a double is an exact rational
    if x == 0:
        return 0, 0
    assert x > 0
    shift = 0
    while x >= 1:
        x /= 2
        shift += 1
    while x < Fraction(1, 2):
        x *= 2
        shift -= 1
    scaled = x * (1 << 31)                # exact
    q = int(scaled + Fraction(1, 2))      # floor(v + 1/2) == round-half-away for v > 0
    if q == 1 << 31:
        q >>= 1
        shift += 1
    if shift < -31:
        return 0, 0
    return q, shift


def _f64_div(num: Fraction, den: Fraction) -> float:
    """IEEE double division of two exactly known operands = the correctly rounded quotient."""
    return float(num / den)               # Fraction -> float is correctly rounded (round-half-even)


def _fr(v) -> Fraction:
    return Fraction(float(np.float32(v)))


def _act_range(act: str, scale, zp: int) -> tuple[int, int]:
    """`CalculateActivationRangeQuantized`: quantise 0 / 6 / -1 / 1 with float32 division and std::round."""
    def q(v):
        r = np.float32(v) / np.float32(scale)
        return zp + int(np.sign(r) * np.floor(np.abs(np.float64(r)) + 0.5))
    lo, hi = -128, 127
    if act == "RELU":
        lo = max(lo, q(0.0))
    elif act == "RELU6":
        lo, hi = max(lo, q(0.0)), min(hi, q(6.0))
    elif act == "RELU_N1_TO_1":
        lo, hi = max(lo, q(-1.0)), min(hi, q(1.0))
    return lo, hi


def derive(path_or_bytes) -> dict[int, dict]:
    tensors, ops = _load(path_or_bytes)
    out: dict[int, dict] = {}
    for idx, code, ins, outs, act, keep in ops:
        name = _OPNAME.get(code)
        if name in ("CONV_2D", "DEPTHWISE_CONV_2D", "FULLY_CONNECTED"):
            x, w, y = tensors[ins[0]], tensors[ins[1]], tensors[outs[0]]
            if x.dtype != np.int8:
                continue
            cout = w.shape[3] if name == "DEPTHWISE_CONV_2D" else w.shape[0]
            sx, sy = _fr(x.scale[0]), _fr(y.scale[0])
            mult, shift = [], []
            for c in range(cout):
                sw = _fr(w.scale[c] if w.scale.size > 1 else w.scale[0])
                # (double)input_scale * (double)filter_scale is exact (2 x 24-bit significands); the division rounds once
                m, e = quantize_multiplier(_f64_div(sx * sw, sy))
                mult.append(m)
                shift.append(e)
            lo, hi = _act_range(act, y.scale[0], int(y.zp[0]))
            out[idx] = dict(kind="conv", mult=np.array(mult, np.int32), shift=np.array(shift, np.int32), in_zp=int(x.zp[0]),
                            out_zp=int(y.zp[0]), act_min=lo, act_max=hi)
        elif name == "ADD":
            a, b, y = tensors[ins[0]], tensors[ins[1]], tensors[outs[0]]
            if a.dtype != np.int8:
                continue
            if a.data is not None and b.data is None:
                a, b = b, a               # the exporter keeps the activation first
            s1, s2, so = np.float32(a.scale[0]), np.float32(b.scale[0]), np.float32(y.scale[0])
            twice = Fraction(float(np.float32(2.0) * max(s1, s2)))            # float product, exact (x2)
            m1, e1 = quantize_multiplier(_f64_div(_fr(s1), twice))
            m2, e2 = quantize_multiplier(_f64_div(_fr(s2), twice))
            den = Fraction(float(np.float32(1 << 20) * so))                    # (1 << left_shift) * float scale, exact (x2^20)
            mo, eo = quantize_multiplier(_f64_div(twice, den))
            lo, hi = _act_range(act, y.scale[0], int(y.zp[0]))
            out[idx] = dict(kind="add", in1_zp=int(a.zp[0]), in2_zp=int(b.zp[0]), out_zp=int(y.zp[0]), left_shift=20,
                            m1=m1, s1=e1, m2=m2, s2=e2, mo=mo, so=eo, act_min=lo, act_max=hi)
        elif name == "MUL":
            a, b, y = tensors[ins[0]], tensors[ins[1]], tensors[outs[0]]
            if a.dtype != np.int8:
                continue
            if a.data is not None and b.data is None:
                a, b = b, a
            prod = Fraction(float(np.float32(a.scale[0]) * np.float32(b.scale[0])))   # float product (rounded to float32)
            m, e = quantize_multiplier(_f64_div(prod, _fr(y.scale[0])))
            lo, hi = _act_range(act, y.scale[0], int(y.zp[0]))
            out[idx] = dict(kind="mul", in1_zp=int(a.zp[0]), in2_zp=int(b.zp[0]), out_zp=int(y.zp[0]), mult=m, shift=e,
                            act_min=lo, act_max=hi)
        elif name == "MEAN":
            x, y = tensors[ins[0]], tensors[outs[0]]
            if x.dtype != np.int8:
                continue
            n = int(x.shape[1] * x.shape[2])
            m, e = quantize_multiplier(_f64_div(_fr(x.scale[0]), _fr(y.scale[0])))
            # reduce.cc (keep_dims = 0 path): fold 1/N -- shift = min(63 - clz64(N), 32, 31 + e)
            sh = min(n.bit_length() - 1, 32, 31 + e)
            mn = ((m << sh) // n) if sh >= 0 else 0
            out[idx] = dict(kind="mean", count=n, in_zp=int(x.zp[0]), out_zp=int(y.zp[0]), mult=m, shift=e, mult_n=int(mn),
                            shift_n=e - sh, keep_dims=keep)
        elif name == "LOGISTIC":
            x, y = tensors[ins[0]], tensors[outs[0]]
            if x.dtype != np.int8:
                continue
            from oracle import bn_oracle      # glibc expf, float32 maths, like LUTPopulate<int8_t>

            out[idx] = dict(kind="logistic", lut=bn_oracle.logistic_lut(x.scale[0], int(x.zp[0]), y.scale[0], int(y.zp[0])))
        elif name == "QUANTIZE":
            x, y = tensors[ins[0]], tensors[outs[0]]
            if x.dtype == np.float32:
                out[idx] = dict(kind="quantize", scale=float(y.scale[0]), zp=int(y.zp[0]))
            else:
                m, e = quantize_multiplier(_f64_div(_fr(x.scale[0]), _fr(y.scale[0])))
                out[idx] = dict(kind="requant", in_zp=int(x.zp[0]), out_zp=int(y.zp[0]), mult=m, shift=e)
        elif name == "DEQUANTIZE":
            x = tensors[ins[0]]
            out[idx] = dict(kind="dequantize", scale=float(x.scale[0]), zp=int(x.zp[0]))
    return out


# ------------------------------------------------------------------------------------------------
# blob side (layout: include/bn_blob.h; read with struct, no product import)
# ------------------------------------------------------------------------------------------------
_OP_BYTES, _HDR = 176, "<8sIIIIQQQQ"
_K_QUANT, _K_DEQ, _K_CONV, _K_DW, _K_FC, _K_ADD, _K_MUL, _K_MEAN, _K_LOGI, _K_REQ = 1, 2, 7, 8, 9, 10, 11, 12, 13, 19


def _blob_ops(blob: bytes):
    magic, ver, hb, n_t, n_o, t_off, o_off, d_off, total = struct.unpack_from(_HDR, blob, 0)
    assert magic == b"BNB200\0\0" and total == len(blob)
    for i in range(n_o):
        base = o_off + i * _OP_BYTES
        kind, tfl, n_in = struct.unpack_from("<3i", blob, base)
        p_at, f_at, off_at = base + 32, base + 32 + 96, base + 32 + 96 + 16
        yield dict(kind=kind, tfl=tfl, p_at=p_at, f_at=f_at, p=list(struct.unpack_from("<24i", blob, p_at)),
                   f=list(struct.unpack_from("<4f", blob, f_at)), off=list(struct.unpack_from("<4Q", blob, off_at)))


def _fields(op: dict, d: dict, blob: bytes):
    """Yield (description, blob value, oracle value, writer) for one op."""
    p = op["p"]
    k = op["kind"]

    def pslot(i, name, want):
        return (f"op {op['tfl']} {name}", p[i], int(want), ("p", op["p_at"] + 4 * i, int(want)))

    if k in (_K_CONV, _K_DW, _K_FC) and d["kind"] == "conv":
        cout = p[11]
        for j, name in ((2, "mult"), (3, "shift")):
            have = np.frombuffer(blob, np.int32, cout, op["off"][j])
            yield (f"op {op['tfl']} {name}[{cout}]", have, d[name], ("arr", op["off"][j], d[name]))
        yield pslot(6, "in_zp", d["in_zp"])
        yield pslot(7, "out_zp", d["out_zp"])
        yield pslot(8, "act_min", d["act_min"])
        yield pslot(9, "act_max", d["act_max"])
    elif k == _K_ADD and d["kind"] == "add":
        for i, name in enumerate(("in1_zp", "in2_zp", "out_zp", "left_shift", "m1", "s1", "m2", "s2", "mo", "so", "act_min", "act_max")):
            yield pslot(i, name, d[name])
    elif k == _K_MUL and d["kind"] == "mul":
        for i, name in enumerate(("in1_zp", "in2_zp", "out_zp", "mult", "shift", "act_min", "act_max")):
            yield pslot(i, name, d[name])
    elif k == _K_MEAN and d["kind"] == "mean":
        for i, name in enumerate(("count", "in_zp", "out_zp", "mult", "shift", "mult_n", "shift_n", "keep_dims")):
            yield pslot(i, name, d[name])
    elif k == _K_LOGI and d["kind"] == "logistic":
        have = np.frombuffer(blob, np.int8, 256, op["off"][0])
        yield (f"op {op['tfl']} LOGISTIC lut", have, d["lut"], ("arr", op["off"][0], d["lut"]))
    elif k == _K_REQ and d["kind"] == "requant":
        for i, name in enumerate(("in_zp", "out_zp", "mult", "shift")):
            yield pslot(i, name, d[name])
    elif k in (_K_QUANT, _K_DEQ) and d["kind"] in ("quantize", "dequantize"):
        yield pslot(0, "zp", d["zp"])
        yield (f"op {op['tfl']} scale", np.float32(op["f"][0]), np.float32(d["scale"]), ("f", op["f_at"], d["scale"]))


def compare_with_blob(blob: bytes, derived: dict[int, dict]) -> tuple[list[str], int]:
    """-> (mismatch descriptions, number of integer constants compared)."""
    bad, n = [], 0
    for op in _blob_ops(blob):
        d = derived.get(op["tfl"])
        if d is None:
            continue
        for what, have, want, _ in _fields(op, d, blob):
            n += int(np.size(want))
            if not np.array_equal(np.asarray(have), np.asarray(want)):
                bad.append(f"{what}: blob {have} != oracle {want}")
    return bad, n


def patch_blob(blob: bytes, derived: dict[int, dict]) -> bytes:
    """The same blob with every derived constant overwritten by the oracle's own value."""
    out = bytearray(blob)
    for op in _blob_ops(blob):
        d = derived.get(op["tfl"])
        if d is None:
            continue
        for _, _, _, (how, at, val) in _fields(op, d, blob):
            if how == "p":
                struct.pack_into("<i", out, at, val)
            elif how == "f":
                struct.pack_into("<f", out, at, val)
            else:
                raw = np.ascontiguousarray(val).tobytes()
                out[at:at + len(raw)] = raw
    return bytes(out)
