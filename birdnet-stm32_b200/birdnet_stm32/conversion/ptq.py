"""Post-training quantisation of small float graphs to int8 `.tflite` files, without TensorFlow.

The reference converts a Keras model with `tf.lite.TFLiteConverter` and a representative dataset
(`conversion/quantize.py:111-168`; float I/O kept, int8 inside, per-channel weights unless `--per_tensor`).
That converter cannot run here, so this module is the replacement for the op set the engine supports: a float graph
is described layer by layer (`FloatGraph`), evaluated with numpy on calibration inputs to record every activation's
range, and written out as an int8 TFLite graph with the conventions observed in the shipped checkpoint (SURVEY
Appendix H): symmetric int8 weights (per output channel or per tensor), int32 biases at `s_in * s_w`, asymmetric int8
activations from calibration min/max, fused ReLU/ReLU6 attributes, fixed 1/256 / -128 LOGISTIC and SOFTMAX outputs,
shared parameters across RESHAPE / TRANSPOSE / PAD, QUANTIZE first and DEQUANTIZE last.

`build_dscnn` assembles the reference architectures on top of it (`models/dscnn.py:198-262`, `models/blocks.py`,
`models/frontend.py:347-358`): plain DS blocks or inverted residuals, squeeze-excite, attention pooling, the
raw-waveform learned filterbank -- with random-init weights, which is all BASELINE configs 3 and 4 ask for.
"""

from __future__ import annotations

import math

import numpy as np

from birdnet_stm32.conversion.tflite_reader import Graph, OpInfo, TensorInfo
from birdnet_stm32.conversion.tflite_writer import write_tflite


# ---------------------------------------------------------------------------------------------------------------
# float graph + numpy evaluation
# ---------------------------------------------------------------------------------------------------------------
def _same_pad(n: int, k: int, s: int) -> tuple[int, int]:
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return out, total // 2


def _windows(x: np.ndarray, kh: int, kw: int, sh: int, sw: int, padding: str) -> np.ndarray:
    """NHWC -> [B, H', W', C, kh, kw] view of the (padded) input."""
    B, H, W, C = x.shape
    if padding == "SAME":
        oh, pt = _same_pad(H, kh, sh)
        ow, pl = _same_pad(W, kw, sw)
        pb = max((oh - 1) * sh + kh - H - pt, 0)
        pr = max((ow - 1) * sw + kw - W - pl, 0)
        x = np.pad(x, ((0, 0), (pt, pb), (pl, pr), (0, 0)))
    win = np.lib.stride_tricks.sliding_window_view(x, (kh, kw), axis=(1, 2))
    return win[:, ::sh, ::sw]


def _act(y: np.ndarray, act: str) -> np.ndarray:
    if act == "RELU":
        return np.maximum(y, 0.0)
    if act == "RELU6":
        return np.clip(y, 0.0, 6.0)
    return y


class FloatGraph:
    """A linear list of float layers over named tensors; `run` evaluates it with numpy (NHWC, float32)."""

    def __init__(self, input_shape: tuple):
        self.input_shape = tuple(input_shape)            # without batch
        self.layers: list[dict] = []
        self.shapes: dict[int, tuple] = {0: self.input_shape}
        self._n = 1

    def _add(self, kind: str, ins: list[int], shape: tuple, **kw) -> int:
        out = self._n
        self._n += 1
        self.layers.append(dict(kind=kind, ins=ins, out=out, **kw))
        self.shapes[out] = tuple(int(v) for v in shape)
        return out

    # -- layer constructors (shapes are per item, no batch dim) --------------------------------------------------
    def conv2d(self, x, w, b, stride=(1, 1), padding="SAME", act="NONE"):
        H, W, _ = self.shapes[x]
        co, kh, kw, _ = w.shape
        oh = _same_pad(H, kh, stride[0])[0] if padding == "SAME" else (H - kh) // stride[0] + 1
        ow = _same_pad(W, kw, stride[1])[0] if padding == "SAME" else (W - kw) // stride[1] + 1
        return self._add("CONV_2D", [x], (oh, ow, co), w=w.astype(np.float32), b=b.astype(np.float32), stride=stride, padding=padding, act=act)

    def dwconv(self, x, w, b, stride=(1, 1), act="NONE"):
        H, W, C = self.shapes[x]
        _, kh, kw, _ = w.shape
        return self._add("DEPTHWISE_CONV_2D", [x], (_same_pad(H, kh, stride[0])[0], _same_pad(W, kw, stride[1])[0], C),
                         w=w.astype(np.float32), b=b.astype(np.float32), stride=stride, padding="SAME", act=act)

    def dense(self, x, w, b=None, act="NONE"):
        shp = self.shapes[x]
        return self._add("FULLY_CONNECTED", [x], shp[:-1] + (w.shape[0],), w=w.astype(np.float32),
                         b=None if b is None else b.astype(np.float32), act=act)

    def mean_hw(self, x, keep_dims: bool):
        H, W, C = self.shapes[x]
        return self._add("MEAN", [x], (1, 1, C) if keep_dims else (C,), keep_dims=keep_dims)

    def logistic(self, x):
        return self._add("LOGISTIC", [x], self.shapes[x])

    def softmax(self, x, beta=1.0):
        return self._add("SOFTMAX", [x], self.shapes[x], beta=beta)

    def mul(self, a, b):
        return self._add("MUL", [a, b], self.shapes[a])

    def add(self, a, b, act="NONE"):
        return self._add("ADD", [a, b], self.shapes[a], act=act)

    def sum_axis(self, x, axis: int):
        shp = list(self.shapes[x])
        del shp[axis]
        return self._add("SUM", [x], tuple(shp), axis=axis)

    def reshape(self, x, shape):
        assert int(np.prod(shape)) == int(np.prod(self.shapes[x]))
        return self._add("RESHAPE", [x], tuple(shape))

    def transpose(self, x, perm):
        shp = self.shapes[x]
        return self._add("TRANSPOSE", [x], tuple(shp[p] for p in perm), perm=tuple(perm))

    def pad(self, x, pads):
        shp = self.shapes[x]
        return self._add("PAD", [x], tuple(s + a + b for s, (a, b) in zip(shp, pads)), pads=tuple(pads))

    # -- numpy forward ------------------------------------------------------------------------------------------
    def run(self, x0: np.ndarray) -> dict[int, np.ndarray]:
        v = {0: np.asarray(x0, dtype=np.float32)}
        for L in self.layers:
            k = L["kind"]
            a = v[L["ins"][0]]
            if k == "CONV_2D":
                win = _windows(a, L["w"].shape[1], L["w"].shape[2], *L["stride"], L["padding"])
                y = np.einsum("bhwcij,oijc->bhwo", win, L["w"], optimize=True) + L["b"]
                y = _act(y, L["act"])
            elif k == "DEPTHWISE_CONV_2D":
                win = _windows(a, L["w"].shape[1], L["w"].shape[2], *L["stride"], "SAME")
                y = np.einsum("bhwcij,ijc->bhwc", win, L["w"][0], optimize=True) + L["b"]
                y = _act(y, L["act"])
            elif k == "FULLY_CONNECTED":
                y = a @ L["w"].T
                if L["b"] is not None:
                    y = y + L["b"]
                y = _act(y, L["act"])
            elif k == "MEAN":
                y = a.mean(axis=(1, 2), keepdims=L["keep_dims"])
            elif k == "LOGISTIC":
                y = 1.0 / (1.0 + np.exp(-a))
            elif k == "SOFTMAX":
                z = L["beta"] * a
                z = np.exp(z - z.max(axis=-1, keepdims=True))
                y = z / z.sum(axis=-1, keepdims=True)
            elif k == "MUL":
                y = a * v[L["ins"][1]]
            elif k == "ADD":
                y = _act(a + v[L["ins"][1]], L["act"])
            elif k == "SUM":
                y = a.sum(axis=L["axis"] + 1)
            elif k == "RESHAPE":
                y = a.reshape((a.shape[0],) + self.shapes[L["out"]])
            elif k == "TRANSPOSE":
                y = a.transpose((0,) + tuple(p + 1 for p in L["perm"]))
            elif k == "PAD":
                y = np.pad(a, ((0, 0),) + L["pads"])
            else:
                raise ValueError(k)
            v[L["out"]] = y.astype(np.float32)
        return v


# ---------------------------------------------------------------------------------------------------------------
# calibration + quantisation -> TFLite graph
# ---------------------------------------------------------------------------------------------------------------
def _act_qparams(mn: float, mx: float) -> tuple[np.float32, int]:
    mn, mx = min(float(mn), 0.0), max(float(mx), 0.0)
    if mx - mn < 1e-12:
        return np.float32(1.0 / 255.0), -128
    scale = np.float32((mx - mn) / 255.0)
    zp = int(round(-128.0 - mn / float(scale)))
    return scale, max(-128, min(127, zp))


def _weight_q(w: np.ndarray, axis: int, per_channel: bool) -> tuple[np.ndarray, np.ndarray]:
    """Symmetric int8 in [-127, 127]; one scale per slice along `axis`, or one for the tensor."""
    if per_channel:
        red = tuple(i for i in range(w.ndim) if i != axis)
        amax = np.abs(w).max(axis=red)
    else:
        amax = np.array([np.abs(w).max()])
    scale = np.maximum(amax / 127.0, 1e-9).astype(np.float32)
    shp = [1] * w.ndim
    if per_channel:
        shp[axis] = -1
    q = np.clip(np.round(w / scale.reshape(shp).astype(np.float64)), -127, 127).astype(np.int8)
    return q, scale


def quantize_graph(fg: FloatGraph, calib: np.ndarray, per_channel: bool = True, batch: int = 8, description: str = "") -> Graph:
    """Calibrate `fg` on `calib` ([N, *input_shape] float32) and return the int8 TFLite :class:`Graph`."""
    lo: dict[int, float] = {}
    hi: dict[int, float] = {}
    for i in range(0, calib.shape[0], batch):
        for t, a in fg.run(calib[i:i + batch]).items():
            lo[t] = min(lo.get(t, np.inf), float(a.min()))
            hi[t] = max(hi.get(t, -np.inf), float(a.max()))

    qp: dict[int, tuple] = {}
    for L in [dict(kind="INPUT", ins=[], out=0)] + fg.layers:
        t = L["out"]
        if L["kind"] in ("LOGISTIC", "SOFTMAX"):
            qp[t] = (np.float32(1.0 / 256.0), -128)
        elif L["kind"] in ("RESHAPE", "TRANSPOSE", "PAD"):
            qp[t] = qp[L["ins"][0]]
        else:
            qp[t] = _act_qparams(lo[t], hi[t])

    tensors: list[TensorInfo] = []
    ops: list[OpInfo] = []

    def new_tensor(name, shape, dtype, scale=None, zp=None, qdim=0, data=None, batched=True):
        full = ((1,) + tuple(shape)) if batched else tuple(shape)
        sig = ((-1,) + tuple(shape)) if batched else tuple(shape)
        sc = np.zeros((0,), np.float32) if scale is None else np.atleast_1d(np.asarray(scale, dtype=np.float32))
        z = np.zeros((0,), np.int64) if zp is None else np.atleast_1d(np.asarray(zp, dtype=np.int64))
        tensors.append(TensorInfo(len(tensors), name, full, sig, dtype, sc, z, qdim, data))
        return len(tensors) - 1

    def new_op(kind, ins, outs, **options):
        ops.append(OpInfo(len(ops), kind, 1, list(ins), list(outs), dict(options)))

    tid: dict[int, int] = {}
    t_in = new_tensor("serving_default_input:0", fg.input_shape, np.float32)
    tid[0] = new_tensor("input_int8", fg.input_shape, np.int8, *qp[0])
    new_op("QUANTIZE", [t_in], [tid[0]])

    for li, L in enumerate(fg.layers):
        k, out = L["kind"], L["out"]
        x = tid[L["ins"][0]]
        s_in = float(qp[L["ins"][0]][0])
        y = new_tensor(f"l{li}_{k.lower()}", fg.shapes[out], np.int8, *qp[out])
        tid[out] = y
        if k in ("CONV_2D", "DEPTHWISE_CONV_2D", "FULLY_CONNECTED"):
            w = L["w"]
            axis = 3 if k == "DEPTHWISE_CONV_2D" else 0
            wq, ws = _weight_q(w, axis, per_channel)
            tw = new_tensor(f"l{li}_w", w.shape, np.int8, ws, np.zeros(ws.size, np.int64), axis if ws.size > 1 else 0, wq, batched=False)
            ins = [x, tw]
            nout = w.shape[axis]
            if L["b"] is not None:
                bs = (np.float64(s_in) * ws.astype(np.float64)).astype(np.float32)
                bs_full = np.broadcast_to(bs, (nout,)) if bs.size == 1 else bs
                bq = np.clip(np.round(L["b"].astype(np.float64) / bs_full.astype(np.float64)), -(1 << 30), 1 << 30).astype(np.int32)
                ins.append(new_tensor(f"l{li}_b", (nout,), np.int32, bs, np.zeros(bs.size, np.int64), 0, bq, batched=False))
            if k == "FULLY_CONNECTED":
                new_op(k, ins, [y], act=L["act"], weights_format=0, keep_num_dims=len(fg.shapes[out]) > 1)
            elif k == "CONV_2D":
                new_op(k, ins, [y], padding=L["padding"], stride_w=L["stride"][1], stride_h=L["stride"][0], act=L["act"], dil_w=1, dil_h=1)
            else:
                new_op(k, ins, [y], padding="SAME", stride_w=L["stride"][1], stride_h=L["stride"][0], depth_multiplier=1, act=L["act"], dil_w=1, dil_h=1)
        elif k in ("MEAN", "SUM"):
            axes = np.array([1, 2], np.int32) if k == "MEAN" else np.array([L["axis"] + 1], np.int32)
            ta = new_tensor(f"l{li}_axes", axes.shape, np.int32, data=axes, batched=False)
            new_op(k, [x, ta], [y], keep_dims=bool(L.get("keep_dims", False)))
        elif k == "LOGISTIC":
            new_op(k, [x], [y])
        elif k == "SOFTMAX":
            new_op(k, [x], [y], beta=float(L["beta"]))
        elif k in ("MUL", "ADD"):
            new_op(k, [x, tid[L["ins"][1]]], [y], act=L.get("act", "NONE"))
        elif k == "RESHAPE":
            shp = np.array((-1,) + fg.shapes[out], np.int32)
            ts = new_tensor(f"l{li}_shape", shp.shape, np.int32, data=shp, batched=False)
            new_op(k, [x, ts], [y], new_shape=tuple(int(v) for v in shp))
        elif k == "TRANSPOSE":
            perm = np.array((0,) + tuple(p + 1 for p in L["perm"]), np.int32)
            tp = new_tensor(f"l{li}_perm", perm.shape, np.int32, data=perm, batched=False)
            new_op(k, [x, tp], [y])
        elif k == "PAD":
            pads = np.array(((0, 0),) + L["pads"], np.int32)
            tp = new_tensor(f"l{li}_paddings", pads.shape, np.int32, data=pads, batched=False)
            new_op(k, [x, tp], [y])
        else:
            raise ValueError(k)

    last = fg.layers[-1]["out"]
    t_out = new_tensor("StatefulPartitionedCall:0", fg.shapes[last], np.float32)
    new_op("DEQUANTIZE", [tid[last]], [t_out])
    return Graph(tensors=tensors, ops=ops, inputs=[t_in], outputs=[t_out], description=description or "birdnet_stm32 B200 PTQ")


# ---------------------------------------------------------------------------------------------------------------
# reference architectures with random-init weights
# ---------------------------------------------------------------------------------------------------------------
def _make_divisible(v, divisor: int = 8) -> int:
    """`models/blocks.py:13-24`"""
    v = int(v + divisor / 2) // divisor * divisor
    return max(divisor, v)


class _Init:
    """He-style random weights with BatchNorm folded away (random-init BN is the identity plus a small bias)."""

    def __init__(self, seed: int):
        self.rng = np.random.default_rng(seed)

    def conv(self, co, kh, kw, ci):
        return self.rng.normal(0.0, math.sqrt(2.0 / (kh * kw * ci)), (co, kh, kw, ci)), self.rng.normal(0.0, 0.05, co)

    def dw(self, kh, kw, c):
        return self.rng.normal(0.0, math.sqrt(2.0 / (kh * kw)), (1, kh, kw, c)), self.rng.normal(0.0, 0.05, c)

    def dense(self, n, k, bias=True):
        return self.rng.normal(0.0, math.sqrt(1.0 / k), (n, k)), (self.rng.normal(0.0, 0.05, n) if bias else None)


def _se(fg: FloatGraph, init: _Init, x: int, reduction: int) -> int:
    """`models/blocks.py:27-46`: GAP(keepdims) -> Dense(relu, no bias) -> Dense(sigmoid, no bias) -> Multiply"""
    C = fg.shapes[x][-1]
    se_ch = max(1, C // reduction)
    s = fg.mean_hw(x, keep_dims=True)
    s = fg.dense(s, init.dense(se_ch, C, bias=False)[0], None, act="RELU")
    s = fg.logistic(fg.dense(s, init.dense(C, se_ch, bias=False)[0], None))
    return fg.mul(x, s)


def build_dscnn(frontend: str = "hybrid_features", num_mels: int = 64, spec_width: int = 256, chunk_len: int = 48000,
                alpha: float = 1.0, depth_multiplier: float = 1.0, num_classes: int = 10, use_se: bool = False,
                se_reduction: int = 8, use_inverted_residual: bool = False, expansion_factor: int = 2,
                use_attention_pooling: bool = False, embeddings_size: int = 256, seed: int = 7) -> FloatGraph:
    """Float graph of `build_dscnn_model` (`models/dscnn.py:198-262`) as it lowers to TFLite.

    frontend "raw": input `[T, 1]`, learned filterbank `Conv1D(num_mels, 16, stride=ceil(T / spec_width))` + ReLU6
    written as the CONV_2D the converter emits, zero padded at the end when T is not a multiple
    (`models/frontend.py:139-147,347-358`).  frontend "precomputed": input `[num_mels, spec_width, 1]`.
    """
    init = _Init(seed)
    if frontend == "raw":
        fg = FloatGraph((chunk_len, 1))
        stride = -(-chunk_len // spec_width)
        k = 16
        need = (spec_width - 1) * stride + k
        x = fg.reshape(0, (1, chunk_len, 1))
        if need > chunk_len:
            x = fg.pad(x, ((0, 0), (0, need - chunk_len), (0, 0)))
        w, b = init.conv(num_mels, 1, k, 1)
        x = fg.conv2d(x, w, b, stride=(1, stride), padding="VALID", act="RELU6")       # [1, W, mels]
        x = fg.transpose(x, (2, 1, 0))                                                  # [mels, W, 1]
    else:
        fg = FloatGraph((num_mels, spec_width, 1))
        x = 0
    stem_ch = _make_divisible(int(16 * alpha))
    w, b = init.conv(stem_ch, 3, 3, 1)
    x = fg.conv2d(x, w, b, stride=(1, 2), act="RELU6")
    for si, (bf, br) in enumerate(zip([32, 64, 128, 256], [2, 3, 4, 2])):
        out_ch = _make_divisible(int(bf * alpha))
        reps = max(1, int(math.ceil(br * depth_multiplier)))
        for bi in range(reps):
            stride = (2, 2) if bi == 0 else (1, 1)
            in_ch = fg.shapes[x][-1]
            if use_inverted_residual:
                # models/blocks.py:83-133: expand 1x1 (ReLU6) -> DW 3x3 (ReLU6) -> [SE] -> project 1x1 (linear) -> ADD (no act)
                hidden = _make_divisible(in_ch * expansion_factor)
                h = fg.conv2d(x, *init.conv(hidden, 1, 1, in_ch), act="RELU6")
                h = fg.dwconv(h, *init.dw(3, 3, hidden), stride=stride, act="RELU6")
                if use_se:
                    h = _se(fg, init, h, se_reduction)
                h = fg.conv2d(h, *init.conv(out_ch, 1, 1, hidden))
                x = fg.add(x, h) if stride == (1, 1) and in_ch == out_ch else h
            else:
                # models/dscnn.py:28-84: DW 3x3 (ReLU6) -> PW 1x1 (ReLU6 | linear + ADD + ReLU6); SE after the block
                h = fg.dwconv(x, *init.dw(3, 3, in_ch), stride=stride, act="RELU6")
                if stride == (1, 1) and in_ch == out_ch:
                    h = fg.conv2d(h, *init.conv(out_ch, 1, 1, in_ch))
                    x = fg.add(x, h, act="RELU6")
                else:
                    x = fg.conv2d(h, *init.conv(out_ch, 1, 1, in_ch), act="RELU6")
                if use_se:
                    x = _se(fg, init, x, se_reduction)
    emb = _make_divisible(int(embeddings_size))
    if fg.shapes[x][-1] != emb:
        x = fg.conv2d(x, *init.conv(emb, 1, 1, fg.shapes[x][-1]), act="RELU6")
    H, W, C = fg.shapes[x]
    if use_attention_pooling:
        # models/blocks.py:151-159: reshape [HW, C] -> Dense(1) -> softmax over HW -> multiply -> reduce_sum
        flat = fg.reshape(x, (H * W, C))
        a = fg.dense(flat, init.dense(1, C, bias=False)[0], None)          # [HW, 1]
        a = fg.softmax(fg.reshape(a, (1, H * W)))                           # softmax runs over the last dim in TFLite
        a = fg.reshape(a, (H * W, 1))
        x = fg.sum_axis(fg.mul(flat, a), 0)                                 # [C]
    else:
        x = fg.mean_hw(x, keep_dims=False)
    w, b = init.dense(num_classes, C)
    fg.logistic(fg.dense(x, w, b))
    return fg


def synth_calibration(fg: FloatGraph, n: int, seed: int = 11) -> np.ndarray:
    """Synthetic calibration inputs in the value range of the frontend (`cli/convert.py:123-141` falls back to
    uniform-random spectrograms when no data is given): raw waveforms in [-1, 1] scaled by their peak, spectrograms in [0, 1]."""
    rng = np.random.default_rng(seed)
    if len(fg.input_shape) == 2:                          # raw [T, 1]
        T = fg.input_shape[0]
        t = np.arange(T) / 24000.0
        x = np.zeros((n, T, 1), np.float32)
        for i in range(n):
            f0, f1 = rng.uniform(300, 9000, 2)
            y = rng.uniform(0.2, 0.9) * np.sin(2 * np.pi * (f0 * t + (f1 - f0) * t * t / (2 * t[-1]))) + rng.uniform(0.02, 0.2) * rng.standard_normal(T)
            x[i, :, 0] = y / (np.abs(y).max() + 1e-6)
        return x
    return (rng.random((n,) + fg.input_shape, dtype=np.float32) ** 2).astype(np.float32)


def convert(fg: FloatGraph, calib: np.ndarray, per_channel: bool = True, description: str = "") -> bytes:
    """Float graph + calibration data -> `.tflite` bytes (the `convert_to_tflite` twin, `quantize.py:111-168`)."""
    return write_tflite(quantize_graph(fg, calib, per_channel=per_channel, description=description))
