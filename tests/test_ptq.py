"""BASELINE configs 3 and 4: the TensorFlow-free `.tflite` writer and PTQ calibrator, and
the graphs they produce -- raw-waveform learned-filterbank DS-CNN, wide DS-CNN (alpha 1.0, depth multiplier 2) with
squeeze-excite and attention pooling, plain-DS and inverted-residual forms, per-tensor vs per-channel weights.

CPU tests: writer round trip, PTQ conventions (Appendix H), oracle int8 result vs the float model.
GPU tests: every int8 tensor of the generic CUDA plan == the oracle, bit for bit, through the C ABI.
"""

import json
import os

import numpy as np
import pytest

from birdnet_stm32.conversion import ptq
from birdnet_stm32.conversion.export_blob import export_blob
from birdnet_stm32.conversion.tflite_reader import read_tflite
from birdnet_stm32.conversion.tflite_writer import write_tflite

HERE = os.path.dirname(os.path.abspath(__file__))

RAW_CFG = dict(audio_frontend="raw", sample_rate=24000, chunk_duration=2, spec_width=256, fft_length=512, num_mels=64)
RAW_PAD_CFG = dict(audio_frontend="raw", sample_rate=22050, chunk_duration=2, spec_width=256, fft_length=512, num_mels=64)
MEL_CFG = dict(audio_frontend="librosa", sample_rate=24000, chunk_duration=3, spec_width=256, fft_length=512, num_mels=64)

CASES = {
    # config 3: T = 48,000 (stride 188, no padding) and T = 44,100 (stride 173, 31 zero samples appended)
    "raw_48000": (dict(frontend="raw", chunk_len=48000, num_classes=10, seed=7), RAW_CFG, True),
    "raw_44100_padded": (dict(frontend="raw", chunk_len=44100, num_classes=10, seed=7), RAW_PAD_CFG, True),
    # config 4: alpha 1.0, depth_multiplier 2 (repeats 4, 6, 8, 4), SE reduction 8, attention pooling
    "wide_se_attn_per_channel": (dict(frontend="precomputed", depth_multiplier=2.0, use_se=True, use_attention_pooling=True,
                                      num_classes=12, seed=7), MEL_CFG, True),
    "wide_se_attn_per_tensor": (dict(frontend="precomputed", depth_multiplier=2.0, use_se=True, use_attention_pooling=True,
                                     num_classes=12, seed=7), MEL_CFG, False),
    "wide_ir_se_attn": (dict(frontend="precomputed", depth_multiplier=2.0, use_se=True, use_inverted_residual=True,
                             use_attention_pooling=True, num_classes=12, seed=8), MEL_CFG, True),
}
_cache: dict = {}


def _case(name):
    if name not in _cache:
        kw, cfg, per_channel = CASES[name]
        fg = ptq.build_dscnn(**kw)
        raw = ptq.convert(fg, ptq.synth_calibration(fg, 6), per_channel=per_channel)
        g = read_tflite(raw)
        _cache[name] = (fg, raw, g, export_blob(g, cfg))
    return _cache[name]


# ---------------------------------------------------------------------------------------------------------------
# CPU
# ---------------------------------------------------------------------------------------------------------------
def test_writer_round_trips_the_shipped_checkpoint(graph, cfg, blob):
    raw = write_tflite(graph)
    assert raw[4:8] == b"TFL3"
    g2 = read_tflite(raw)
    assert len(g2.ops) == len(graph.ops) and len(g2.tensors) == len(graph.tensors)
    for a, b in zip(graph.ops, g2.ops):
        assert (a.kind, a.inputs, a.outputs, a.options) == (b.kind, b.inputs, b.outputs, b.options)
    for a, b in zip(graph.tensors, g2.tensors):
        assert (a.name, a.shape, a.shape_signature, a.dtype, a.quantized_dimension) == (b.name, b.shape, b.shape_signature, b.dtype, b.quantized_dimension)
        assert np.array_equal(a.scale, b.scale) and np.array_equal(a.zero_point, b.zero_point)
        assert (a.data is None) == (b.data is None) and (a.data is None or np.array_equal(a.data, b.data))
    assert export_blob(g2, cfg) == blob
    # and once more: writing what was read back is a fixed point
    assert write_tflite(g2) == raw


def test_ptq_follows_the_conventions_of_the_shipped_file():
    fg, raw, g, _ = _case("wide_se_attn_per_channel")
    assert g.ops[0].kind == "QUANTIZE" and g.ops[-1].kind == "DEQUANTIZE"
    assert g.tensor(g.inputs[0]).dtype == np.float32 and g.tensor(g.outputs[0]).dtype == np.float32
    assert g.tensor(g.inputs[0]).shape_signature[0] == -1
    kinds = {op.kind for op in g.ops}
    assert {"CONV_2D", "DEPTHWISE_CONV_2D", "FULLY_CONNECTED", "MEAN", "LOGISTIC", "MUL", "ADD", "SOFTMAX", "SUM", "RESHAPE"} <= kinds
    for op in g.ops:
        if op.kind in ("CONV_2D", "DEPTHWISE_CONV_2D", "FULLY_CONNECTED"):
            x, w = g.tensor(op.inputs[0]), g.tensor(op.inputs[1])
            nout = w.shape[3] if op.kind == "DEPTHWISE_CONV_2D" else w.shape[0]
            assert w.dtype == np.int8 and not w.zero_point.any() and np.abs(w.data.astype(int)).max() <= 127
            assert w.scale.size == nout and w.quantized_dimension == (3 if op.kind == "DEPTHWISE_CONV_2D" else 0)
            if len(op.inputs) > 2:
                b = g.tensor(op.inputs[2])
                assert b.dtype == np.int32 and np.allclose(b.scale, x.s() * w.scale, rtol=1e-6)
        if op.kind in ("LOGISTIC", "SOFTMAX"):
            y = g.tensor(op.outputs[0])
            assert y.s() == np.float32(1 / 256) and y.zp() == -128
        if op.kind in ("RESHAPE", "TRANSPOSE", "PAD"):
            x, y = g.tensor(op.inputs[0]), g.tensor(op.outputs[0])
            assert (x.s(), x.zp()) == (y.s(), y.zp())
    # --per_tensor: one scale per weight tensor (conversion/quantize.py:140-141)
    _, _, gt, _ = _case("wide_se_attn_per_tensor")
    for op in gt.ops:
        if op.kind in ("CONV_2D", "DEPTHWISE_CONV_2D", "FULLY_CONNECTED"):
            assert gt.tensor(op.inputs[1]).scale.size == 1
    # the raw frontend lowers to (PAD) RESHAPE CONV_2D[64,1,16,1] stride (1, ceil(T/256)) VALID + ReLU6, TRANSPOSE
    _, _, gr, _ = _case("raw_44100_padded")
    assert [op.kind for op in gr.ops[:6]] == ["QUANTIZE", "RESHAPE", "PAD", "CONV_2D", "TRANSPOSE", "CONV_2D"]
    conv = gr.ops[3]
    assert gr.tensor(conv.inputs[1]).shape == (64, 1, 16, 1) and conv.options["stride_w"] == 173 and conv.options["padding"] == "VALID"
    assert gr.tensor(conv.outputs[0]).shape == (1, 1, 256, 64)
    _, _, g48, _ = _case("raw_48000")
    assert "PAD" not in {op.kind for op in g48.ops} and g48.ops[2].options["stride_w"] == 188


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_int8_graph_tracks_the_float_model(name):
    from oracle import bn_oracle

    fg, _, _, blob = _case(name)
    m = bn_oracle.OracleModel(blob)
    x = ptq.synth_calibration(fg, 4, seed=99)
    yq = m.predict(x)
    yf = fg.run(x)[fg.layers[-1]["out"]]
    assert yq.shape == yf.shape
    # int8 quantisation noise through 40 - 200 layers; scores are sigmoid outputs in [0, 1]
    assert np.abs(yq - yf).max() < 0.08, np.abs(yq - yf).max()
    assert np.mean(np.abs(yq - yf)) < 0.02


def test_softmax_sum_pad_oracle_semantics():
    """Known answers for the three ops configs 3/4 add to the op set (int8 kernels of TFLite's builtin resolver)."""
    from oracle import bn_oracle

    # attention pooling alone on a tiny tensor: [HW=4, C=2]
    fg = ptq.FloatGraph((2, 2, 2))
    flat = fg.reshape(0, (4, 2))
    w = np.array([[1.0, -0.5]], np.float32)
    a = fg.softmax(fg.reshape(fg.dense(flat, w, None), (1, 4)))
    fg.sum_axis(fg.mul(flat, fg.reshape(a, (4, 1))), 0)
    rng = np.random.default_rng(0)
    cal = rng.random((512, 2, 2, 2), dtype=np.float32)
    g = ptq.quantize_graph(fg, cal)
    m = bn_oracle.OracleModel(export_blob(read_tflite(write_tflite(g)), {}))
    x = cal[:5]                                          # inside the calibrated ranges: no saturation
    want = fg.run(x)[fg.layers[-1]["out"]]
    got = m.predict(x)
    assert np.abs(got - want).max() < 0.03
    # softmax tap: rows sum to ~1 (in units of 1/256) and follow the float softmax within 1.5 LSB
    sm_id = [op.outputs[0] for op in g.ops if op.kind == "SOFTMAX"][0]
    _, tap = m.run(x, tap_id=sm_id)
    p = (tap.astype(np.int32) + 128) / 256.0
    assert np.abs(p.sum(axis=1) - 1.0).max() <= 3 / 256
    sm_f = fg.run(x)[[L["out"] for L in fg.layers if L["kind"] == "SOFTMAX"][0]]
    assert np.abs(p - sm_f.reshape(5, 4)).max() <= 0.03
    # PAD writes the zero point
    fp = ptq.FloatGraph((1, 6, 1))
    fp.reshape(fp.pad(0, ((0, 0), (0, 3), (0, 0))), (9,))      # scores are read as [B, last dim]
    calp = rng.random((64, 1, 6, 1), dtype=np.float32) + 0.5
    gp = ptq.quantize_graph(fp, calp)
    mp = bn_oracle.OracleModel(export_blob(gp, {}))
    xp = calp[:2]
    yp = mp.predict(xp)
    assert yp.shape == (2, 9) and np.abs(yp[:, 6:]).max() < 1e-6 and np.abs(yp[:, :6] - xp[:, 0, :, 0]).max() < 0.01


# ---------------------------------------------------------------------------------------------------------------
# GPU: generic CUDA plan == oracle, every tensor
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_gpu_generic_plan_bit_exact_on_synthesised_graphs(name):
    from birdnet_stm32.evaluation.gpu_runner import GpuRunner
    from oracle import bn_oracle

    fg, _, g, blob = _case(name)
    _, cfg, _ = CASES[name]
    B = 3
    x = ptq.synth_calibration(fg, B, seed=123)
    m = bn_oracle.OracleModel(blob)
    runner = GpuRunner(blob, cfg)
    assert runner.query().fast_path == 0          # these topologies run on the one-kernel-per-op plan
    got = runner.predict(x)
    want = m.predict(x)
    assert got.dtype == np.float32 and got.shape == want.shape
    assert np.array_equal(got, want), np.abs(got - want).max()
    # every int8 activation tensor, via the debug taps.  With the default BN_OPT_FUSION a 1x1 convolution whose only consumer is
    # the ADD right behind it runs fused with that ADD (its own output is not materialised); with BN_OPT_FUSION = 0 every op
    # writes its output.  The SE gate (MEAN -> FC -> FC -> LOGISTIC as one launch) still writes all four tensors.
    from birdnet_stm32 import _lib as L

    users = {}
    for op in g.ops:
        for t in op.inputs:
            users[t] = users.get(t, 0) + 1
    add_inputs = {t for op in g.ops if op.kind == "ADD" for t in op.inputs[:2]}
    folded = {op.outputs[0] for op in g.ops if op.kind == "CONV_2D" and users.get(op.outputs[0], 0) == 1 and op.outputs[0] in add_inputs}
    # ... and a depthwise 3x3 whose only consumer is a 1x1 convolution runs inside the fused DS-block kernel with it
    conv_inputs = {op.inputs[0] for op in g.ops if op.kind == "CONV_2D"}
    folded |= {op.outputs[0] for op in g.ops if op.kind == "DEPTHWISE_CONV_2D" and users.get(op.outputs[0], 0) == 1 and op.outputs[0] in conv_inputs}
    for fusion in (None, 0):
        if fusion is not None:
            runner.set_option(L.BN_OPT_FUSION, fusion)
            assert np.array_equal(runner.predict(x), want)
        checked = 0
        for op in g.ops:
            t = g.tensor(op.outputs[0])
            if t.dtype != np.int8 or (fusion is None and t.index in folded):
                continue
            n = int(np.prod(t.shape[1:]))
            _, tap = m.run(x, tap_id=t.index)
            dev = runner.dump_tensor(t.index, B * n)
            assert np.array_equal(dev.reshape(B, n), tap.reshape(B, n)), (name, op.index, op.kind, fusion)
            checked += 1
        assert checked >= len(g.ops) - 2 - (len(folded) if fusion is None else 0)
    runner.set_option(L.BN_OPT_FUSION, 139)
    runner.close()


@pytest.mark.gpu
def test_gpu_generic_plan_fusions_follow_the_rounding_and_mean_options():
    """The single-launch SE gate (MEAN -> FC -> FC -> LOGISTIC) and the other fast kernels of the generic plan under the
    non-default TFLite builds the engine can be told to match (single rounding; MEAN variants 1 and 3): scores equal to the
    oracle with the same options, with the multi-op launches on (default) and off."""
    from birdnet_stm32 import _lib as L
    from birdnet_stm32.evaluation.gpu_runner import GpuRunner
    from oracle import bn_oracle

    name = "wide_se_attn_per_channel"
    fg, _, g, blob = _case(name)
    _, cfg, _ = CASES[name]
    x = ptq.synth_calibration(fg, 5, seed=321)
    runner = GpuRunner(blob, cfg)
    try:
        for rounding, variant in ((0, 0), (1, 0), (0, 1), (0, 3), (1, 2)):
            want = bn_oracle.OracleModel(blob, rounding=rounding, mean_variant=variant).predict(x)
            runner.set_option(L.BN_OPT_ROUNDING, rounding)
            runner.set_option(L.BN_OPT_MEAN_VARIANT, variant)
            for fusion in (139, 0):
                runner.set_option(L.BN_OPT_FUSION, fusion)
                got = runner.predict(x)
                assert np.array_equal(got, want), (rounding, variant, fusion, np.abs(got - want).max())
    finally:
        runner.close()


def test_ptq_weight_quantiser_reproduces_the_real_converter_output(graph):
    """Row f3 against the reference's own artefacts: the weight quantiser of `conversion/ptq.py`, fed with the
    BatchNorm-folded float weights of the shipped Keras checkpoint, must give the int8 weights and scales that the REAL
    TensorFlow Lite converter wrote into the shipped `.tflite` (build container only: needs the reference checkout)."""
    import sys

    path = "/root/reference/checkpoints/birdnet_stm32n6_100.keras"
    if not os.path.exists(path):
        pytest.skip("reference checkout not mounted")
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"))
    from make_golden_weights import expected_layers

    from birdnet_stm32.conversion import ptq
    from oracle.keras_float_model import KerasFloatModel

    layers = expected_layers(KerasFloatModel(path))
    ops = [op for op in graph.ops if op.kind in ("CONV_2D", "DEPTHWISE_CONV_2D", "FULLY_CONNECTED")
           and not (op.kind == "DEPTHWISE_CONV_2D" and graph.tensors[op.inputs[1]].shape[1:3] == (1, 1))]
    assert len(layers) == len(ops) == 25
    same = total = 0
    for (name, wf, axis, _), op in zip(layers, ops):
        wt = graph.tensors[op.inputs[1]]
        q, scale = ptq._weight_q(wf.astype(np.float32), axis, per_channel=True)
        live = scale > 4e-9                                   # dead channels (|w| ~ 1e-40) carry converter floor scales
        shp = [1] * q.ndim
        shp[axis] = -1
        m = np.broadcast_to(live.reshape(shp), q.shape)
        assert np.array_equal(q[m], wt.data[m]), name
        assert np.all(np.abs(scale[live].astype(np.float64) - wt.scale.astype(np.float64)[live]) <= 1e-6 * scale[live]), name
        same += int(m.sum())
        total += q.size
    assert same > 200000 and same / total > 0.9            # the rest are dead channels (6.5 % of the weights)


def test_own_ptq_of_the_shipped_float_checkpoint_passes_the_reference_gate(synth):
    """Row f3 end to end on the reference's real network: the body of the shipped Keras checkpoint (stem ... dense head,
    BatchNorm folded) is described as a FloatGraph, quantised by `conversion/ptq.py` on 24 calibration chunks, written as a
    `.tflite`, exported and run by the int8 oracle; its scores must pass the reference's conversion gate against the float
    model (mean cosine >= 0.95, `conversion/validate.py`) on 12 held-out chunks -- like the shipped converter output does
    (build container only: needs the reference checkout)."""
    path = "/root/reference/checkpoints/birdnet_stm32n6_100.keras"
    if not os.path.exists(path):
        pytest.skip("reference checkout not mounted")
    from conftest import CONFIG

    from oracle import bn_oracle
    from oracle.keras_float_model import KerasFloatModel, cosine_similarity

    km = KerasFloatModel(path)
    cfg = json.load(open(CONFIG))
    T = int(cfg["sample_rate"] * cfg["chunk_duration"])

    def frontend_out(seed, n):
        pcm = synth.synth_pcm16(n, T, cfg["sample_rate"], seed=seed, edge_cases=False)
        spec = bn_oracle.frontend_hybrid(pcm, synth.file_peaks(pcm), cfg["fft_length"], T // cfg["spec_width"], cfg["spec_width"])
        taps: dict = {}
        scores = km.predict(spec, taps)
        return taps["frontend"].astype(np.float32), scores          # [n, 64, 256, 1] NHWC

    fg = ptq.FloatGraph((64, 256, 1))
    x, block_in = 0, None
    layers = km.layers
    for i, l in enumerate(layers):
        cls, c = l["class_name"], l["config"]
        if cls in ("Conv2D", "DepthwiseConv2D"):
            bn = layers[i + 1]
            g, b, m, v = (km.var(bn["config"]["name"], k).astype(np.float64) for k in range(4))
            sc = g / np.sqrt(v + bn["config"]["epsilon"])
            nxt = next(t for t in layers[i + 2:] if t["class_name"] not in ("SpatialDropout2D",))
            act = "RELU6" if nxt["class_name"] == "ReLU" else "NONE"          # a residual Add comes before the ReLU
            w = km.var(c["name"], 0).astype(np.float64)
            if cls == "Conv2D":
                x = fg.conv2d(x, np.transpose(w * sc[None, None, None, :], (3, 0, 1, 2)), b - m * sc, stride=tuple(c["strides"]), act=act)
            else:
                block_in = x
                x = fg.dwconv(x, (w[:, :, :, 0] * sc[None, None, :])[None], b - m * sc, stride=tuple(c["strides"]), act=act)
        elif cls == "Add":
            x = fg.add(block_in, x, act="RELU6")
        elif cls == "GlobalAveragePooling2D":
            x = fg.mean_hw(x, keep_dims=False)
        elif cls == "Dense":
            x = fg.logistic(fg.dense(x, km.var(c["name"], 0).T, km.var(c["name"], 1)))
    calib, _ = frontend_out(101, 24)
    test_x, want = frontend_out(202, 12)
    np.testing.assert_allclose(fg.run(test_x)[x], want, atol=2e-5)       # the FloatGraph IS the float model's body
    blob = export_blob(read_tflite(ptq.convert(fg, calib, per_channel=True, description="shipped checkpoint, own PTQ")), {})
    got = bn_oracle.OracleModel(blob).predict(test_x)
    cos = [cosine_similarity(want[i].astype(np.float64), got[i].astype(np.float64)) for i in range(len(want))]
    print(f"own PTQ of the shipped float checkpoint: cosine mean {np.mean(cos):.4f}, min {min(cos):.4f}, MAE {np.abs(got - want).mean():.5f}")
    assert np.mean(cos) >= 0.95 and np.abs(got - want).mean() <= 0.01
