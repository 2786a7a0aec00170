"""Pin of the int8 path against the reference's shipped FLOAT checkpoint, with the reference's own conversion gate.

`tests/golden/keras_float_reference.npz` holds the sigmoid scores of `checkpoints/birdnet_stm32n6_100.keras` (the Keras
model the shipped `.tflite` was converted from) on the seeded 16-chunk batch, computed from the archive's float weights and
BatchNorm statistics (`tests/golden/make_golden_keras.py`).  The reference accepts a conversion when the int8 outputs and
the Keras outputs have mean cosine similarity >= 0.95 (`conversion/validate.py:51-105`, `cli/convert.py:187-195`); the same
gate is applied here to the CPU oracle and to the CUDA engine.  It is an independent route to the network output (no
quantisation parameters, no folded weights), so it checks the reading of the `.tflite` -- layouts, padding, per-channel
axes, residual wiring, the PWL frontend -- that the bit-exact tests take for granted.

The last four chunks of the batch are the synthetic edge cases (silence, full-scale square, impulse, pure sine): far outside
the training distribution, all scores near zero, where cosine similarity is ill-conditioned (the reference's own rule gives
0.0 when exactly one side is all-zero).  They are reported but the gate is evaluated on the 12 chirp + noise chunks.
"""

import json
import os

import numpy as np
import pytest

from conftest import CONFIG, GOLDEN

N_GATED = 12


def cosine(a, b, eps=1e-8):
    an, bn = np.linalg.norm(a), np.linalg.norm(b)
    if an < eps and bn < eps:
        return 1.0
    if an < eps or bn < eps:
        return 0.0
    return float(np.dot(a, b) / (an * bn))


def batch(synth):
    cfg = json.load(open(CONFIG))
    T = int(cfg["sample_rate"] * cfg["chunk_duration"])
    pcm = synth.synth_pcm16(16, T, cfg["sample_rate"], seed=1234, edge_cases=True)
    return cfg, T, pcm, synth.file_peaks(pcm)


def gate(scores, ref):
    cos = [cosine(ref[i].astype(np.float64), scores[i].astype(np.float64)) for i in range(len(ref))]
    return cos, float(np.mean(cos[:N_GATED])), float(np.mean(np.abs(scores[:N_GATED] - ref[:N_GATED])))


def test_int8_oracle_passes_the_reference_conversion_gate(synth, oracle_model):
    from oracle import bn_oracle

    ref = np.load(os.path.join(GOLDEN, "keras_float_reference.npz"))
    cfg, T, pcm, peak = batch(synth)
    spec = bn_oracle.frontend_hybrid(pcm, peak, cfg["fft_length"], T // cfg["spec_width"], cfg["spec_width"])
    np.testing.assert_allclose(spec.astype(np.float64).sum(axis=(1, 2, 3)), ref["spec_sum"], rtol=1e-6)   # same inputs as the golden run
    scores = oracle_model.predict(spec)
    cos, cos_mean, mae = gate(scores, ref["scores"])
    print(f"int8 oracle vs Keras float: cosine mean {cos_mean:.4f} (min {min(cos[:N_GATED]):.4f}), MAE {mae:.5f}; edge cases {np.round(cos[N_GATED:], 3)}")
    assert cos_mean >= 0.95                       # the reference's default --min_cosine_sim
    assert min(cos[:N_GATED]) >= 0.95
    assert mae <= 0.01
    agree = int((scores[:N_GATED].argmax(1) == ref["scores"][:N_GATED].argmax(1)).sum())
    assert agree >= N_GATED - 2, agree            # top-1 may flip between near-tied classes after quantisation


def test_keras_archive_reader_when_reference_is_mounted():
    """Re-derives the golden scores from the archive itself (build container only; skipped on the GPU box)."""
    path = "/root/reference/checkpoints/birdnet_stm32n6_100.keras"
    if not os.path.exists(path):
        pytest.skip("reference checkout not mounted")
    from oracle.h5min import read_keras_weights

    cfg, w = read_keras_weights(path)
    layer_w = {k: v for k, v in w.items() if k.startswith("/layers/")}
    ref = np.load(os.path.join(GOLDEN, "keras_float_reference.npz"))
    assert sum(v.size for v in layer_w.values()) == int(ref["n_layer_params"])
    assert w["/layers/dense/vars/0"].shape == (256, 100) and w["/layers/conv2d/vars/0"].shape == (3, 3, 1, 16)
    assert len([l for l in cfg["config"]["layers"] if l["class_name"] == "BatchNormalization"]) == 23


@pytest.mark.gpu
def test_gpu_engine_passes_the_reference_conversion_gate(synth, blob):
    from birdnet_stm32.evaluation.gpu_runner import GpuRunner

    ref = np.load(os.path.join(GOLDEN, "keras_float_reference.npz"))
    cfg, T, pcm, peak = batch(synth)
    runner = GpuRunner(blob, cfg)
    try:
        scores = runner.predict_pcm16(pcm, peak)          # full CUDA path: STFT frontend + int8 graph
    finally:
        runner.close()
    cos, cos_mean, mae = gate(scores, ref["scores"])
    print(f"CUDA engine vs Keras float: cosine mean {cos_mean:.4f} (min {min(cos[:N_GATED]):.4f}), MAE {mae:.5f}")
    assert cos_mean >= 0.95 and min(cos[:N_GATED]) >= 0.95 and mae <= 0.01


def test_int8_layers_track_the_float_checkpoint_layer_by_layer(synth, oracle_model, graph):
    """Every post-activation tensor of the int8 graph, dequantised, against the float model's activation at the same
    layer: per-chunk profiles along channels, rows and columns (mean over the other two axes).  Correlation >= 0.97 and
    relative L2 error <= 0.25 per layer (quantisation noise accumulates with depth; a permuted channel axis, a
    transposed map or a wrong padding side gives correlations near zero)."""
    from oracle import bn_oracle

    ref = np.load(os.path.join(GOLDEN, "keras_float_reference.npz"))
    names = [str(n) for n in ref["tap_names"]]
    cfg, T, pcm, peak = batch(synth)
    spec = bn_oracle.frontend_hybrid(pcm[:N_GATED], peak[:N_GATED], cfg["fft_length"], T // cfg["spec_width"], cfg["spec_width"])
    # tflite tensors in the order of the float taps: frontend output = input of the stem conv; then every op with a fused
    # activation from the stem on (CONV_2D / DEPTHWISE_CONV_2D with RELU6, residual ADD with RELU6); MEAN; FULLY_CONNECTED
    stem = next(op for op in graph.ops if op.kind == "CONV_2D" and op.options.get("stride_w") == 2)
    tids = [stem.inputs[0]]
    tids += [op.outputs[0] for op in graph.ops if op.index >= stem.index and op.kind in ("CONV_2D", "DEPTHWISE_CONV_2D", "ADD")
             and op.options.get("act", "NONE") != "NONE"]
    tids += [next(op.outputs[0] for op in graph.ops if op.kind == "MEAN"), next(op.outputs[0] for op in graph.ops if op.kind == "FULLY_CONNECTED")]
    assert len(tids) == len(names) == 26
    worst_corr, worst_rel = 1.0, 0.0
    for i, (name, tid) in enumerate(zip(names, tids)):
        t = graph.tensors[tid]
        _, tap = oracle_model.run(spec, tap_id=tid)
        q = tap.reshape((spec.shape[0],) + tuple(t.shape[1:])).astype(np.float64)
        x = (q - float(t.zero_point[0])) * float(t.scale[0])
        if x.ndim == 4 and name == "frontend":
            pass                                       # [B, 64 mel, 256, 1] in both graphs
        profiles = [("c", x.mean(axis=(1, 2)) if x.ndim == 4 else x)]
        if x.ndim == 4:
            profiles += [("h", x.mean(axis=(2, 3))), ("w", x.mean(axis=(1, 3)))]
        for ax, got in profiles:
            want = ref[f"tap{i}_{ax}"].astype(np.float64)
            assert got.shape == want.shape, (name, ax, got.shape, want.shape)
            if want.shape[1] < 3 or np.std(want) < 1e-9:
                continue
            corr = float(np.corrcoef(got.reshape(-1), want.reshape(-1))[0, 1])
            rel = float(np.linalg.norm(got - want) / (np.linalg.norm(want) + 1e-12))
            worst_corr, worst_rel = min(worst_corr, corr), max(worst_rel, rel)
            assert corr >= 0.97 and rel <= 0.25, (name, ax, corr, rel)
    print(f"26 layers vs float checkpoint: worst profile correlation {worst_corr:.4f}, worst relative L2 error {worst_rel:.3f}")


def test_shipped_tflite_weights_are_reproduced_from_the_float_checkpoint(graph):
    """Bit-level pin of the weight / bias path against the REAL TensorFlow Lite converter: quantising the float
    checkpoint's BatchNorm-folded weights per output channel (scale = max|w| / 127, round to nearest) must give exactly
    the int8 weights the shipped `.tflite` holds, its scales to float32 accuracy, and its int32 biases
    (bias / (input scale x weight scale)) to within one unit.  This fixes the reading of both files: OHWI / 1HWC layouts,
    quantised dimension, BatchNorm folding, which convolution is which."""
    import hashlib

    ref = np.load(os.path.join(GOLDEN, "keras_weight_reference.npz"))
    names = [str(n) for n in ref["names"]]
    ops = [op for op in graph.ops if op.kind in ("CONV_2D", "DEPTHWISE_CONV_2D", "FULLY_CONNECTED")
           and not (op.kind == "DEPTHWISE_CONV_2D" and graph.tensors[op.inputs[1]].shape[1:3] == (1, 1))]      # skip the 1x1 PWL depthwise ops
    assert len(ops) == len(names) == 25
    n_weights = n_bias = bias_off = 0
    for name, op in zip(names, ops):
        wt = graph.tensors[op.inputs[1]]
        live = ref[f"{name}/live"]
        axis = int(ref[f"{name}/axis"])
        assert wt.quantized_dimension == axis and wt.shape[axis] == live.size, name
        shape = [1] * len(wt.shape)
        shape[axis] = -1
        q = np.where(np.broadcast_to(live.reshape(shape), wt.data.shape), wt.data, 0).astype(np.int8)
        dig = np.frombuffer(hashlib.sha256(np.ascontiguousarray(q).tobytes()).digest()[:8], dtype=np.uint64)[0]
        assert dig == ref[f"{name}/digest"], f"{name}: int8 weights differ from the quantised float checkpoint"
        n_weights += int(np.broadcast_to(live.reshape(shape), wt.data.shape).sum())
        s_ref = ref[f"{name}/scale"][live]
        s = wt.scale.astype(np.float64)[live]
        assert np.all(np.abs(s - s_ref) <= 1e-6 * s_ref), name
        assert np.all(wt.scale.astype(np.float64)[~live] <= 2e-8), name      # dead channels (|w| ~ 1e-40): the converter's floor scales
        if f"{name}/bias" in ref.files and len(op.inputs) > 2 and op.inputs[2] >= 0:
            bt, it = graph.tensors[op.inputs[2]], graph.tensors[op.inputs[0]]
            want = np.round(ref[f"{name}/bias"] / (float(it.scale[0]) * wt.scale.astype(np.float64)))
            ok = np.abs(want - bt.data.astype(np.float64)) <= 1
            n_bias += int(live.sum())
            bias_off += int((~ok[live]).sum())
    print(f"{n_weights} int8 weights of 25 layers identical to the quantised float checkpoint; {n_bias - bias_off}/{n_bias} biases within 1")
    assert n_weights > 180000 and bias_off <= 2
