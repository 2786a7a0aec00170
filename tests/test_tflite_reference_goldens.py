"""Consumes `tests/golden/tflite_reference.npz` -- outputs of the REAL reference (`get_spectrogram_from_audio` +
`TFLiteRunner.predict` + per-tensor dumps) captured by `scripts/capture_reference_goldens.py` on a machine that has
TensorFlow 2.19 / librosa 0.11.  This image has neither, so the file cannot be produced here: the tests below are then
reported as XFAIL (strict: they cannot silently pass), and "parity unpinned" stays in DESIGN.md.  With the file present they
are ordinary tests and are the pin of the oracle -- and through it of the CUDA engine -- at the TFLite / librosa boundaries.
"""

import os

import numpy as np
import pytest

from conftest import GOLDEN

PATH = os.path.join(GOLDEN, "tflite_reference.npz")
absent = pytest.mark.xfail(condition=not os.path.exists(PATH), strict=True, raises=FileNotFoundError,
                           reason="tests/golden/tflite_reference.npz has not been captured (needs TensorFlow + librosa; "
                                  "run scripts/capture_reference_goldens.py once on such a machine)")
CASES = (("sr22050", 22050, 66150), ("sr24000", 24000, 72000))


def _load():
    if not os.path.exists(PATH):
        raise FileNotFoundError(PATH)
    return np.load(PATH)


def test_capture_script_generator_matches_synth(synth):
    """The standalone generator of the capture script is the one the parity tests use."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("cap", os.path.join(os.path.dirname(GOLDEN), "..", "scripts", "capture_reference_goldens.py"))
    cap = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cap)
    a = cap.synth_pcm16(6, 4000, 22050, seed=9, edge_cases=True)
    np.testing.assert_array_equal(a, synth.synth_pcm16(6, 4000, 22050, seed=9, edge_cases=True))
    y = cap.reference_float_chunk(a[0])
    assert y.dtype == np.float32 and np.max(np.abs(y)) == np.float32(1.0)


@absent
@pytest.mark.parametrize("tag,sr,T", CASES)
def test_oracle_frontend_equals_the_reference_spectrogram(tag, sr, T, synth):
    from oracle import bn_oracle

    g = _load()
    pcm = g[f"{tag}_pcm"]
    np.testing.assert_array_equal(pcm, synth.synth_pcm16(len(pcm), T, sr, seed=1234, edge_cases=True))
    ours = bn_oracle.frontend_hybrid(pcm, synth.file_peaks(pcm), 512, T // 256, 256)
    ref = g[f"{tag}_spec"]
    assert np.abs(ours - ref).max() <= 2e-6, "oracle STFT / normalise differs from librosa's"
    q = lambda s: np.clip(np.round(s / np.float32(0.003921568859368563)) - 128, -128, 127)
    assert (q(ours) == q(ref)).mean() >= 0.9999


@absent
@pytest.mark.parametrize("tag,sr,T", CASES)
def test_oracle_graph_is_bit_exact_against_the_tflite_interpreter(tag, sr, T, blob):
    """Scores and every dumped int8 tensor; also reports WHICH rounding / MEAN variant the real kernels use."""
    from oracle import bn_oracle, tflite_quant
    from conftest import TFLITE

    g = _load()
    spec, want = g[f"{tag}_spec"], g[f"{tag}_scores"]
    b = tflite_quant.patch_blob(blob, tflite_quant.derive(TFLITE))
    hits = [(r, mv) for r in (0, 1) for mv in (0, 1, 2, 3)
            if np.array_equal(bn_oracle.OracleModel(b, rounding=r, mean_variant=mv).predict(spec), want)]
    assert (0, 0) in hits, f"default (double rounding, auto MEAN) does not reproduce TFLite; matching variants: {hits}"
    model = bn_oracle.OracleModel(b)
    for key in g.files:
        if key.startswith(f"{tag}_tensor_"):
            tid = int(key.rsplit("_", 1)[1])
            try:
                _, ours = model.run(spec, tap_id=tid)
            except KeyError:
                continue                                   # a tensor the lowering folds away (shape arithmetic)
            assert np.array_equal(ours.reshape(-1), g[key].reshape(-1)), f"tensor {tid} differs"


@absent
@pytest.mark.gpu
@pytest.mark.parametrize("tag,sr,T", CASES)
def test_engine_against_the_tflite_interpreter(tag, sr, T, blob, cfg, synth):
    from birdnet_stm32.conversion.export_blob import export_blob
    from birdnet_stm32.evaluation.gpu_runner import GpuRunner
    from conftest import TFLITE

    g = _load()
    c = dict(cfg, sample_rate=sr)
    r = GpuRunner(export_blob(TFLITE, c), c)
    try:
        np.testing.assert_array_equal(r.predict(g[f"{tag}_spec"]), g[f"{tag}_scores"])          # int8 body: bit-exact
        pcm = g[f"{tag}_pcm"]
        got = r.predict_pcm16(pcm, synth.file_peaks(pcm))
        assert (got.argmax(1) == g[f"{tag}_scores"].argmax(1)).mean() >= 0.9
    finally:
        r.close()
