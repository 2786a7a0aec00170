#!/usr/bin/env python
"""One-shot pin against the REAL reference arithmetic.

Run this ONCE on any machine that has the reference's pinned dependencies (requirements.txt: tensorflow==2.19.0,
librosa==0.11.0, numpy==1.26.4, soundfile) and a checkout of birdnet-team/birdnet-stm32:

    python scripts/capture_reference_goldens.py --reference /path/to/birdnet-stm32 [--out tests/golden/tflite_reference.npz]

It feeds the seeded synthetic batch of this repo's parity tests through the reference's own code --
`birdnet_stm32.audio.spectrogram.get_spectrogram_from_audio` exactly as `make_chunks_for_file` calls it for the hybrid
frontend (`evaluation/metrics.py:55-61`), then `birdnet_stm32.models.runners.TFLiteRunner.predict`
(`models/runners.py:82-95`) on the shipped `checkpoints/birdnet_stm32n6_100.tflite` -- and additionally dumps every int8
activation tensor with `tf.lite.Interpreter(experimental_preserve_all_tensors=True)` (same construction as
`conversion/validate.py:64-77`).  The resulting `.npz` is consumed by `tests/test_tflite_reference_goldens.py`: once it is
committed, the oracle (and through it the CUDA engine) is pinned bit for bit at the TFLite and librosa boundaries and the
`BN_OPT_ROUNDING` / `BN_OPT_MEAN_VARIANT` defaults stop being assumptions.  This image has neither TensorFlow nor librosa
(no network), which is why the file is not produced here.

The script needs nothing from this repository except numpy: the signal generator below is a copy of
`birdnet_stm32/audio/synth.py` (kept identical by `test_capture_script_generator_matches_synth`).
"""

from __future__ import annotations

import argparse
import os
import sys

import numpy as np

CASES = (("sr22050", 22050, 66150), ("sr24000", 24000, 72000))      # shipped config / BASELINE "3 s at 24 kHz"
N_CHUNKS, SEED, N_FFT, SPEC_WIDTH = 12, 1234, 512, 256


def synth_wave(n: int, T: int, sample_rate: int, seed: int = 1234, edge_cases: bool = False) -> np.ndarray:
    rng = np.random.default_rng(seed)
    t = np.arange(T, dtype=np.float64) / sample_rate
    dur = T / sample_rate
    x = np.zeros((n, T), dtype=np.float64)
    for b in range(n):
        for _ in range(int(rng.integers(1, 4))):
            f0, f1 = rng.uniform(300.0, min(10000.0, 0.45 * sample_rate), 2)
            amp = rng.uniform(0.1, 0.8)
            x[b] += amp * np.sin(2 * np.pi * (f0 * t + (f1 - f0) * t * t / (2 * dur)))
        x[b] += rng.uniform(0.05, 0.3) * rng.standard_normal(T)
    if edge_cases and n >= 5:
        x[n - 1] = 0.0
        x[n - 2] = np.where((np.arange(T) // 50) % 2 == 0, 1.0, -1.0)
        x[n - 3] = 0.0
        x[n - 3, T // 2] = 1.0
        x[n - 4] = 0.5 * np.sin(2 * np.pi * 1000.0 * t)
    return x


def synth_pcm16(n: int, T: int, sample_rate: int, seed: int = 1234, edge_cases: bool = False) -> np.ndarray:
    x = synth_wave(n, T, sample_rate, seed, edge_cases)
    return np.round(32767.0 * np.clip(x, -1.0, 1.0)).astype(np.int16)


def reference_float_chunk(pcm: np.ndarray) -> np.ndarray:
    """What `load_audio_window` hands on for a one-chunk PCM16 file: soundfile float32 = s / 32768, then / max|y| (io.py:114-126)."""
    y = pcm.astype(np.float32) / np.float32(32768.0)
    peak = np.max(np.abs(y))
    return (y / peak).astype(np.float32) if peak > 0 else y


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", required=True, help="checkout of birdnet-team/birdnet-stm32")
    ap.add_argument("--out", default=os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "tflite_reference.npz"))
    args = ap.parse_args()
    sys.path.insert(0, args.reference)
    import librosa
    import tensorflow as tf

    from birdnet_stm32.audio.spectrogram import get_spectrogram_from_audio
    from birdnet_stm32.models.runners import TFLiteRunner

    model = os.path.join(args.reference, "checkpoints", "birdnet_stm32n6_100.tflite")
    out: dict = {"versions": np.array([f"tensorflow {tf.__version__}", f"librosa {librosa.__version__}", f"numpy {np.__version__}"])}
    for tag, sr, T in CASES:
        pcm = synth_pcm16(N_CHUNKS, T, sr, seed=SEED, edge_cases=True)
        specs = []
        for b in range(N_CHUNKS):
            S = get_spectrogram_from_audio(reference_float_chunk(pcm[b]), sample_rate=sr, n_fft=N_FFT, mel_bins=-1, spec_width=SPEC_WIDTH)
            specs.append(S[: N_FFT // 2 + 1, :SPEC_WIDTH, None].astype(np.float32))
        spec = np.stack(specs)
        scores = TFLiteRunner(model).predict(spec)
        out[f"{tag}_pcm"] = pcm
        out[f"{tag}_spec"] = spec
        out[f"{tag}_scores"] = scores
        # every tensor of the graph (int8 activations are what the tests compare)
        it = tf.lite.Interpreter(model_path=model, experimental_delegates=[], experimental_preserve_all_tensors=True)
        inp = it.get_input_details()[0]["index"]
        it.resize_tensor_input(inp, spec.shape)
        it.allocate_tensors()
        it.set_tensor(inp, spec)
        it.invoke()
        np.testing.assert_array_equal(it.get_tensor(it.get_output_details()[0]["index"]), scores)
        for d in it.get_tensor_details():
            if d["dtype"] == np.int8 and len(d["shape"]) >= 2 and d["shape"][0] == N_CHUNKS:
                try:
                    out[f"{tag}_tensor_{d['index']}"] = it.get_tensor(d["index"])
                except ValueError:
                    pass
    np.savez_compressed(args.out, **out)
    print(f"wrote {args.out}: {len(out)} arrays; commit it and run `python -m pytest tests/test_tflite_reference_goldens.py`")


if __name__ == "__main__":
    main()
