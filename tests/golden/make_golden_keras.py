"""Generates tests/golden/keras_float_reference.npz.  Run ONCE in the build container (where /root/reference is mounted).

The reference ships the float Keras checkpoint its `.tflite` was converted from
(`checkpoints/birdnet_stm32n6_100.keras`).  TensorFlow / Keras / h5py are not installed here, so the archive is read with
`oracle/h5min.py` and evaluated with `oracle/keras_float_model.py` (torch CPU convolutions, layer sequence from the
archive's own config.json).  Stored: the float model's sigmoid scores for the seeded 16-chunk batch the other parity
tests use -- the reference-side operand of the reference's own conversion gate (`conversion/validate.py`, cosine >= 0.95).
"""

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "birdnet-stm32_b200"))

from birdnet_stm32.audio import synth
from oracle import bn_oracle
from oracle.keras_float_model import KerasFloatModel


def main():
    cfg = json.load(open(os.path.join(ROOT, "tests", "fixtures", "birdnet_stm32n6_100_model_config.json")))
    T = int(cfg["sample_rate"] * cfg["chunk_duration"])
    pcm = synth.synth_pcm16(16, T, cfg["sample_rate"], seed=1234, edge_cases=True)
    peak = synth.file_peaks(pcm)
    spec = bn_oracle.frontend_hybrid(pcm, peak, cfg["fft_length"], T // cfg["spec_width"], cfg["spec_width"])
    km = KerasFloatModel("/root/reference/checkpoints/birdnet_stm32n6_100.keras")
    taps: dict = {}
    scores = km.predict(spec, taps)
    n_params = int(sum(v.size for k, v in km.w.items() if k.startswith("/layers/")))
    out = dict(scores=scores, n_layer_params=np.int64(n_params), spec_sum=spec.astype(np.float64).sum(axis=(1, 2, 3)))
    # layer-by-layer pin: per-chunk activation profiles of the float model along each axis (mean over the other two) for
    # the first 12 chunks -- small, but a channel permutation, a transposed map or a shifted padding changes them
    names = list(taps)
    out["tap_names"] = np.array(names)
    for i, k in enumerate(names):
        a = taps[k][:12].astype(np.float64)
        if a.ndim == 4:
            out[f"tap{i}_c"] = a.mean(axis=(1, 2)).astype(np.float32)
            out[f"tap{i}_h"] = a.mean(axis=(2, 3)).astype(np.float32)
            out[f"tap{i}_w"] = a.mean(axis=(1, 3)).astype(np.float32)
        else:
            out[f"tap{i}_c"] = a.astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "keras_float_reference.npz"), **out)
    print("taps", len(names), names[:4], "...")
    print("scores", scores.shape, "layer parameters", n_params, "top-1", scores.argmax(1).tolist())


if __name__ == "__main__":
    main()
