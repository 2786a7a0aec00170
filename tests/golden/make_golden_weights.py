"""Generates tests/golden/keras_weight_reference.npz.  Run ONCE in the build container (where /root/reference is mounted).

Bit-level pin of the weight / bias path: the float checkpoint `checkpoints/birdnet_stm32n6_100.keras` is read with
`oracle/h5min.py`, BatchNorm is folded into the preceding convolution (w' = w * gamma / sqrt(var + eps),
b' = beta - mean * gamma / sqrt(var + eps)) and the weights are quantised the way the TensorFlow Lite converter does
(symmetric int8 per output channel, scale = max|w'| / 127, q = round(w' / scale)).  Stored per layer: the expected scales,
the expected bias in real units, a 64-bit digest of the expected int8 weights and the mask of live channels.  The test
compares them with what the shipped `.tflite` (the REAL converter's output) holds.
"""

import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "birdnet-stm32_b200"))

from oracle.keras_float_model import KerasFloatModel


def digest(q: np.ndarray) -> np.uint64:
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(q, dtype=np.int8).tobytes()).digest()[:8], dtype=np.uint64)[0]


def expected_layers(km: KerasFloatModel):
    """[(name, folded float weights in the .tflite layout, output-channel axis, folded bias or None)] in graph order."""
    out = []
    fe = next(l["config"]["name"] for l in km.layers if l["class_name"] == "AudioFrontendLayer")
    mix = km.var(fe, 0, "/mel_mixer").astype(np.float64)                       # [1, 1, 264, 64] HWIO
    out.append(("mel_mixer", np.transpose(mix, (3, 0, 1, 2)), 0, None))
    for i, l in enumerate(km.layers):
        cls, name = l["class_name"], l["config"]["name"]
        if cls in ("Conv2D", "DepthwiseConv2D"):
            bn = km.layers[i + 1]
            g, b, m, v = (km.var(bn["config"]["name"], k).astype(np.float64) for k in range(4))
            sc = g / np.sqrt(v + bn["config"]["epsilon"])
            w = km.var(name, 0).astype(np.float64)
            if cls == "Conv2D":
                wf, axis = np.transpose(w * sc[None, None, None, :], (3, 0, 1, 2)), 0      # HWIO -> OHWI
            else:
                wf, axis = (w[:, :, :, 0] * sc[None, None, :])[None], 3                      # HWC1 -> 1HWC
            out.append((name, wf, axis, b - m * sc))
        elif cls == "Dense":
            out.append((name, km.var(name, 0).astype(np.float64).T, 0, km.var(name, 1).astype(np.float64)))   # [in, out] -> [out, in]
    return out


def main():
    km = KerasFloatModel("/root/reference/checkpoints/birdnet_stm32n6_100.keras")
    res = {}
    names = []
    for name, wf, axis, bias in expected_layers(km):
        axes = tuple(a for a in range(wf.ndim) if a != axis)
        s = np.abs(wf).max(axis=axes) / 127.0
        live = s > 4e-9                      # below max|w| = 5e-7 the converter stores its floor scale 5e-7 / 127 (dead channels)
        shape = [1] * wf.ndim
        shape[axis] = -1
        q = np.round(wf / np.where(live, s, 1.0).reshape(shape))
        q = np.where(np.broadcast_to(live.reshape(shape), q.shape), q, 0).astype(np.int8)
        names.append(name)
        res[f"{name}/scale"] = s
        res[f"{name}/live"] = live
        res[f"{name}/digest"] = digest(q)
        res[f"{name}/axis"] = np.int64(axis)
        if bias is not None:
            res[f"{name}/bias"] = bias
    res["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "keras_weight_reference.npz"), **res)
    print(len(names), "layers;", os.path.getsize(os.path.join(HERE, "keras_weight_reference.npz")), "bytes")


if __name__ == "__main__":
    main()
