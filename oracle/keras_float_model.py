"""Float32 forward pass of the reference's shipped Keras checkpoint (TEST INFRASTRUCTURE ONLY).

`/root/reference/checkpoints/birdnet_stm32n6_100.keras` is the float model the shipped `.tflite` was converted from.  The
reference accepts a conversion when the int8 TFLite outputs and the Keras outputs have cosine similarity >= 0.95 on the
validation batches (`conversion/validate.py:51-105`, gate in `cli/convert.py:187-195`).  This module evaluates the same
gate WITHOUT TensorFlow: the weights are read from the archive with `oracle/h5min.py`, the layer sequence from its
`config.json`, and the layers (Conv2D / DepthwiseConv2D / BatchNormalization / ReLU(max 6) / Add /
GlobalAveragePooling2D / Dense + sigmoid, `models/dscnn.py:28-84,198-262`; the hybrid frontend `models/frontend.py:299-345`
and its PWL magnitude layer `models/magnitude.py:179-192`) are evaluated with torch CPU convolutions.  It is an independent
route to the network output: Keras float weights + BatchNorm statistics instead of the converter's folded int8 weights,
scales and zero points, so it pins the oracle's reading of the `.tflite` (weight layouts, padding, per-channel axes,
residual wiring) at the reference's own acceptance level.

The frontend follows the graph the shipped `.tflite` holds: transpose -> slice -> zero-pad 257 -> 264
channels -> 1x1 mel mixer -> ReLU -> PWL, i.e. the checkpoint predates the per-sample max normalisation of the current
`AudioFrontendLayer.call`.
"""

from __future__ import annotations

import numpy as np

from oracle.h5min import read_keras_weights


def _same_pad(x, k, s):
    """TensorFlow SAME padding for NCHW `x` (extra pixel goes after)."""
    import torch.nn.functional as F

    pads = []
    for size, kk, ss in ((x.shape[3], k[1], s[1]), (x.shape[2], k[0], s[0])):      # F.pad: last dim first
        out = -(-size // ss)
        total = max((out - 1) * ss + kk - size, 0)
        pads += [total // 2, total - total // 2]
    return F.pad(x, pads)


class KerasFloatModel:
    def __init__(self, keras_path: str):
        self.cfg, self.w = read_keras_weights(keras_path)
        self.layers = self.cfg["config"]["layers"]
        # Keras 3 stores weights under class-based unique names in layer order: conv2d, conv2d_1, ...
        counters: dict[str, int] = {}
        self.h5name = {}
        snake = {"Conv2D": "conv2d", "DepthwiseConv2D": "depthwise_conv2d", "BatchNormalization": "batch_normalization", "Dense": "dense",
                 "AudioFrontendLayer": "audio_frontend_layer"}
        for layer in self.layers:
            base = snake.get(layer["class_name"])
            if base is None:
                continue
            i = counters.get(base, 0)
            counters[base] = i + 1
            self.h5name[layer["config"]["name"]] = base if i == 0 else f"{base}_{i}"

    def var(self, layer_name: str, idx: int, sub: str = "") -> np.ndarray:
        return self.w[f"/layers/{self.h5name[layer_name]}{sub}/vars/{idx}"]

    def frontend(self, spec):
        """[B, 257, W, 1] normalised |STFT| -> [B, 1(C), 64(H = mel), W] NCHW input of the stem."""
        import torch

        name = next(l["config"]["name"] for l in self.layers if l["class_name"] == "AudioFrontendLayer")
        x = torch.from_numpy(np.ascontiguousarray(spec[..., 0], dtype=np.float32))           # [B, bins, T]
        mix = torch.from_numpy(self.var(name, 0, "/mel_mixer")[0, 0])                          # [264, 64]
        y = torch.einsum("bft,fm->bmt", x, mix[: x.shape[1]]).clamp_min(0.0)                  # 1x1 conv over bins, ReLU
        k0 = torch.from_numpy(self.var(name, 0, "/_pwl_k0_dw").reshape(-1))                    # [64]
        out = y * k0[None, :, None]
        for j, suffix in enumerate(("", "_1", "_2")):
            sw = torch.from_numpy(self.var(name, 0, f"/_pwl_shift_dws/depthwise_conv2d{suffix}").reshape(-1))
            sb = torch.from_numpy(self.var(name, 1, f"/_pwl_shift_dws/depthwise_conv2d{suffix}").reshape(-1))
            kk = torch.from_numpy(self.var(name, 0, f"/_pwl_k_dws/depthwise_conv2d{suffix}").reshape(-1))
            out = out + kk[None, :, None] * (y * sw[None, :, None] + sb[None, :, None]).clamp_min(0.0)
        return out[:, None, :, :]                                                              # [B, 1, mel, T]

    def predict(self, spec: np.ndarray, taps: dict | None = None) -> np.ndarray:
        """float32 [B, 257, 256, 1] -> sigmoid scores float32 [B, 100].  If `taps` is a dict it receives the activations
        after the frontend ("frontend"), after every ReLU layer (by layer name), after the pooling ("gap") and the
        pre-sigmoid logits ("logits") as NHWC numpy arrays."""
        import torch
        import torch.nn.functional as F

        def keep(key, t):
            if taps is not None:
                taps[key] = (t.permute(0, 2, 3, 1) if t.dim() == 4 else t).numpy().copy()

        with torch.no_grad():
            x = None
            block_in = None
            for layer in self.layers:
                cls, c = layer["class_name"], layer["config"]
                name = c.get("name")
                if cls == "InputLayer" or cls in ("SpatialDropout2D", "Dropout"):
                    continue
                if cls == "AudioFrontendLayer":
                    x = self.frontend(spec)
                    keep("frontend", x)
                elif cls == "Conv2D":
                    w = torch.from_numpy(self.var(name, 0)).permute(3, 2, 0, 1).contiguous()  # HWIO -> OIHW
                    x = F.conv2d(_same_pad(x, c["kernel_size"], c["strides"]), w, stride=tuple(c["strides"]))
                elif cls == "DepthwiseConv2D":
                    block_in = x
                    w = torch.from_numpy(self.var(name, 0)).permute(2, 3, 0, 1).contiguous()  # HWC1 -> C1HW
                    x = F.conv2d(_same_pad(x, c["kernel_size"], c["strides"]), w, stride=tuple(c["strides"]), groups=x.shape[1])
                elif cls == "BatchNormalization":
                    g, b, m, v = (torch.from_numpy(self.var(name, i)) for i in range(4))
                    scale = g / torch.sqrt(v + float(c["epsilon"]))
                    x = x * scale[None, :, None, None] + (b - m * scale)[None, :, None, None]
                elif cls == "ReLU":
                    x = x.clamp(0.0, float(c["max_value"])) if c.get("max_value") is not None else x.clamp_min(0.0)
                    keep(name, x)
                elif cls == "Add":
                    x = x + block_in
                elif cls == "GlobalAveragePooling2D":
                    x = x.mean(dim=(2, 3))
                    keep("gap", x)
                elif cls == "Dense":
                    x = x @ torch.from_numpy(self.var(name, 0)) + torch.from_numpy(self.var(name, 1))
                    keep("logits", x)
                    if c.get("activation") == "sigmoid":
                        x = torch.sigmoid(x)
                else:
                    raise NotImplementedError(cls)
            return x.numpy().astype(np.float32)


def cosine_similarity(a: np.ndarray, b: np.ndarray, eps: float = 1e-8) -> float:
    """`conversion/validate.py:7-29`."""
    an, bn = np.linalg.norm(a), np.linalg.norm(b)
    if an < eps and bn < eps:
        return 1.0
    if an < eps or bn < eps:
        return 0.0
    return float(np.dot(a, b) / (an * bn))
