"""Runner loading: the plugin boundary of the reference (`models/runners.py:98-114`).

`load_model_runner(path)` keeps its signature and its dispatch on the file extension.  `.tflite`
(and `.b200blob`) models get the B200 `GpuRunner`; `.keras` needs TensorFlow, which this package
never imports, so it raises with a clear message instead of silently doing something else.
"""

from __future__ import annotations

from birdnet_stm32.evaluation.gpu_runner import GpuRunner


def load_model_runner(model_path: str, model_config: dict | None = None, device: int = 0, backend: str = "gpu"):
    """Return a runner exposing `predict(x_batch) -> float32 [B, C]`.

    Args:
        model_path: `.tflite` or `.b200blob` path.
        model_config: optional `_model_config.json` dict (looked up next to the model otherwise).
        device: CUDA device ordinal.
        backend: only "gpu" exists in this package.
    """
    if backend != "gpu":
        raise ValueError(f"Unknown backend '{backend}': this package only ships the B200 engine ('gpu').")
    lower = model_path.lower()
    if lower.endswith(".tflite") or lower.endswith(".b200blob"):
        return GpuRunner(model_path, model_config=model_config, device=device)
    raise RuntimeError(
        f"Cannot load '{model_path}': Keras models need TensorFlow; convert to .tflite with the reference "
        "`birdnet_stm32 convert` and pass that file."
    )
