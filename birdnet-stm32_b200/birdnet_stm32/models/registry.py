"""Frontend registry (API of the reference `birdnet_stm32/models/registry.py:13-140`), TF-free.

Kept unchanged on purpose: name -> (mode, precomputed, n6_compatible, description).  The B200
capability of a frontend is a separate table (`GPU_FRONTENDS`) so the reference contract is not
altered.
"""

from __future__ import annotations

from dataclasses import dataclass


@dataclass(frozen=True)
class FrontendInfo:
    name: str
    mode: str
    precomputed: bool
    n6_compatible: bool
    description: str = ""


_REGISTRY: dict[str, FrontendInfo] = {}

# frontends whose host-side work has a CUDA kernel in libbn_b200.so: bn_frontend_pcm16 (hybrid) and
# bn_features_pcm16 (librosa / log_mel / mfcc); "raw" only needs x / (max|x| + 1e-6) on the host
GPU_FRONTENDS = frozenset({"hybrid", "librosa", "log_mel", "mfcc"})


def register_frontend(info: FrontendInfo) -> None:
    if info.name in _REGISTRY:
        raise ValueError(f"Frontend '{info.name}' is already registered.")
    _REGISTRY[info.name] = info


def list_frontends() -> list[str]:
    return sorted(_REGISTRY)


def get_frontend_info(name: str) -> FrontendInfo:
    try:
        return _REGISTRY[name]
    except KeyError:
        raise KeyError(f"Frontend '{name}' is not registered. Available: {list_frontends()}") from None


def is_precomputed(name: str) -> bool:
    return get_frontend_info(name).precomputed


def is_n6_compatible(name: str) -> bool:
    return get_frontend_info(name).n6_compatible


def has_gpu_frontend(name: str) -> bool:
    """True when PCM16 -> model input for this frontend runs on the B200 engine."""
    get_frontend_info(name)
    return name in GPU_FRONTENDS


for _name, _mode, _pre, _desc in (
    ("librosa", "precomputed", True, "Host-side mel spectrogram (librosa). Pass-through in model."),
    ("hybrid", "hybrid", False, "Offline STFT + in-model 1x1 Conv2D mel mixer."),
    ("raw", "raw", False, "Raw waveform -> learned Conv2D filterbank (requires T < 65536)."),
    ("mfcc", "precomputed", True, "Host-side MFCC (mel -> DCT -> truncate). Pass-through in model."),
    ("log_mel", "precomputed", True, "Host-side log-mel spectrogram (log1p). Quantization-friendly."),
):
    register_frontend(FrontendInfo(name=_name, mode=_mode, precomputed=_pre, n6_compatible=True, description=_desc))
