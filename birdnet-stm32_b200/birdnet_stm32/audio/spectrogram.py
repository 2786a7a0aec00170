"""Spectrogram entry points of the B200 path.

The reference computes the hybrid model input on the host with librosa
(`audio/spectrogram.py:24-149`: STFT n_fft=512, hop=len//spec_width, Hann, centred, |.|, keep
`spec_width` frames, min-max normalise).  Here that computation is the CUDA kernel `k_stft_mag`
(csrc/bn_frontend.cu) reached through `GpuRunner.frontend`; this module only keeps the function name
callers know and the normalisation contract.
"""

from __future__ import annotations

import numpy as np


def normalize(S: np.ndarray) -> np.ndarray:
    """`(S - min) / (max - min + 1e-10)` -- contract of reference `spectrogram.py:12-21` (host helper for
    small arrays in tests and tools; the engine normalises on the device)."""
    return (S - S.min()) / (S.max() - S.min() + 1e-10)


def get_spectrogram_from_pcm16(runner, pcm: np.ndarray, peak=None) -> np.ndarray:
    """Linear-magnitude model input `[B, n_fft/2+1, spec_width]` in [0, 1] for PCM16 chunks `[B, T]`.

    Equivalent of `get_spectrogram_from_audio(chunk, mel_bins=-1)` per chunk; runs on the GPU.
    """
    pcm = np.asarray(pcm)
    if pcm.ndim == 1:
        pcm = pcm[None, :]
    return runner.frontend(pcm, peak)[..., 0]
