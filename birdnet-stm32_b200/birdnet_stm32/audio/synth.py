"""Synthetic PCM16 chunk generators shared by the tests and `bench.py`.

The recipes follow the reference's synthesised fixtures (sine / silence /
seeded noise, `tests/conftest.py:49-81`, `tests/fixtures/generate_fixtures.py:17-32`)
extended with chirp mixtures.  Nothing here
reads audio files.
"""

from __future__ import annotations

import numpy as np


def synth_wave(n: int, T: int, sample_rate: int, seed: int = 1234, edge_cases: bool = False) -> np.ndarray:
    """float64 [n, T] mixtures of 1-3 linear chirps + white noise, in [-1, 1]-ish."""
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
        x[n - 1] = 0.0                                             # silence
        x[n - 2] = np.where((np.arange(T) // 50) % 2 == 0, 1.0, -1.0)   # full-scale square
        x[n - 3] = 0.0
        x[n - 3, T // 2] = 1.0                                     # single impulse
        x[n - 4] = 0.5 * np.sin(2 * np.pi * 1000.0 * t)            # 1 kHz sine (reference conftest)
    return x


def synth_pcm16(n: int, T: int, sample_rate: int, seed: int = 1234, edge_cases: bool = False) -> np.ndarray:
    """int16 [n, T]: round(32767 * clip(x, -1, 1))."""
    x = synth_wave(n, T, sample_rate, seed, edge_cases)
    return np.round(32767.0 * np.clip(x, -1.0, 1.0)).astype(np.int16)


def file_peaks(pcm: np.ndarray) -> np.ndarray:
    """Per-chunk peak of pcm/32768 as float32 (each chunk treated as its own file;
    reference peak normalisation `audio/io.py:124-126`).  0 for silent chunks."""
    return np.abs(pcm.astype(np.float32) / np.float32(32768.0)).max(axis=1).astype(np.float32)
