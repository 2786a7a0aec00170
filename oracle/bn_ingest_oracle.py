"""CPU oracle for the device ingest (TEST INFRASTRUCTURE ONLY -- never imported by the package).

Restates the host side of the reference's `load_audio_window` / `fast_resample` /
`split_audio_into_chunks` (`birdnet_stm32/audio/io.py:14-30,112-128,133-174`) on samples that were already read
from the container:

  * decode: what `soundfile.SoundFile.read(dtype="float32", always_2d=True)` returns for the WAV sample formats
    (libsndfile 1.2 `pcm.c`, `float32.c`; third-party, not vendored in /root/reference): int16 / 2^15, int24 / 2^23,
    (float32)int32 / 2^31, float32 as is, (uint8 - 128) / 2^7;
  * `y.mean(axis=1)` with numpy itself (float32 accumulation, division by the channel count);
  * `scipy.signal.resample_poly(y, up, down)` -- scipy IS in this image, so the resampler is the reference's own
    third-party code, not a restatement;
  * `peak = max|y|`, `y / peak` (numpy, float32);
  * chunk geometry of `split_audio_into_chunks`.

Pinned by `tests/golden/ingest_reference.npz`: outputs of the REAL `load_audio_window` and
`split_audio_into_chunks` of /root/reference, run in the build container on WAV files through a stub `soundfile`
module (the stub only parses the container and decodes samples as listed above; mean, resampling, peak
normalisation and chunking are the reference's code).  Generator: `tests/golden/make_golden.py`.
"""

from __future__ import annotations

from math import gcd

import numpy as np
from scipy.signal import resample_poly


def decode(raw: np.ndarray, kind: str, channels: int) -> np.ndarray:
    """Interleaved raw samples -> float32 [frames, channels] (libsndfile float32 read, normalised)."""
    a = np.asarray(raw).reshape(-1)
    if kind == "s16":
        x = a.astype(np.int16).astype(np.float32) * np.float32(1.0 / 32768.0)
    elif kind == "s32":
        x = a.astype(np.int32).astype(np.float32) * np.float32(1.0 / 2147483648.0)
    elif kind == "f32":
        x = a.astype(np.float32)
    elif kind == "u8":
        x = (a.astype(np.int32) - 128).astype(np.float32) * np.float32(1.0 / 128.0)
    elif kind == "s24":
        b = a.astype(np.uint8).reshape(-1, 3).astype(np.int32)
        v = (b[:, 0] << 8) | (b[:, 1] << 16) | (b[:, 2] << 24)
        v = v.astype(np.int32) >> 8
        x = v.astype(np.float32) * np.float32(1.0 / 8388608.0)
    else:
        raise ValueError(kind)
    return x.reshape(-1, channels)


def fast_resample(y: np.ndarray, sr_in: int, sr_out: int) -> np.ndarray:
    """`audio/io.py:14-30`."""
    if sr_in == sr_out:
        return y.astype(np.float32, copy=False)
    g = gcd(sr_in, sr_out)
    return resample_poly(y, sr_out // g, sr_in // g).astype(np.float32, copy=False)


def load_window(raw: np.ndarray, kind: str, channels: int, sr_in: int, sr_out: int, normalize: bool = True):
    """`audio/io.py:112-128` after the read: (mono float32 at sr_out, peak before normalisation)."""
    y = decode(raw, kind, channels)
    if y.size == 0:
        return np.empty((0,), dtype=np.float32), 0.0
    y = y.mean(axis=1).astype(np.float32, copy=False)
    if sr_in != sr_out:
        y = fast_resample(y, sr_in, sr_out)
    peak = float(np.max(np.abs(y))) if y.size else 0.0
    if normalize and peak > 0.0:
        y = y / peak
    return y.astype(np.float32, copy=False), peak


def split_chunks(audio: np.ndarray, sample_rate: int, chunk_duration: float, chunk_overlap: float = 0.0) -> np.ndarray:
    """`audio/io.py:133-174`: `[n_chunks, chunk_size]` float32."""
    size = int(sample_rate * chunk_duration)
    y = np.asarray(audio, dtype=np.float32).reshape(-1)
    if y.size == 0:
        return np.empty((0, size), dtype=np.float32)
    if y.size <= size:
        out = np.zeros((1, size), dtype=np.float32)
        out[0, : y.size] = y
        return out
    overlap = max(0.0, min(chunk_overlap, chunk_duration - 0.1))
    step = max(1, int(sample_rate * (chunk_duration - overlap)))
    starts = list(range(0, y.size - size + 1, step))
    if starts[-1] + size < y.size:
        starts.append(y.size - size)
    return np.stack([y[s : s + size] for s in starts], axis=0)


def resample_filter(up: int, down: int) -> tuple[np.ndarray, int]:
    """The float32 taps `resample_poly` hands to upfirdn (padded, times up) and `n_pre_remove` -- for the tap test."""
    from scipy.signal import firwin

    g = gcd(up, down)
    up //= g
    down //= g
    max_rate = max(up, down)
    half_len = 10 * max_rate
    h = firwin(2 * half_len + 1, 1.0 / max_rate, window=("kaiser", 5.0)).astype(np.float32)
    h *= up
    n_pre_pad = down - half_len % down
    return np.concatenate((np.zeros(n_pre_pad, np.float32), h)), (half_len + n_pre_pad) // down
