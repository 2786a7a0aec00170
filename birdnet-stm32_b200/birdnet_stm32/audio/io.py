"""Audio ingest for the B200 path: PCM16 WAV -> mono int16 + file peak, and chunk geometry.

Semantics follow the reference `birdnet_stm32/audio/io.py`:
  * `load_audio_window` (`:63-130`): read at most `max_duration` seconds from the start, mono,
    peak-normalise the whole window (`y / max|y|` when the peak is > 0).  The engine keeps the
    samples as int16 and carries the peak as a float32 scalar per chunk; the division happens in
    float32 on the device exactly where the reference does it (`pcm/32768` then `/ peak`).
  * `split_audio_into_chunks` (`:133-174`) and `estimate_num_chunks` (`:33-60`): chunk_size =
    int(sr * cd); a short window gives one right-zero-padded chunk; otherwise starts every
    step = int(sr * (cd - overlap)) with overlap clamped to [0, cd - 0.1], plus an end-anchored
    tail chunk when samples remain.

Decoding uses the stdlib `wave` module (soundfile/libsndfile are not part of this image).  Files
that would need the reference's host-side resampling or channel averaging (sample rate different from
the model's, multi-channel, non-16-bit) are reported with `UnsupportedAudio`; `evaluate()` skips them
the way the reference skips unreadable files (`metrics.py:125-126`) and counts them.
"""

from __future__ import annotations

import wave

import numpy as np


class UnsupportedAudio(Exception):
    """The file cannot be fed to the PCM16 device path without host-side resampling / mixing."""


def chunk_starts(num_samples: int, sample_rate: int, chunk_duration: float, chunk_overlap: float = 0.0) -> np.ndarray:
    """Start offsets of the chunks `split_audio_into_chunks` emits (empty for an empty window)."""
    size = int(sample_rate * chunk_duration)
    if num_samples <= 0 or size <= 0:
        return np.zeros((0,), dtype=np.int64)
    if num_samples <= size:
        return np.zeros((1,), dtype=np.int64)
    overlap = max(0.0, min(chunk_overlap, chunk_duration - 0.1))
    step = max(1, int(sample_rate * (chunk_duration - overlap)))
    starts = np.arange(0, num_samples - size + 1, step, dtype=np.int64)
    if starts.size == 0 or starts[-1] + size < num_samples:
        starts = np.append(starts, num_samples - size)
    return starts


def estimate_num_chunks(num_samples: int, sample_rate: int, chunk_duration: float, chunk_overlap: float = 0.0) -> int:
    return int(chunk_starts(num_samples, sample_rate, chunk_duration, chunk_overlap).size)


def split_audio_into_chunks(audio: np.ndarray, sample_rate: int = 24000, chunk_duration: float = 3.0,
                            chunk_overlap: float = 0.0) -> np.ndarray:
    """[N] -> [n_chunks, chunk_size], dtype preserved for int16, float32 otherwise."""
    size = int(sample_rate * chunk_duration)
    y = np.asarray(audio).reshape(-1)
    dtype = np.int16 if y.dtype == np.int16 else np.float32
    if y.size == 0 or size <= 0:
        return np.empty((0, max(size, 0)), dtype=dtype)
    y = y.astype(dtype, copy=False)
    starts = chunk_starts(y.size, sample_rate, chunk_duration, chunk_overlap)
    out = np.zeros((starts.size, size), dtype=dtype)
    if y.size <= size:
        out[0, : y.size] = y
        return out
    for i, s in enumerate(starts):
        out[i] = y[s : s + size]
    return out


def read_wav_pcm16(path: str, max_frames: int | None = None) -> tuple[np.ndarray, int]:
    """Mono 16-bit PCM WAV -> (int16 [frames], sample_rate).  Anything else raises UnsupportedAudio."""
    try:
        with wave.open(path, "rb") as wf:
            if wf.getsampwidth() != 2 or wf.getcomptype() != "NONE":
                raise UnsupportedAudio(f"{path}: only 16-bit PCM WAV is decoded on this path")
            if wf.getnchannels() != 1:
                raise UnsupportedAudio(f"{path}: {wf.getnchannels()} channels (channel averaging needs the float path)")
            n = wf.getnframes() if max_frames is None else min(wf.getnframes(), max_frames)
            raw = wf.readframes(n)
            return np.frombuffer(raw, dtype="<i2").astype(np.int16, copy=False), wf.getframerate()
    except wave.Error as exc:
        raise UnsupportedAudio(f"{path}: {exc}") from exc


def load_pcm16_window(path: str, sample_rate: int, max_duration: float | None = 60) -> tuple[np.ndarray, np.float32]:
    """(int16 mono window, file peak) -- the device-path twin of `load_audio_window`.

    peak = max|pcm / 32768| in float32 (0 for silence, meaning "do not normalise").
    """
    limit = None if not max_duration or max_duration <= 0 else int(float(max_duration) * sample_rate)
    pcm, sr0 = read_wav_pcm16(path, limit)
    if sr0 != sample_rate:
        raise UnsupportedAudio(f"{path}: sample rate {sr0} != model rate {sample_rate} (host resampling is a next-step row)")
    peak = np.float32(np.abs(pcm.astype(np.float32) / np.float32(32768.0)).max()) if pcm.size else np.float32(0)
    return pcm, peak


def load_audio_window(path: str, sample_rate: int = 24000, max_duration: float | None = 30,
                      chunk_duration: float = 3.0, random_offset: bool = False) -> np.ndarray:
    """Reference-compatible float view: mono float32 in [-1, 1], peak normalised; empty on error."""
    try:
        pcm, peak = load_pcm16_window(path, sample_rate, max_duration)
    except Exception:
        return np.empty((0,), dtype=np.float32)
    y = pcm.astype(np.float32) / np.float32(32768.0)
    if peak > 0:
        y = y / peak
    return y.astype(np.float32, copy=False)


def load_audio_file(path: str, sample_rate: int = 24000, max_duration: int = 30, chunk_duration: float = 3.0,
                    chunk_overlap: float = 0.0, random_offset: bool = False):
    """Float32 chunks `[n, chunk_size]` like the reference (`io.py:177-213`); `[]` on error."""
    audio = load_audio_window(path, sample_rate, max_duration, chunk_duration, random_offset)
    if audio.size == 0:
        return []
    return split_audio_into_chunks(audio, sample_rate, chunk_duration, chunk_overlap)


def load_pcm16_chunks(path: str, sample_rate: int, chunk_duration: float, chunk_overlap: float = 0.0,
                      max_duration: float | None = 60) -> tuple[np.ndarray, np.float32]:
    """PCM16 chunks `[n, chunk_size]` + the file peak for the device path (`[0, size]` if the file is empty)."""
    pcm, peak = load_pcm16_window(path, sample_rate, max_duration)
    return split_audio_into_chunks(pcm, sample_rate, chunk_duration, chunk_overlap), peak


def save_wav(audio: np.ndarray, path: str, sample_rate: int = 24000) -> None:
    """Write mono 16-bit PCM (float input in [-1, 1] is scaled by 32767 and rounded)."""
    a = np.asarray(audio)
    if a.dtype != np.int16:
        a = np.round(32767.0 * np.clip(a.astype(np.float64), -1.0, 1.0)).astype(np.int16)
    with wave.open(path, "wb") as wf:
        wf.setnchannels(1)
        wf.setsampwidth(2)
        wf.setframerate(int(sample_rate))
        wf.writeframes(a.astype("<i2").tobytes())
