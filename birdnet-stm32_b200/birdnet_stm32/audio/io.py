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

Container parsing is done here (RIFF/WAVE: PCM 8/16/24/32-bit and IEEE float32, any channel count) and, for FLAC, by
the native decoder of the reader (`csrc/bn_flac.h`, RFC 9639; soundfile/libsndfile are not part of this image).  Mono 16-bit files at the model rate take the PCM16 path
above.  Everything else -- other sample rates, several channels, other sample formats -- is what the
reference handles with `y.mean(axis=1)` + `scipy.signal.resample_poly` on the host (`io.py:118-120`); here
the raw interleaved samples go to the device once and `bn_ingest_*` (`audio/ingest.py`, `csrc/bn_ingest.cu`)
decodes, mixes, resamples, peak-normalises and chunks them on the GPU.  Lossy containers (MP3 / OGG / M4A) raise
`UnsupportedAudio`; `evaluate()` skips them the way the reference skips unreadable files
(`metrics.py:125-126`), counts them by reason and, with `strict_files=True`, raises instead.
"""

from __future__ import annotations

import struct
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
    if y.size == starts.size * size and (starts.size == 1 or int(starts[1]) == size):
        return y.reshape(starts.size, size)            # back-to-back chunks of an exact multiple: a view, no copy
    out = np.zeros((starts.size, size), dtype=dtype)
    if y.size <= size:
        out[0, : y.size] = y
        return out
    for i, s in enumerate(starts):
        out[i] = y[s : s + size]
    return out


def is_flac(path: str) -> bool:
    try:
        with open(path, "rb") as fh:
            head = fh.read(4)
    except OSError:
        return False
    return head == b"fLaC" or head[:3] == b"ID3"


def read_flac_frames(path: str, max_seconds: float | None = None) -> tuple[np.ndarray, str, int, int]:
    """FLAC -> (interleaved samples left-justified in int16 / int32, "s16" | "s32", channels, sample_rate), decoded by the
    native reader (`bn_read_raw_batch`).  (s << (16 - bits)) / 32768 is the float libsndfile hands the reference."""
    from birdnet_stm32.audio import reader as _rd

    info = _rd.probe(path, float(max_seconds or 0.0))
    if info.status == _rd.RD_UNREADABLE or info.container != 1:
        raise UnsupportedAudio(f"{path}: not a decodable FLAC stream")
    nbytes = int(info.n_frames) * int(info.channels) * (2 if _rd.FMT_NAMES[info.fmt] == "s16" else 4)
    buf = np.empty(nbytes + 16, dtype=np.uint8)
    n, items = _rd.read_raw_batch([path], buf, max_seconds=float(max_seconds or 0.0), threads=1)
    if n != 1 or items[0] is None:
        raise UnsupportedAudio(f"{path}: FLAC decode failed")
    raw, kind, ch, sr = items[0]
    return raw.copy(), kind, ch, sr


def read_wav_pcm16(path: str, max_frames: int | None = None) -> tuple[np.ndarray, int]:
    """Mono PCM of <= 16 bits (WAV, FLAC) -> (int16 [frames], sample_rate).  Anything else raises UnsupportedAudio."""
    if is_flac(path):
        from birdnet_stm32.audio import reader as _rd

        info = _rd.probe(path, 0.0)
        if info.status == _rd.RD_UNREADABLE or info.channels != 1 or _rd.FMT_NAMES.get(info.fmt) != "s16":
            raise UnsupportedAudio(f"{path}: FLAC stream is not mono with <= 16 bits per sample")
        secs = None if max_frames is None else (max_frames + 0.5) / float(info.sample_rate)
        raw, _, _, sr = read_flac_frames(path, secs)
        return (raw if max_frames is None else raw[:max_frames]).astype(np.int16, copy=False), sr
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


def read_wav_frames(path: str, max_seconds: float | None = None) -> tuple[np.ndarray, str, int, int]:
    """RIFF/WAVE -> (interleaved raw samples, format, channels, sample_rate) without converting anything.

    format is one of `s16 | s24 | s32 | f32 | u8` (`_lib.BN_SAMPLE_FORMAT`); the array is int16 / uint8 (3 bytes per
    sample, packed) / int32 / float32 / uint8 with `frames * channels` samples.  `max_seconds` limits the read to
    `int(min(frames, max_seconds * sr))` frames from the start, the window `load_audio_window` reads (`io.py:97-109`).
    """
    if is_flac(path):
        return read_flac_frames(path, max_seconds)
    with open(path, "rb") as fh:
        head = fh.read(12)
        if len(head) < 12 or head[:4] != b"RIFF" or head[8:12] != b"WAVE":
            raise UnsupportedAudio(f"{path}: not a RIFF/WAVE or FLAC file")
        fmt = None
        while True:
            hdr = fh.read(8)
            if len(hdr) < 8:
                raise UnsupportedAudio(f"{path}: no data chunk")
            cid, size = hdr[:4], struct.unpack("<I", hdr[4:])[0]
            if cid == b"fmt ":
                body = fh.read(size + (size & 1))
                tag, ch, sr, _, _, bits = struct.unpack("<HHIIHH", body[:16])
                if tag == 0xFFFE and len(body) >= 26:
                    tag = struct.unpack("<H", body[24:26])[0]
                fmt = (tag, ch, sr, bits)
            elif cid == b"data":
                if fmt is None:
                    raise UnsupportedAudio(f"{path}: data chunk before fmt chunk")
                tag, ch, sr, bits = fmt
                kind = {(1, 8): "u8", (1, 16): "s16", (1, 24): "s24", (1, 32): "s32", (3, 32): "f32"}.get((tag, bits))
                if kind is None or ch < 1 or sr <= 0:
                    raise UnsupportedAudio(f"{path}: WAVE format tag {tag} with {bits} bits is not decoded on this path")
                bps = bits // 8
                frames = size // (bps * ch)
                if max_seconds and max_seconds > 0:
                    frames = int(min(frames, float(max_seconds) * sr))
                raw = fh.read(frames * bps * ch)
                frames = len(raw) // (bps * ch)
                raw = raw[: frames * bps * ch]
                dt = {"u8": np.uint8, "s16": "<i2", "s24": np.uint8, "s32": "<i4", "f32": "<f4"}[kind]
                return np.frombuffer(raw, dtype=dt), kind, ch, sr
            else:
                fh.seek(size + (size & 1), 1)


def load_pcm16_window(path: str, sample_rate: int, max_duration: float | None = 60) -> tuple[np.ndarray, np.float32]:
    """(int16 mono window, file peak) -- the device-path twin of `load_audio_window`.

    peak = max|pcm / 32768| in float32 (0 for silence, meaning "do not normalise").
    """
    limit = None if not max_duration or max_duration <= 0 else int(float(max_duration) * sample_rate)
    pcm, sr0 = read_wav_pcm16(path, limit)
    if sr0 != sample_rate:
        raise UnsupportedAudio(f"{path}: sample rate {sr0} != model rate {sample_rate} (host resampling is a next-step row)")
    # max|s / 32768| in float32 == max|s| / 32768 (division by a power of two is exact and monotone): two int16 passes
    # instead of a float32 copy, abs and max (2.0 -> 0.2 ms per 60 s file)
    peak = np.float32(max(int(pcm.max()), -int(pcm.min()))) / np.float32(32768.0) if pcm.size else np.float32(0)
    return pcm, peak


def load_audio_window(path: str, sample_rate: int = 24000, max_duration: float | None = 30,
                      chunk_duration: float = 3.0, random_offset: bool = False) -> np.ndarray:
    """Reference-compatible float view: mono float32 in [-1, 1], peak normalised; empty on error."""
    try:
        pcm, peak = load_pcm16_window(path, sample_rate, max_duration)
    except UnsupportedAudio:
        # other rate / channels / sample format: decode + mix + resample + normalise on the device (no host resampler here)
        try:
            raw, kind, ch, sr0 = read_wav_frames(path, max_duration)
        except Exception:
            return np.empty((0,), dtype=np.float32)
        from birdnet_stm32.audio.ingest import shared_ingest

        return shared_ingest().window(raw, kind, ch, sr0, sample_rate, normalize=True)
    except Exception:
        return np.empty((0,), dtype=np.float32)
    y = pcm.astype(np.float32) / np.float32(32768.0)
    if peak > 0:
        y = y / peak
    return y.astype(np.float32, copy=False)


def fast_resample(y: np.ndarray, sr_in: int, sr_out: int) -> np.ndarray:
    """`scipy.signal.resample_poly` twin of the reference (`audio/io.py:14-30`) on the device: float32 mono in, float32 out,
    bit-identical to scipy (see `audio/ingest.py`).  Equal rates return the input as float32."""
    if sr_in == sr_out:
        return np.asarray(y).astype(np.float32, copy=False)
    from birdnet_stm32.audio.ingest import shared_ingest

    return shared_ingest().window(np.ascontiguousarray(y, dtype=np.float32), "f32", 1, int(sr_in), int(sr_out), normalize=False)


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


def prefetch_ordered(fn, items, workers: int = 8, depth: int | None = None):
    """Yield `fn(item)` for every item, in order, computed by a pool of reader threads that runs ahead of the
    consumer by at most `depth` items.  File reads and the numpy copies behind `fn` release the GIL, so the GPU
    path is fed while the next files are still being read (the reference reads, decodes and resamples each file
    serially before its `predict` calls, `evaluation/metrics.py:117-147`).  `workers <= 1` is a plain loop."""
    items = list(items)
    if workers <= 1 or len(items) <= 1:
        for it in items:
            yield fn(it)
        return
    from collections import deque
    from concurrent.futures import ThreadPoolExecutor

    depth = depth or 4 * workers
    with ThreadPoolExecutor(max_workers=workers, thread_name_prefix="bn-io") as pool:
        pending: deque = deque()
        nxt = 0
        while nxt < len(items) and len(pending) < depth:
            pending.append(pool.submit(fn, items[nxt]))
            nxt += 1
        while pending:
            fut = pending.popleft()
            if nxt < len(items):
                pending.append(pool.submit(fn, items[nxt]))
                nxt += 1
            yield fut.result()
