"""Native batch reader: many WAV files -> PCM16 chunks in one (pinned) buffer, read by a pool of C++ threads.

Twin of the per-file host work the reference does before inference (`audio/io.py:90-174`, the loop of
`evaluation/metrics.py:117-147`); see `include/bn_reader.h` / `csrc/bn_reader.cu`.  Pure host code: it works without a
GPU (the buffer is then ordinary memory), which is how the CPU tests exercise it.
"""

from __future__ import annotations

import ctypes as C
import os

import numpy as np

from birdnet_stm32 import _lib as L

RD_OK, RD_NEEDS_INGEST, RD_UNREADABLE = 0, 1, 2
FMT_NAMES = {v: k for k, v in L.BN_SAMPLE_FORMAT.items()}


def probe(path: str, max_seconds: float = 0.0) -> L.BnReaderFile:
    out = L.BnReaderFile()
    L.check(L.load().bn_wav_probe(os.fsencode(path), float(max_seconds), C.byref(out)))
    return out


def read_pcm16_batch(paths: list[str], sample_rate: int, chunk_len: int, step: int, out: np.ndarray, max_seconds: float = 60.0,
                     threads: int = 8):
    """Fill `out` (int16 [cap, chunk_len], C-contiguous) with the chunks of `paths` in order.

    Returns (files_consumed, chunks_used, info) where info is a ctypes array of `BnReaderFile` for the consumed files:
    status RD_OK files own `n_chunks` consecutive rows (in order), RD_NEEDS_INGEST files must go through the device ingest,
    RD_UNREADABLE files are skipped.  files_consumed < len(paths) when the buffer is full."""
    if out.dtype != np.int16 or out.ndim != 2 or out.shape[1] != chunk_len or not out.flags.c_contiguous:
        raise ValueError("out must be a C-contiguous int16 [cap, chunk_len] array")
    n = len(paths)
    arr = (C.c_char_p * max(n, 1))(*[os.fsencode(p) for p in paths])
    info = (L.BnReaderFile * max(n, 1))()
    used = C.c_int(0)
    rc = L.load().bn_read_pcm16_batch(arr, n, int(sample_rate), int(chunk_len), int(step), float(max_seconds or 0.0),
                                      out.ctypes.data_as(C.c_void_p), int(out.shape[0]), int(threads), info, C.byref(used))
    if rc < 0:
        L.check(rc)
    return rc, used.value, info


_RAW_DTYPE = {"s16": "<i2", "s24": np.uint8, "s32": "<i4", "f32": "<f4", "u8": np.uint8}


def read_raw_batch(paths: list[str], out: np.ndarray, max_seconds: float = 60.0, threads: int = 8):
    """Raw interleaved samples of `paths` (in order) into the byte buffer `out` (uint8, C-contiguous).

    Returns (files_consumed, items) where items[i] is None for an unreadable file or
    `(raw, kind, channels, sample_rate)` -- the tuple `audio.io.read_wav_frames` returns, with `raw` a view into `out`."""
    if out.dtype != np.uint8 or out.ndim != 1 or not out.flags.c_contiguous:
        raise ValueError("out must be a C-contiguous uint8 vector")
    n = len(paths)
    arr = (C.c_char_p * max(n, 1))(*[os.fsencode(p) for p in paths])
    info = (L.BnReaderFile * max(n, 1))()
    offs = (C.c_int64 * (n + 1))()
    rc = L.load().bn_read_raw_batch(arr, n, float(max_seconds or 0.0), out.ctypes.data_as(C.c_void_p), int(out.size), int(threads), info, offs)
    if rc < 0:
        L.check(rc)
    items = []
    for i in range(rc):
        fi = info[i]
        if fi.status == RD_UNREADABLE:
            items.append(None)
            continue
        kind = FMT_NAMES[fi.fmt]
        nbytes = fi.n_frames * fi.channels * {"s16": 2, "s24": 3, "s32": 4, "f32": 4, "u8": 1}[kind]
        raw = out[offs[i]:offs[i] + nbytes].view(_RAW_DTYPE[kind])
        items.append((raw, kind, int(fi.channels), int(fi.sample_rate)))
    return rc, items
