"""GpuIngest: device-side twin of the reference's `load_audio_window` arithmetic (`audio/io.py:112-128`).

The reference decodes a file window to float32 with libsndfile, averages the channels with numpy, resamples
with `scipy.signal.resample_poly` when the file rate differs from the model rate, divides by the peak and
cuts chunks, all on the host and per file.  `GpuIngest` hands the raw interleaved samples to
`bn_ingest_window` / `bn_ingest_chunks` (`include/bn_ingest.h`, `csrc/bn_ingest.cu`), which do the same
steps with hand-written CUDA kernels and leave float32 chunks in device memory for
`GpuRunner.predict_wave` / `infer_pool_wave_ptr`.  There is no host resampler in this package: without the
CUDA library or a device these calls raise.
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from birdnet_stm32 import _lib as L


def resample_ratio(sr_in: int, sr_out: int) -> tuple[int, int]:
    """(up, down) of `fast_resample` (`audio/io.py:26-29`)."""
    from math import gcd

    g = gcd(int(sr_in), int(sr_out))
    return int(sr_out) // g, int(sr_in) // g


def chunk_step(sample_rate: int, chunk_duration: float, chunk_overlap: float = 0.0) -> tuple[int, int]:
    """(chunk_len, step) of `split_audio_into_chunks` (`audio/io.py:152-163`)."""
    size = int(sample_rate * chunk_duration)
    overlap = max(0.0, min(chunk_overlap, chunk_duration - 0.1))
    return size, max(1, int(sample_rate * (chunk_duration - overlap)))


def _vp(a):
    return a.ctypes.data_as(C.c_void_p)


class GpuIngest:
    def __init__(self, device: int = 0):
        self._lib = L.load()
        self._h = C.c_void_p()
        L.check(self._lib.bn_ingest_create(int(device), C.byref(self._h)))
        self.device = device

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.bn_ingest_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launches(self) -> int:
        return int(self._lib.bn_ingest_launch_count(self._h))

    # -- geometry (host arithmetic only) ---------------------------------------------------
    def out_len(self, n_frames: int, sr_in: int, sr_out: int) -> int:
        return int(self._lib.bn_ingest_out_len(int(n_frames), int(sr_in), int(sr_out)))

    def num_chunks(self, n_samples: int, chunk_len: int, step: int) -> int:
        return int(self._lib.bn_ingest_num_chunks(int(n_samples), int(chunk_len), int(step)))

    def filter_taps(self, up: int, down: int) -> tuple[np.ndarray, int]:
        """The float32 taps handed to upfirdn (zero padded, times `up`) and the number of leading outputs dropped."""
        n, pre = C.c_int(), C.c_int()
        L.check(self._lib.bn_ingest_filter(int(up), int(down), None, 0, C.byref(n), C.byref(pre)))
        h = np.zeros((n.value,), dtype=np.float32)
        if n.value:
            L.check(self._lib.bn_ingest_filter(int(up), int(down), _vp(h), n.value, C.byref(n), C.byref(pre)))
        return h, pre.value

    @staticmethod
    def _frames(raw: np.ndarray, kind: str, channels: int) -> tuple[np.ndarray, int]:
        if kind not in L.BN_SAMPLE_FORMAT:
            raise ValueError(f"unknown sample format {kind!r}")
        a = np.ascontiguousarray(raw).reshape(-1)
        per = 3 if kind == "s24" else 1
        if a.size % (per * channels):
            raise ValueError("sample buffer is not a whole number of frames")
        return a, a.size // (per * channels)

    # -- host in, host out -------------------------------------------------------------------
    def window(self, raw: np.ndarray, kind: str, channels: int, sr_in: int, sr_out: int, normalize: bool = True,
               return_peak: bool = False):
        """`load_audio_window` on already-read samples: mono float32 at `sr_out`, peak-normalised."""
        a, n = self._frames(raw, kind, channels)
        out = np.empty((self.out_len(n, sr_in, sr_out),), dtype=np.float32)
        peak = np.zeros((1,), dtype=np.float32)
        if n:
            L.check(self._lib.bn_ingest_window(self._h, _vp(a), L.BN_SAMPLE_FORMAT[kind], n, int(channels), int(sr_in), int(sr_out),
                                               int(bool(normalize)), _vp(out), _vp(peak), None))
        return (out, float(peak[0])) if return_peak else out

    def chunks(self, raw: np.ndarray, kind: str, channels: int, sr_in: int, sr_out: int, chunk_len: int, step: int) -> np.ndarray:
        """`load_audio_file`: float32 chunks `[n, chunk_len]` on the host."""
        a, n = self._frames(raw, kind, channels)
        nc = self.num_chunks(self.out_len(n, sr_in, sr_out), chunk_len, step)
        out = np.empty((nc, chunk_len), dtype=np.float32)
        got = C.c_int()
        if n:
            L.check(self._lib.bn_ingest_chunks(self._h, _vp(a), L.BN_SAMPLE_FORMAT[kind], n, int(channels), int(sr_in), int(sr_out),
                                               int(chunk_len), int(step), _vp(out), nc, C.byref(got), None, None))
        return out

    # -- host in, device out (the evaluation path) ---------------------------------------------
    def chunks_to_ptr(self, raw: np.ndarray, kind: str, channels: int, sr_in: int, sr_out: int, chunk_len: int, step: int,
                      out_ptr: int, max_chunks: int) -> int:
        """Chunks written to device memory at `out_ptr` (float32 `[max_chunks, chunk_len]`); returns the count."""
        a, n = self._frames(raw, kind, channels)
        got = C.c_int()
        if n:
            L.check(self._lib.bn_ingest_chunks(self._h, _vp(a), L.BN_SAMPLE_FORMAT[kind], n, int(channels), int(sr_in), int(sr_out),
                                               int(chunk_len), int(step), C.c_void_p(out_ptr), int(max_chunks), C.byref(got), None, None))
        return got.value


_shared: GpuIngest | None = None


def shared_ingest(device: int = 0) -> GpuIngest:
    global _shared
    if _shared is None or _shared.device != device:
        _shared = GpuIngest(device)
    return _shared
