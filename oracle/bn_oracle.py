"""ctypes binding of the CPU ORACLE (`oracle/bn_oracle.c`).

TEST INFRASTRUCTURE ONLY.  Importable from `tests/`, `__graft_entry__.smoke()`
and `bench.py`'s cpu_baseline / `--impl reference` legs; the product package
(`birdnet-stm32_b200/birdnet_stm32`) never imports it.  Parity status: see the
header of `bn_oracle.c` ("PARITY UNPINNED" at the TFLite / librosa boundaries).
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libbn_oracle.so")
_lib = None

POOL_METHODS = {"avg": 0, "mean": 0, "average": 0, "max": 1, "lme": 2, "log_mean_exp": 2, "log_mean_exponential": 2}


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc, no GPU needed)."""
    src = os.path.join(_HERE, "bn_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "libbn_oracle.so"], check=True, capture_output=True)
    return _LIB_PATH


def use_native_build() -> str:
    """Switch to `libbn_oracle_native.so` (-O3 -march=native), compiled on THIS machine -- the build bench.py times for
    `cpu_baseline` / `--impl reference`.  Must be called before the first `lib()` use; falls back to the portable build."""
    global _LIB_PATH, _lib
    native = os.path.join(_HERE, "libbn_oracle_native.so")
    try:
        subprocess.run(["make", "-C", _HERE, "-B", "native"], check=True, capture_output=True)
        if _lib is None or _LIB_PATH != native:
            _LIB_PATH, _lib = native, None
    except Exception:
        pass
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        L.bno_load.restype = C.c_void_p
        L.bno_load.argtypes = [C.c_void_p, C.c_size_t]
        L.bno_free.argtypes = [C.c_void_p]
        L.bno_set_option.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.bno_num_classes.argtypes = [C.c_void_p]
        L.bno_input_elems.restype = C.c_long
        L.bno_input_elems.argtypes = [C.c_void_p]
        L.bno_tensor_bytes.restype = C.c_long
        L.bno_tensor_bytes.argtypes = [C.c_void_p, C.c_int]
        L.bno_run_graph.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        L.bno_frontend_hybrid.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.bno_frontend_hybrid_f32.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.bno_frontend_raw.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.bno_pool.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p]
        L.bno_logistic_lut.argtypes = [C.c_float, C.c_int, C.c_float, C.c_int, C.c_void_p]
        L.bno_mbqm.restype = C.c_int32
        L.bno_mbqm.argtypes = [C.c_int32, C.c_int32, C.c_int, C.c_int]
        L.bno_rq_fast.restype = C.c_int32
        L.bno_rq_fast.argtypes = [C.c_int32, C.c_int32, C.c_int]
        L.bno_last_error.restype = C.c_char_p
        _lib = L
    return _lib


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class OracleModel:
    """The int8 graph executed on the CPU from the exported blob."""

    def __init__(self, blob: bytes, rounding: int = 0, mean_variant: int = 0, threads: int = 0):
        self._L = lib()
        self._blob = bytes(blob)
        self._h = self._L.bno_load(self._blob, len(self._blob))
        if not self._h:
            raise RuntimeError(self._L.bno_last_error().decode())
        self._L.bno_set_option(self._h, rounding, mean_variant, threads)
        self.num_classes = self._L.bno_num_classes(self._h)
        self.input_elems = self._L.bno_input_elems(self._h)

    def set_option(self, rounding: int = 0, mean_variant: int = 0, threads: int = 0):
        self._L.bno_set_option(self._h, rounding, mean_variant, threads)

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.bno_free(self._h)
            self._h = None

    def predict(self, x_batch: np.ndarray) -> np.ndarray:
        """Same protocol as the reference runners (`models/runners.py:82-95`)."""
        out, _ = self.run(x_batch)
        return out

    def run(self, x_batch: np.ndarray, tap_id: int = -1, tap_dtype=np.int8):
        x = np.ascontiguousarray(x_batch, dtype=np.float32)
        B = x.shape[0]
        if x.size != B * self.input_elems:
            raise ValueError(f"input has {x.size // max(B, 1)} elements per chunk, model expects {self.input_elems}")
        out = np.empty((B, self.num_classes), dtype=np.float32)
        tap = None
        if tap_id >= 0:
            nb = self._L.bno_tensor_bytes(self._h, tap_id)
            if nb < 0:
                raise KeyError(f"no tensor {tap_id}")
            tap = np.empty((B, nb), dtype=np.uint8)
        rc = self._L.bno_run_graph(self._h, _ptr(x), B, _ptr(out), tap_id, _ptr(tap) if tap is not None else None)
        if rc:
            raise RuntimeError(self._L.bno_last_error().decode())
        if tap is not None:
            tap = tap.view(tap_dtype)
        return out, tap


def frontend_hybrid(pcm: np.ndarray, peak: np.ndarray | None, n_fft: int, hop: int, spec_width: int,
                    threads: int = 0) -> np.ndarray:
    """PCM16 [B,T] (+ per-chunk file peak) -> float32 [B, n_fft/2+1, spec_width, 1]."""
    L = lib()
    pcm = np.ascontiguousarray(pcm, dtype=np.int16)
    B, T = pcm.shape
    pk = np.ascontiguousarray(peak, dtype=np.float32) if peak is not None else None
    out = np.empty((B, n_fft // 2 + 1, spec_width, 1), dtype=np.float32)
    rc = L.bno_frontend_hybrid(_ptr(pcm), _ptr(pk) if pk is not None else None, B, T, n_fft, hop, spec_width, _ptr(out), threads)
    if rc:
        raise RuntimeError(L.bno_last_error().decode())
    return out


def frontend_hybrid_f32(wav: np.ndarray, n_fft: int, hop: int, spec_width: int, threads: int = 0) -> np.ndarray:
    L = lib()
    wav = np.ascontiguousarray(wav, dtype=np.float32)
    B, T = wav.shape
    out = np.empty((B, n_fft // 2 + 1, spec_width, 1), dtype=np.float32)
    rc = L.bno_frontend_hybrid_f32(_ptr(wav), B, T, n_fft, hop, spec_width, _ptr(out), threads)
    if rc:
        raise RuntimeError(L.bno_last_error().decode())
    return out


def frontend_raw(pcm: np.ndarray, peak: np.ndarray | None) -> np.ndarray:
    L = lib()
    pcm = np.ascontiguousarray(pcm, dtype=np.int16)
    B, T = pcm.shape
    pk = np.ascontiguousarray(peak, dtype=np.float32) if peak is not None else None
    out = np.empty((B, T, 1), dtype=np.float32)
    L.bno_frontend_raw(_ptr(pcm), _ptr(pk) if pk is not None else None, B, T, _ptr(out))
    return out


def pool_scores(chunk_scores: np.ndarray, method: str = "average", beta: float = 10.0) -> np.ndarray:
    """Restatement of `evaluation/pooling.py:25-47` (same errors, same empty behaviour)."""
    method = method.lower()
    chunk_scores = np.asarray(chunk_scores)
    if chunk_scores.ndim != 2:
        raise ValueError("chunk_scores must be [N_chunks, C]")
    if method not in POOL_METHODS:
        raise ValueError(f"Unsupported pooling method: {method}")
    s = np.ascontiguousarray(chunk_scores, dtype=np.float32)
    out = np.zeros((s.shape[1],), dtype=np.float32)
    rc = lib().bno_pool(_ptr(s), s.shape[0], s.shape[1], POOL_METHODS[method], float(beta), _ptr(out))
    if rc:
        raise RuntimeError(lib().bno_last_error().decode())
    return out


def logistic_lut(in_scale, in_zp, out_scale, out_zp) -> np.ndarray:
    lut = np.zeros(256, dtype=np.int8)
    lib().bno_logistic_lut(float(in_scale), int(in_zp), float(out_scale), int(out_zp), _ptr(lut))
    return lut


def mbqm(x: int, qm: int, shift: int, rounding: int = 0) -> int:
    return int(lib().bno_mbqm(int(x), int(qm), int(shift), int(rounding)))


def rq_fast(x: int, qm: int, n: int) -> int:
    return int(lib().bno_rq_fast(int(x), int(qm), int(n)))
