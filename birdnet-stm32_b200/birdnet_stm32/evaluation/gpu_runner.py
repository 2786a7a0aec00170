"""GpuRunner: the B200 engine behind the reference's runner protocol.

Sits beside the reference's `TFLiteRunner` (`birdnet_stm32/models/runners.py:48-95`):
same `predict(x_batch) -> float32 [B, num_classes]` contract, any batch size per
call, fresh output array per call, Python exceptions on error.  On top of that
it exposes the full hot path the reference spreads over
`make_chunks_for_file` + `predict` + `pool_scores`
(`evaluation/metrics.py:55-61,129-143`): PCM16 in, pooled file scores out.

All compute happens in `libbn_b200.so` (hand-written sm_100a kernels) through
the C ABI of `include/bn_engine.h`; this module only moves pointers.
"""

from __future__ import annotations

import ctypes as C
import json
import os

import numpy as np

from birdnet_stm32 import _lib as L
from birdnet_stm32.conversion.export_blob import export_blob


def _find_config(model_path: str) -> dict:
    """Locate `<stem>_model_config.json` next to the model like the reference CLI does
    (`cli/evaluate.py:86-94`)."""
    stem = os.path.splitext(model_path)[0]
    for cand in (stem + "_model_config.json", stem.replace("_quantized", "") + "_model_config.json"):
        if os.path.exists(cand):
            with open(cand) as fh:
                return json.load(fh)
    return {}


def _vp(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class PinnedArray:
    """A numpy view over page-locked host memory from `bn_host_alloc`."""

    def __init__(self, shape, dtype):
        self._lib = L.load()
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self._ptr = self._lib.bn_host_alloc(max(self.nbytes, 1))
        if not self._ptr:
            raise L.EngineError(-3, self._lib.bn_last_error().decode())
        buf = (C.c_char * max(self.nbytes, 1)).from_address(self._ptr)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def free(self):
        if self._ptr:
            self.array = None
            self._lib.bn_host_free(self._ptr)
            self._ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class GpuRunner:
    """Batched inference on one B200 through `libbn_b200.so`.

    Args:
        model: path to a `.tflite` (flattened on the fly), a `.b200blob`, or blob bytes.
        model_config: the `_model_config.json` dict; looked up next to the model if omitted.
        device: CUDA device ordinal.
        wave: chunks per internal wave (workspace size); None keeps the engine default.
    """

    def __init__(self, model, model_config: dict | None = None, device: int = 0, wave: int | None = None):
        self._lib = L.load()
        self._h = C.c_void_p()
        if isinstance(model, (bytes, bytearray, memoryview)):
            blob = bytes(model)
            self.cfg = dict(model_config or {})
        else:
            path = os.fspath(model)
            self.cfg = dict(model_config) if model_config is not None else _find_config(path)
            if path.lower().endswith(".tflite"):
                blob = export_blob(path, self.cfg)
            else:
                with open(path, "rb") as fh:
                    blob = fh.read()
        self._blob = blob
        L.check(self._lib.bn_create(blob, len(blob), int(device), C.byref(self._h)))
        if wave:
            self.set_option(L.BN_OPT_WAVE, int(wave))
        self.info = self.query()
        self.num_classes = self.info.num_classes
        self.input_elems = self.info.input_elems
        self.device = device

    # -- lifecycle -----------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.bn_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def query(self) -> L.BnInfo:
        info = L.BnInfo()
        L.check(self._lib.bn_query(self._h, C.byref(info)))
        return info

    def set_option(self, key: int, value: int):
        L.check(self._lib.bn_set_option(self._h, key, value))

    @property
    def launches(self) -> int:
        return int(self._lib.bn_launch_count(self._h))

    # -- reference runner protocol -------------------------------------------
    def predict(self, x_batch: np.ndarray) -> np.ndarray:
        """`TFLiteRunner.predict` twin: float32 `[B, ...]` model input -> float32 `[B, C]`."""
        x = np.ascontiguousarray(x_batch, dtype=np.float32)
        B = int(x.shape[0]) if x.ndim else 0
        if B and x.size != B * self.input_elems:
            raise ValueError(
                f"Cannot set tensor: got {tuple(x.shape)}, model expects {self.input_elems} elements per sample"
            )
        out = np.empty((B, self.num_classes), dtype=np.float32)
        if B:
            L.check(self._lib.bn_infer_spec_f32(self._h, _vp(x), B, _vp(out), None))
        return out

    # -- full hot path --------------------------------------------------------
    def _pcm_args(self, pcm, peak):
        pcm = np.ascontiguousarray(pcm, dtype=np.int16)
        if pcm.ndim != 2 or pcm.shape[1] != self.info.chunk_len:
            raise ValueError(f"pcm must be int16 [B, {self.info.chunk_len}], got {pcm.shape}")
        pk = None
        if peak is not None:
            pk = np.ascontiguousarray(peak, dtype=np.float32)
            if pk.shape != (pcm.shape[0],):
                raise ValueError("peak must be float32 [B]")
        return pcm, pk

    def frontend(self, pcm: np.ndarray, peak: np.ndarray | None = None) -> np.ndarray:
        """PCM16 `[B, T]` -> the model's float input (hybrid: `[B, 257, 256, 1]`); K1 only."""
        pcm, pk = self._pcm_args(pcm, peak)
        B = pcm.shape[0]
        out = np.empty((B, self.info.fft_bins, self.info.spec_width, 1), dtype=np.float32)
        if B:
            L.check(self._lib.bn_frontend_pcm16(self._h, _vp(pcm), _vp(pk) if pk is not None else None, B, _vp(out), None))
        return out

    def predict_pcm16(self, pcm: np.ndarray, peak: np.ndarray | None = None) -> np.ndarray:
        """PCM16 `[B, T]` (+ per-chunk file peak) -> chunk scores float32 `[B, C]`."""
        pcm, pk = self._pcm_args(pcm, peak)
        B = pcm.shape[0]
        out = np.empty((B, self.num_classes), dtype=np.float32)
        if B:
            L.check(self._lib.bn_infer_pcm16(self._h, _vp(pcm), _vp(pk) if pk is not None else None, B, _vp(out), None))
        return out

    def predict_pooled(self, pcm: np.ndarray, peak, file_offsets, pooling: str = "average", beta: float = 10.0) -> np.ndarray:
        """PCM16 chunks of F files -> pooled float32 `[F, C]` (pooling on the device)."""
        method = pooling.lower()
        if method not in L.BN_POOL:
            raise ValueError(f"Unsupported pooling method: {pooling}")
        pcm, pk = self._pcm_args(pcm, peak)
        offs = np.ascontiguousarray(file_offsets, dtype=np.int32)
        F = offs.size - 1
        if F < 0 or (F >= 0 and int(offs[-1]) != pcm.shape[0]):
            raise ValueError("file_offsets must be [F+1] with file_offsets[F] == number of chunks")
        out = np.empty((F, self.num_classes), dtype=np.float32)
        L.check(self._lib.bn_infer_pool(self._h, _vp(pcm), _vp(pk) if pk is not None else None, _vp(offs), F,
                                        L.BN_POOL[method], float(beta), _vp(out), None))
        return out

    # -- float32 waveform chunks (files the device ingest resampled / mixed, audio/ingest.py) --------
    def _wave_args(self, wave, peak):
        w = np.ascontiguousarray(wave, dtype=np.float32)
        if w.ndim != 2 or w.shape[1] != self.info.chunk_len:
            raise ValueError(f"wave must be float32 [B, {self.info.chunk_len}], got {w.shape}")
        pk = None
        if peak is not None:
            pk = np.ascontiguousarray(peak, dtype=np.float32)
            if pk.shape != (w.shape[0],):
                raise ValueError("peak must be float32 [B]")
        return w, pk

    def frontend_wave(self, wave: np.ndarray, peak: np.ndarray | None = None) -> np.ndarray:
        """float32 chunks `[B, T]` -> the model's float input; what `make_chunks_for_file` computes per chunk."""
        w, pk = self._wave_args(wave, peak)
        B = w.shape[0]
        out = np.empty((B, self.info.fft_bins, self.info.spec_width, 1), dtype=np.float32)
        if B:
            L.check(self._lib.bn_frontend_wave_f32(self._h, _vp(w), _vp(pk) if pk is not None else None, B, _vp(out), None))
        return out

    def predict_wave(self, wave: np.ndarray, peak: np.ndarray | None = None) -> np.ndarray:
        """float32 chunks `[B, T]` -> chunk scores float32 `[B, C]`."""
        w, pk = self._wave_args(wave, peak)
        B = w.shape[0]
        out = np.empty((B, self.num_classes), dtype=np.float32)
        if B:
            L.check(self._lib.bn_infer_wave_f32(self._h, _vp(w), _vp(pk) if pk is not None else None, B, _vp(out), None))
        return out

    def predict_pooled_wave(self, wave: np.ndarray, peak, file_offsets, pooling: str = "average", beta: float = 10.0) -> np.ndarray:
        method = pooling.lower()
        if method not in L.BN_POOL:
            raise ValueError(f"Unsupported pooling method: {pooling}")
        w, pk = self._wave_args(wave, peak)
        offs = np.ascontiguousarray(file_offsets, dtype=np.int32)
        F = offs.size - 1
        if F < 0 or int(offs[-1]) != w.shape[0]:
            raise ValueError("file_offsets must be [F+1] with file_offsets[F] == number of chunks")
        out = np.empty((F, self.num_classes), dtype=np.float32)
        L.check(self._lib.bn_infer_pool_wave_f32(self._h, _vp(w), _vp(pk) if pk is not None else None, _vp(offs), F,
                                                 L.BN_POOL[method], float(beta), _vp(out), None))
        return out

    def infer_pool_wave_ptr(self, wave_ptr: int, peak_ptr: int | None, offs_ptr: int, F: int, pooling: str, beta: float,
                            out_ptr: int, stream: int | None = None):
        L.check(self._lib.bn_infer_pool_wave_f32(self._h, C.c_void_p(wave_ptr), C.c_void_p(peak_ptr) if peak_ptr else None,
                                                 C.c_void_p(offs_ptr), F, L.BN_POOL[pooling.lower()], float(beta),
                                                 C.c_void_p(out_ptr), C.c_void_p(stream) if stream else None))

    def infer_wave_ptr(self, wave_ptr: int, peak_ptr: int | None, B: int, scores_ptr: int, stream: int | None = None):
        L.check(self._lib.bn_infer_wave_f32(self._h, C.c_void_p(wave_ptr), C.c_void_p(peak_ptr) if peak_ptr else None, B,
                                            C.c_void_p(scores_ptr), C.c_void_p(stream) if stream else None))

    def pool_scores(self, chunk_scores: np.ndarray, file_offsets, pooling: str = "average", beta: float = 10.0) -> np.ndarray:
        method = pooling.lower()
        if method not in L.BN_POOL:
            raise ValueError(f"Unsupported pooling method: {pooling}")
        s = np.ascontiguousarray(chunk_scores, dtype=np.float32)
        if s.ndim != 2:
            raise ValueError("chunk_scores must be [N_chunks, C]")
        offs = np.ascontiguousarray(file_offsets, dtype=np.int32)
        F = offs.size - 1
        out = np.zeros((F, s.shape[1]), dtype=np.float32)
        L.check(self._lib.bn_pool_scores(self._h, _vp(s), _vp(offs), F, s.shape[1], L.BN_POOL[method], float(beta), _vp(out), None))
        return out

    # -- device-pointer entry points (torch CUDA tensors or raw pointers) ------
    def infer_pcm16_ptr(self, pcm_ptr: int, peak_ptr: int | None, B: int, scores_ptr: int, stream: int | None = None):
        L.check(self._lib.bn_infer_pcm16(self._h, C.c_void_p(pcm_ptr), C.c_void_p(peak_ptr) if peak_ptr else None, B,
                                         C.c_void_p(scores_ptr), C.c_void_p(stream) if stream else None))

    def infer_pool_ptr(self, pcm_ptr: int, peak_ptr: int | None, offs_ptr: int, F: int, pooling: str, beta: float,
                       out_ptr: int, stream: int | None = None):
        L.check(self._lib.bn_infer_pool(self._h, C.c_void_p(pcm_ptr), C.c_void_p(peak_ptr) if peak_ptr else None,
                                        C.c_void_p(offs_ptr), F, L.BN_POOL[pooling.lower()], float(beta), C.c_void_p(out_ptr),
                                        C.c_void_p(stream) if stream else None))

    def infer_spec_ptr(self, spec_ptr: int, B: int, scores_ptr: int, stream: int | None = None):
        L.check(self._lib.bn_infer_spec_f32(self._h, C.c_void_p(spec_ptr), B, C.c_void_p(scores_ptr),
                                            C.c_void_p(stream) if stream else None))

    def frontend_ptr(self, pcm_ptr: int, peak_ptr: int | None, B: int, spec_ptr: int, stream: int | None = None):
        L.check(self._lib.bn_frontend_pcm16(self._h, C.c_void_p(pcm_ptr), C.c_void_p(peak_ptr) if peak_ptr else None, B,
                                            C.c_void_p(spec_ptr), C.c_void_p(stream) if stream else None))

    def profile(self, on: bool = True, reset: bool = False):
        """Switch per-kernel CUDA-event timing on/off (BN_OPT_PROFILE)."""
        self.set_option(L.BN_OPT_PROFILE, 2 if (on and reset) else int(on))

    def profile_read(self) -> dict:
        """{kernel name: (total ms, launches)} accumulated while profiling was on."""
        out = {}
        name = C.create_string_buffer(64)
        ms, cnt = C.c_double(), C.c_int64()
        i = 0
        while self._lib.bn_profile_read(self._h, i, name, 64, C.byref(ms), C.byref(cnt)) == 0:
            out[name.value.decode()] = (ms.value, cnt.value)
            i += 1
        return out

    # -- debug taps -------------------------------------------------------------
    def dump_tensor(self, tfl_tensor_id: int, nbytes: int, dtype=np.int8) -> np.ndarray:
        """Tensor `tfl_tensor_id` of the last wave as a flat array (see BN_OPT_FORCE_GENERIC)."""
        out = np.empty((nbytes,), dtype=np.uint8)
        L.check(self._lib.bn_dump_tensor(self._h, int(tfl_tensor_id), _vp(out), nbytes))
        return out.view(dtype)
