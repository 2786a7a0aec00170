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


class FeatureExtractor:
    """Batched GPU twin of `get_spectrogram_from_audio` for the precomputed-spectrogram frontends.

    Reference `audio/spectrogram.py:24-149` computes one chunk at a time with librosa on the host; this object
    computes `[B, rows, spec_width]` features for a batch of PCM16 chunks with the CUDA kernels of
    `csrc/bn_features.cu` (through `bn_features_create` / `bn_features_pcm16`, `include/bn_features.h`).  Same
    arguments and the same modes: `mode` in {"mel", "log_mel", "mfcc"}, `mag_scale` in {"none", "pwl", "pcen", "db"}
    (used in "mel" mode only, as in the reference); `mel_bins <= 0` / `mode="linear"` is the hybrid model input and is
    served by `GpuRunner.frontend`.  There is no host fallback.
    """

    def __init__(self, sample_rate: int = 24000, chunk_len: int = 72000, n_fft: int = 512, mel_bins: int = 64,
                 spec_width: int = 256, mag_scale: str = "none", mode: str = "mel", n_mfcc: int = 20, device: int = 0):
        import ctypes as C

        from birdnet_stm32 import _lib
        from birdnet_stm32.audio import mel as melmod

        if mode not in _lib.BN_FEAT_MODE:
            raise ValueError(f"Unsupported spectrogram mode for the GPU feature path: {mode}")
        if mag_scale not in _lib.BN_MAG_SCALE:
            raise ValueError(f"Unsupported mag_scale: {mag_scale}")
        if mel_bins <= 0:
            raise ValueError("linear spectrograms are the hybrid model input: use GpuRunner.frontend")
        self._L = _lib.load()
        self._lib = _lib
        hop = chunk_len // spec_width
        self.params = _lib.BnFeatParams(sample_rate=sample_rate, chunk_len=chunk_len, n_fft=n_fft, spec_width=spec_width,
                                        n_mels=mel_bins, mode=_lib.BN_FEAT_MODE[mode], mag_scale=_lib.BN_MAG_SCALE[mag_scale],
                                        n_mfcc=n_mfcc, pcen_b=melmod.pcen_coefficient(sample_rate, hop))
        # librosa.feature.melspectrogram(..., fmin=150, fmax=sample_rate // 2, htk=False, norm="slaney")
        self.mel_basis = np.ascontiguousarray(melmod.mel_filterbank(sample_rate, n_fft, mel_bins, 150.0, float(sample_rate // 2)))
        self.dct = np.ascontiguousarray(melmod.dct_matrix(n_mfcc, mel_bins)) if mode == "mfcc" else None
        h = C.c_void_p()
        _lib.check(self._L.bn_features_create(C.byref(self.params), self.mel_basis.ctypes.data_as(C.c_void_p),
                                              self.dct.ctypes.data_as(C.c_void_p) if self.dct is not None else None,
                                              int(device), C.byref(h)))
        self._h = h
        self.rows = int(self._L.bn_features_rows(h))
        self.chunk_len, self.spec_width = chunk_len, spec_width

    def __call__(self, pcm: np.ndarray, peak=None) -> np.ndarray:
        """PCM16 `[B, T]` (+ per-chunk file peak `[B]`) -> float32 `[B, rows, spec_width]` in [0, 1]."""
        import ctypes as C

        pcm = np.ascontiguousarray(pcm, dtype=np.int16)
        if pcm.ndim == 1:
            pcm = pcm[None, :]
        if pcm.shape[1] != self.chunk_len:
            raise ValueError(f"chunks have {pcm.shape[1]} samples, extractor was built for {self.chunk_len}")
        B = pcm.shape[0]
        pk = None
        if peak is not None:
            pk = np.ascontiguousarray(np.broadcast_to(np.asarray(peak, dtype=np.float32), (B,)))
        out = np.empty((B, self.rows, self.spec_width), dtype=np.float32)
        self._lib.check(self._L.bn_features_pcm16(self._h, pcm.ctypes.data_as(C.c_void_p),
                                                  pk.ctypes.data_as(C.c_void_p) if pk is not None else None, B,
                                                  out.ctypes.data_as(C.c_void_p), None))
        return out

    def run_device(self, pcm_ptr: int, peak_ptr: int, B: int, out_ptr: int, stream: int = 0) -> None:
        """Device-pointer entry (no copies, enqueued on `stream`)."""
        self._lib.check(self._L.bn_features_pcm16(self._h, pcm_ptr, peak_ptr or None, int(B), out_ptr, stream or None))

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._L.bn_features_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def get_spectrograms_from_pcm16(pcm: np.ndarray, peak=None, sample_rate: int = 24000, n_fft: int = 512, mel_bins: int = 64,
                                spec_width: int = 256, mag_scale: str = "none", mode: str = "mel", n_mfcc: int = 20,
                                device: int = 0) -> np.ndarray:
    """One-shot convenience wrapper around :class:`FeatureExtractor` (argument names of the reference function)."""
    pcm = np.asarray(pcm)
    if pcm.ndim == 1:
        pcm = pcm[None, :]
    fx = FeatureExtractor(sample_rate, pcm.shape[1], n_fft, mel_bins, spec_width, mag_scale, mode, n_mfcc, device)
    try:
        return fx(pcm, peak)
    finally:
        fx.close()


def audio_to_pcm16_peak(audio: np.ndarray) -> tuple[np.ndarray, np.float32]:
    """Float waveform -> (int16 samples, peak) such that the device's `(pcm / 32768) / peak` reproduces it.

    The feature kernels take PCM16 plus a per-chunk divisor.  A float chunk `a` with A = max|a| is stored as
    `pcm = round(a / A * 32767)` and `peak = 32767 / (32768 * A)`, so the device sees `pcm * A / 32767`: `a` rounded to 16 bits
    of ITS OWN range (absolute error <= A / 65534).  After the min-max normalisation every frontend ends with, that is an
    error of about 1e-5 on the mel / PWL / log-mel / MFCC features, inside the 1e-4 frontend tolerance; the dB scaling amplifies
    the rounding of quiet bins to at most 5e-4, PCEN to 1e-4 (measured against the oracle in `tests/test_features.py`).  Silence gives zeros and peak 0 (= no division)."""
    a = np.asarray(audio, dtype=np.float32).reshape(-1)
    amax = float(np.max(np.abs(a))) if a.size else 0.0
    if not amax > 0.0:
        return np.zeros(a.shape, dtype=np.int16), np.float32(0.0)
    pcm = np.round(a.astype(np.float64) / amax * 32767.0).astype(np.int16)
    return pcm, np.float32(32767.0 / (32768.0 * amax))


def get_spectrogram_from_audio(audio: np.ndarray, sample_rate: int = 24000, n_fft: int = 512, mel_bins: int = 64, spec_width: int = 256,
                               mag_scale: str = "none", mode: str = "mel", n_mfcc: int = 20, device: int = 0) -> np.ndarray:
    """The reference's per-chunk entry point (`audio/spectrogram.py:24-149`), same arguments and return shape
    `(mel_bins | n_mfcc, spec_width)` in [0, 1], computed by the CUDA feature kernels for one float chunk.

    Batches should use :class:`FeatureExtractor` (one launch for many chunks).  The linear mode (`mel_bins <= 0`, the hybrid
    model input) needs the model's engine and is served by `GpuRunner.frontend` / `frontend_wave`."""
    if mel_bins <= 0 or mode == "linear":
        raise ValueError("linear STFT magnitudes are the hybrid model input: use GpuRunner.frontend_wave(chunks)")
    pcm, peak = audio_to_pcm16_peak(audio)
    out = get_spectrograms_from_pcm16(pcm[None, :], np.array([peak], dtype=np.float32), sample_rate, n_fft, mel_bins, spec_width, mag_scale,
                                      mode, n_mfcc, device)
    return out[0]
