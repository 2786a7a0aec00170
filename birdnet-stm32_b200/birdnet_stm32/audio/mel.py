"""Host-side constants of the precomputed-spectrogram ("librosa") frontends.

The reference builds its mel features with `librosa.feature.melspectrogram(..., fmin=150, fmax=sr//2,
htk=False, norm="slaney")` (`audio/spectrogram.py:64-77,88-101,117-130`), PCEN with `librosa.pcen`
(`:135-136`) and MFCCs with `librosa.feature.mfcc(norm="ortho")` (`:79-83`).  librosa is a third-party
dependency (pinned `==0.11.0`, `requirements.txt:1`) that is not installed here, so the small, data-independent
tables those calls derive -- the Slaney mel filterbank, the orthonormal DCT-II matrix and the PCEN smoothing
coefficient -- are restated from the published algorithms and handed to the CUDA feature kernels
(`csrc/bn_features.cu`) as plain float32 arrays.
"""

from __future__ import annotations

import numpy as np

_F_SP = 200.0 / 3
_MIN_LOG_HZ = 1000.0
_MIN_LOG_MEL = _MIN_LOG_HZ / _F_SP
_LOGSTEP = np.log(6.4) / 27.0


def hz_to_mel(freqs) -> np.ndarray:
    """Slaney (Auditory Toolbox) mel scale: linear below 1 kHz, logarithmic above (`htk=False`)."""
    f = np.asanyarray(freqs, dtype=np.float64)
    mels = f / _F_SP
    log_t = f >= _MIN_LOG_HZ
    return np.where(log_t, _MIN_LOG_MEL + np.log(np.maximum(f, 1e-300) / _MIN_LOG_HZ) / _LOGSTEP, mels)


def mel_to_hz(mels) -> np.ndarray:
    m = np.asanyarray(mels, dtype=np.float64)
    freqs = _F_SP * m
    log_t = m >= _MIN_LOG_MEL
    return np.where(log_t, _MIN_LOG_HZ * np.exp(_LOGSTEP * (m - _MIN_LOG_MEL)), freqs)


def mel_frequencies(n_mels: int, fmin: float, fmax: float) -> np.ndarray:
    return mel_to_hz(np.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels))


def mel_filterbank(sr: int, n_fft: int, n_mels: int, fmin: float = 150.0, fmax: float | None = None) -> np.ndarray:
    """float32 `[n_mels, 1 + n_fft//2]` triangular filters, Slaney area normalisation.

    Same construction as `librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax, htk=False, norm="slaney")`:
    ramps between consecutive mel centre frequencies, `max(0, min(lower, upper))`, scaled by
    `2 / (f[i+2] - f[i])`, stored as float32.
    """
    if fmax is None:
        fmax = float(sr) / 2
    weights = np.zeros((n_mels, 1 + n_fft // 2), dtype=np.float32)
    fftfreqs = np.fft.rfftfreq(n=n_fft, d=1.0 / sr)
    mel_f = mel_frequencies(n_mels + 2, fmin, fmax)
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2 : n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, np.newaxis]
    return weights


def filter_bands(basis: np.ndarray) -> np.ndarray:
    """int32 `[n_mels, 2]`: first bin and one-past-last bin with a non-zero weight (empty filters give 0, 0)."""
    bands = np.zeros((basis.shape[0], 2), dtype=np.int32)
    for m in range(basis.shape[0]):
        nz = np.nonzero(basis[m])[0]
        if nz.size:
            bands[m] = (nz[0], nz[-1] + 1)
    return bands


def dct_matrix(n_out: int, n_in: int) -> np.ndarray:
    """float32 `[n_out, n_in]` orthonormal DCT-II (`scipy.fftpack.dct(type=2, norm="ortho")` rows)."""
    n = np.arange(n_in, dtype=np.float64)
    k = np.arange(n_out, dtype=np.float64)[:, None]
    mat = 2.0 * np.cos(np.pi * (2.0 * n + 1.0) * k / (2.0 * n_in))
    mat[0] *= np.sqrt(1.0 / (4.0 * n_in))
    mat[1:] *= np.sqrt(1.0 / (2.0 * n_in))
    return mat.astype(np.float32)


def pcen_coefficient(sr: int, hop_length: int, time_constant: float = 0.400) -> float:
    """First-order IIR smoothing coefficient `b` of `librosa.pcen` for its default `time_constant`."""
    t_frames = time_constant * sr / float(hop_length)
    return float((np.sqrt(1 + 4 * t_frames**2) - 1) / (2 * t_frames**2))
