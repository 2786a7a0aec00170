"""CPU ORACLE for the precomputed-spectrogram frontends (reference `audio/spectrogram.py:24-149`).

TEST INFRASTRUCTURE ONLY (imported from `tests/` and the `cpu_baseline` leg of `bench_frontend.py`; the
product package never imports it).

What it restates.  `get_spectrogram_from_audio` calls into librosa 0.11 (`requirements.txt:1`), which is not
installed here and not vendored in `/root/reference`; every librosa routine on the path is restated below from
its published algorithm, with the numpy dtype flow of the original kept (float64 window and FFT stored as
complex64, float32 magnitudes and mel projection, float64 PCEN because `scipy.signal.lfilter_zi` returns
float64).  The scipy pieces librosa delegates to (`get_window`, `lfilter`, `lfilter_zi`, `fftpack.dct`) ARE
available in this image and are called directly, so those steps run the reference's own third-party code.

PARITY UNPINNED at the librosa boundary: the reference holds no golden vectors for these features
(`tests/test_spectrogram.py:11-35` checks shapes, dtype and silence only; those behaviours are re-tested in
`tests/test_features.py`).
"""

from __future__ import annotations

import numpy as np
import scipy.fftpack
import scipy.signal


def normalize(S: np.ndarray) -> np.ndarray:
    """reference `audio/spectrogram.py:12-21`"""
    return (S - S.min()) / (S.max() - S.min() + 1e-10)


# ---- librosa.filters.mel(htk=False, norm="slaney") --------------------------------------------------------
def _hz_to_mel(f):
    f = np.atleast_1d(np.asarray(f, dtype=float))
    out = f / (200.0 / 3)
    hi = f >= 1000.0
    out[hi] = 15.0 + np.log(f[hi] / 1000.0) / (np.log(6.4) / 27.0)
    return out


def _mel_to_hz(m):
    m = np.atleast_1d(np.asarray(m, dtype=float))
    out = (200.0 / 3) * m
    hi = m >= 15.0
    out[hi] = 1000.0 * np.exp((np.log(6.4) / 27.0) * (m[hi] - 15.0))
    return out


def mel_basis(sr: int, n_fft: int, n_mels: int, fmin: float, fmax: float) -> np.ndarray:
    lo, hi = _hz_to_mel(fmin)[0], _hz_to_mel(fmax)[0]
    centres = _mel_to_hz(np.linspace(lo, hi, n_mels + 2))
    freqs = np.linspace(0.0, sr / 2.0, 1 + n_fft // 2)
    w = np.zeros((n_mels, freqs.size), dtype=np.float32)
    for i in range(n_mels):
        left, mid, right = centres[i], centres[i + 1], centres[i + 2]
        up = (freqs - left) / (mid - left)
        down = (right - freqs) / (right - mid)
        w[i] = np.maximum(0, np.minimum(up, down))
    w *= (2.0 / (centres[2:] - centres[:-2]))[:, None]
    return w


# ---- librosa.stft(center=True, pad_mode="constant", window="hann") -> |.| ------------------------------------
def stft_mag(y: np.ndarray, n_fft: int, hop: int) -> np.ndarray:
    y = np.asarray(y, dtype=np.float32)
    yp = np.pad(y, n_fft // 2, mode="constant")
    n_frames = 1 + (yp.size - n_fft) // hop
    idx = np.arange(n_fft)[:, None] + hop * np.arange(n_frames)[None, :]
    win = scipy.signal.get_window("hann", n_fft, fftbins=True)
    D = np.fft.rfft(win[:, None] * yp[idx], axis=0).astype(np.complex64)
    return np.abs(D)


def _power_to_db(S, ref_value, amin, top_db=80.0):
    log_spec = 10.0 * np.log10(np.maximum(amin, S))
    log_spec -= 10.0 * np.log10(np.maximum(amin, ref_value))
    return np.maximum(log_spec, log_spec.max() - top_db)


def _pcen(S, sr, hop_length, gain=0.98, bias=2.0, power=0.5, time_constant=0.400, eps=1e-6):
    t_frames = time_constant * sr / float(hop_length)
    b = (np.sqrt(1 + 4 * t_frames**2) - 1) / (2 * t_frames**2)
    zi = np.empty((1, 1))
    zi[:] = scipy.signal.lfilter_zi([b], [1, b - 1])[:]
    S_smooth, _ = scipy.signal.lfilter([b], [1, b - 1], S, zi=zi, axis=1)
    smooth = np.exp(-gain * (np.log(eps) + np.log1p(S_smooth / eps)))
    return (bias**power) * np.expm1(power * np.log1p(S * smooth / bias))


def get_spectrogram_from_audio(audio, sample_rate=24000, n_fft=512, mel_bins=64, spec_width=256, mag_scale="none",
                               mode="mel", n_mfcc=20) -> np.ndarray:
    """Restatement of reference `audio/spectrogram.py:24-149` (same arguments, same return)."""
    audio = np.asarray(audio, dtype=np.float32)
    hop = (len(audio) // spec_width) if spec_width > 0 else n_fft // 2
    mag = stft_mag(audio, n_fft, hop)
    fb = None
    if not (mel_bins <= 0 or mode == "linear") or mode in ("mfcc", "log_mel"):
        fb = mel_basis(sample_rate, n_fft, mel_bins, 150, sample_rate // 2)
    if mode == "mfcc":
        S_mel = fb @ (mag**2.0)
        S_log = _power_to_db(S_mel, np.abs(S_mel.max()), 1e-10)
        S = scipy.fftpack.dct(S_log, axis=-2, type=2, norm="ortho")[:n_mfcc, :]
        return normalize(S[:, :spec_width])
    if mode == "log_mel":
        S = (fb @ mag)[:, :spec_width]
        return normalize(np.log1p(S))
    S = mag if (mel_bins <= 0 or mode == "linear") else fb @ mag
    S = S[:, :spec_width]
    if mag_scale == "pcen":
        S = _pcen(S * (2.0**31), sample_rate, hop)
    elif mag_scale == "pwl":
        Smin, Smax = S.min(), S.max()
        Snorm = (S - Smin) / (Smax - Smin + 1e-10)
        relu = lambda z: np.maximum(z, 0.0)  # noqa: E731
        S = 0.40 * Snorm + 0.25 * relu(Snorm - 0.10) + 0.15 * relu(Snorm - 0.35) + 0.08 * relu(Snorm - 0.65)
    elif mag_scale == "db":
        magnitude = np.abs(S)
        ref_value = magnitude.max()
        S = _power_to_db(np.square(magnitude), ref_value**2, 1e-5**2)
    return normalize(S)


def features_from_pcm16(pcm: np.ndarray, peak, **kw) -> np.ndarray:
    """PCM16 `[B, T]` (+ file peaks) -> float32 `[B, rows, spec_width]`; decode as `audio/io.py:114-126`."""
    pcm = np.asarray(pcm, dtype=np.int16)
    out = []
    for b in range(pcm.shape[0]):
        y = pcm[b].astype(np.float32) / np.float32(32768.0)
        if peak is not None and float(peak[b]) > 0:
            y = y / np.float32(peak[b])
        out.append(get_spectrogram_from_audio(y, **kw).astype(np.float32))
    return np.stack(out)
