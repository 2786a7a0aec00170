"""Reference-signature shims on the GPU (run last: they were added after the round-1 GPU budget was spent, so a failure
here must not hide the parity tests above): `get_spectrogram_from_audio(float chunk)` and `fast_resample`."""

import numpy as np
import pytest


@pytest.mark.gpu
def test_get_spectrogram_from_audio_reference_signature():
    from oracle import bn_features_oracle as FO

    from birdnet_stm32.audio.spectrogram import get_spectrogram_from_audio

    rng = np.random.default_rng(21)
    T, sr = 66150, 22050
    t = np.arange(T) / sr
    a = (0.5 * np.sin(2 * np.pi * (700 + 3000 * t) * t) + 0.08 * rng.standard_normal(T)).astype(np.float32)
    for mode, mag, tol, rows in (("mel", "none", 2e-4, 64), ("mel", "pwl", 2e-4, 64), ("log_mel", "none", 2e-4, 64), ("mfcc", "none", 2e-4, 20),
                                 ("mel", "db", 6e-4, 64)):
        got = get_spectrogram_from_audio(a, sample_rate=sr, n_fft=512, mel_bins=64, spec_width=256, mag_scale=mag, mode=mode, n_mfcc=20)
        want = FO.get_spectrogram_from_audio(a, sr, 512, 64, 256, mag, mode, 20)
        assert got.shape == want.shape == (rows, 256) and got.dtype == np.float32
        assert got.min() >= 0.0 and got.max() <= 1.0 + 1e-6
        assert np.abs(got - want).max() <= tol, (mode, mag, float(np.abs(got - want).max()))
    with pytest.raises(ValueError):
        get_spectrogram_from_audio(a, sample_rate=sr, mel_bins=-1)


@pytest.mark.gpu
def test_fast_resample_reference_signature():
    from scipy.signal import resample_poly

    from birdnet_stm32.audio.io import fast_resample

    rng = np.random.default_rng(22)
    y = (rng.standard_normal(48000) * 0.3).astype(np.float32)
    got = fast_resample(y, 48000, 22050)
    want = resample_poly(y, 147, 320).astype(np.float32)
    assert got.shape == want.shape and got.dtype == np.float32 and np.abs(got - want).max() <= 2e-6
    assert fast_resample(y, 22050, 22050) is not None and np.array_equal(fast_resample(y, 22050, 22050), y)
