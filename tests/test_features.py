"""Precomputed-spectrogram frontends (BASELINE config 2): mel / log-mel / MFCC features with
none | pwl | pcen | db scaling.  CPU tests pin the oracle and the host tables; GPU tests compare the CUDA feature
kernels (through the C ABI of include/bn_features.h) with the oracle on the same seeded PCM.

Mirrors reference tests/test_spectrogram.py:11-35 (shapes, silence, dtype) on both implementations.
"""

import numpy as np
import pytest
import scipy.fftpack

from birdnet_stm32.audio import mel as melmod
from oracle import bn_features_oracle as fo

MODES = [("mel", "none"), ("mel", "pwl"), ("mel", "pcen"), ("mel", "db"), ("log_mel", "none"), ("mfcc", "none")]
# float32 CUDA pipeline vs the float64-FFT / float64-PCEN oracle, on min-max normalised features in [0, 1]
TOL = 1e-4


# ---------------------------------------------------------------------------------------------------------------
# CPU: host tables and oracle behaviour
# ---------------------------------------------------------------------------------------------------------------
def test_mel_filterbank_matches_oracle_and_slaney_properties():
    for sr, n_mels in ((24000, 64), (22050, 64), (24000, 32)):
        a = melmod.mel_filterbank(sr, 512, n_mels, 150.0, float(sr // 2))
        b = fo.mel_basis(sr, 512, n_mels, 150, sr // 2)
        assert a.dtype == np.float32 and a.shape == (n_mels, 257)
        assert np.array_equal(a, b)                      # two independent restatements of librosa.filters.mel
        assert (a >= 0).all() and (a.max(axis=1) > 0).all()
        # triangles: one contiguous run of non-zeros per filter, centres increasing
        bands = melmod.filter_bands(a)
        assert (np.diff(bands[:, 0]) >= 0).all() and (bands[:, 1] > bands[:, 0]).all()
        for m in range(n_mels):
            assert (a[m, bands[m, 0]:bands[m, 1]] > 0).all()
        # Slaney area normalisation: each continuous triangle integrates to 1 over Hz; sampled on the FFT grid the
        # wide (high) filters come close
        area = a[-8:].sum(axis=1) * (sr / 512)
        assert np.allclose(area, 1.0, atol=0.05)
    # the mel scale itself: linear below 1 kHz, log above, inverse consistent
    f = np.array([0.0, 150.0, 999.0, 1000.0, 4000.0, 12000.0])
    assert np.allclose(melmod.mel_to_hz(melmod.hz_to_mel(f)), f)
    assert np.isclose(melmod.hz_to_mel(1000.0), 15.0)


def test_dct_matrix_matches_scipy():
    x = np.random.default_rng(0).normal(size=(64, 7))
    want = scipy.fftpack.dct(x, axis=0, type=2, norm="ortho")[:20]
    got = melmod.dct_matrix(20, 64).astype(np.float64) @ x
    assert np.abs(got - want).max() < 1e-6


def test_pcen_coefficient_matches_filter_design():
    # b solves the librosa design equation  b^2 T^2 + b - 1 = 0  for T = time_constant * sr / hop frames
    for sr, hop in ((24000, 281), (22050, 258)):
        b = melmod.pcen_coefficient(sr, hop)
        T = 0.4 * sr / hop
        assert abs(b * b * T * T + b - 1.0) < 1e-12 and 0 < b < 1


@pytest.mark.parametrize("mode,mag", MODES)
def test_oracle_shapes_range_dtype(mode, mag):
    sr, T = 24000, 72000
    t = np.arange(T) / sr
    sine = (0.5 * np.sin(2 * np.pi * 1000.0 * t)).astype(np.float32)
    S = fo.get_spectrogram_from_audio(sine, sr, 512, 64, 256, mag, mode, 20)
    rows = 20 if mode == "mfcc" else 64
    assert S.shape == (rows, 256)
    assert np.isfinite(S).all() and S.min() >= 0.0 and S.max() <= 1.0 + 1e-6
    assert abs(float(S.max()) - 1.0) < 1e-5 and float(S.min()) == 0.0


def test_oracle_linear_and_silence():
    sr, T = 22050, 66150
    S = fo.get_spectrogram_from_audio(np.zeros(T, np.float32), sr, 512, 64, 256)
    assert S.shape == (64, 256) and np.max(np.abs(S)) < 1e-3           # reference test_silence_low_energy
    L = fo.get_spectrogram_from_audio(np.zeros(T, np.float32), sr, 512, -1, 256)
    assert L.shape == (257, 256)                                          # reference test_output_shape_linear
    x = np.random.default_rng(1).normal(size=T).astype(np.float32) * 0.1
    assert fo.get_spectrogram_from_audio(x, sr, 512, 32, 64).astype(np.float32).dtype == np.float32


def test_oracle_linear_matches_c_oracle_frontend(pcm_batch):
    """The numpy STFT of this oracle == the C oracle's hybrid frontend (float64 FFT stored as complex64)."""
    from oracle import bn_oracle

    pcm, peak = pcm_batch
    want = bn_oracle.frontend_hybrid(pcm[:3], peak[:3], 512, 66150 // 256, 256)[..., 0]
    got = fo.features_from_pcm16(pcm[:3], peak[:3], sample_rate=22050, n_fft=512, mel_bins=-1, spec_width=256)
    assert np.abs(got - want).max() < 2e-6


def test_feature_abi_symbols_declared_and_exported():
    import ctypes
    import os
    import re

    from birdnet_stm32 import _lib

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "bn_features.h")).read()
    declared = set(re.findall(r"BN_API\s+[\w\s\*]+?\b(bn_\w+)\s*\(", hdr))
    assert declared == {"bn_features_create", "bn_features_destroy", "bn_features_rows", "bn_features_pcm16"}
    L = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(L, name)
    assert declared <= set(_lib.EXPORTS)


def test_feature_extractor_argument_errors():
    from birdnet_stm32.audio.spectrogram import FeatureExtractor

    with pytest.raises(ValueError):
        FeatureExtractor(mode="nope")
    with pytest.raises(ValueError):
        FeatureExtractor(mag_scale="nope")
    with pytest.raises(ValueError):
        FeatureExtractor(mel_bins=-1)


# ---------------------------------------------------------------------------------------------------------------
# GPU: CUDA kernels vs oracle
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("sr,T", [(24000, 72000), (22050, 66150)])
@pytest.mark.parametrize("mode,mag", MODES)
def test_gpu_features_match_oracle(synth, sr, T, mode, mag):
    from birdnet_stm32.audio.spectrogram import FeatureExtractor

    pcm = synth.synth_pcm16(6, T, sr, seed=77, edge_cases=True)        # chirps + 1 kHz sine, impulse, square, silence
    peak = synth.file_peaks(pcm)
    fx = FeatureExtractor(sr, T, 512, 64, 256, mag, mode, 20)
    got = fx(pcm, peak)
    fx.close()
    want = fo.features_from_pcm16(pcm, peak, sample_rate=sr, n_fft=512, mel_bins=64, spec_width=256, mag_scale=mag,
                                  mode=mode, n_mfcc=20)
    assert got.shape == want.shape and got.dtype == np.float32
    assert np.isfinite(got).all()
    err = np.abs(got.astype(np.float64) - want.astype(np.float64)).reshape(got.shape[0], -1).max(axis=1)
    # the silent chunk is 0/1e-10 everywhere on both sides; every other chunk within TOL of the oracle
    tol = np.full(err.shape, TOL)
    if mag == "pcen":
        # PCEN divides every band by its own smoothed energy.  For the noiseless digital 1 kHz sine (chunk 2) most
        # bands hold nothing but FFT round-off (float32 on the GPU, float64 in the oracle, ~1e-7 of the peak), and
        # the gain control amplifies exactly that difference; any signal with a noise floor is unaffected.
        tol[2] = 1e-2
    assert (err <= tol).all(), (mode, mag, err)


@pytest.mark.gpu
def test_gpu_features_device_pointers_and_batches(synth):
    """Device-pointer entry, B > one sub-wave, 32 mel bands: same numbers as the host-pointer entry."""
    import torch

    from birdnet_stm32.audio.spectrogram import FeatureExtractor

    sr, T, B = 24000, 72000, 300
    base = synth.synth_pcm16(10, T, sr, seed=5)
    pcm = np.ascontiguousarray(base[np.arange(B) % 10])
    peak = synth.file_peaks(pcm)
    fx = FeatureExtractor(sr, T, 512, 32, 256, "pwl", "mel")
    host = fx(pcm, peak)
    dp = torch.from_numpy(pcm).cuda()
    dk = torch.from_numpy(peak).cuda()
    do = torch.empty((B, 32, 256), dtype=torch.float32, device="cuda")
    fx.run_device(dp.data_ptr(), dk.data_ptr(), B, do.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    fx.close()
    assert np.array_equal(do.cpu().numpy(), host)
    assert np.array_equal(host[:10], host[290:300])


def test_oracle_agrees_with_torchaudio_independent_implementation():
    """librosa is not installable here, but torch / torchaudio ship their own implementations of the same published
    conventions: Slaney mel filterbank with area normalisation (`melscale_fbanks(norm="slaney", mel_scale="slaney")`,
    which torchaudio tests against librosa.filters.mel) and a centred STFT with a periodic Hann window and constant
    padding (`torch.stft`).  The oracle's restatement of `librosa.filters.mel` / `librosa.stft`
    (`audio/spectrogram.py:106-130`) must agree with them: an independent cross-check of the librosa boundary."""
    torch = pytest.importorskip("torch")
    AF = pytest.importorskip("torchaudio.functional")
    from oracle import bn_features_oracle as FO

    from birdnet_stm32.audio.mel import mel_filterbank

    for sr in (22050, 24000, 16000):
        ta = AF.melscale_fbanks(257, 150.0, float(sr // 2), 64, sr, norm="slaney", mel_scale="slaney").numpy().T
        assert np.abs(FO.mel_basis(sr, 512, 64, 150.0, sr // 2) - ta).max() <= 2e-7
        assert np.abs(mel_filterbank(sr, 512, 64, 150.0, sr // 2) - ta).max() <= 2e-7       # the table the CUDA kernel receives
    rng = np.random.default_rng(0)
    for T in (66150, 72000):
        y = (rng.standard_normal(T) * 0.1).astype(np.float32)
        hop = T // 256
        m = FO.stft_mag(y, 512, hop)
        Y = torch.stft(torch.from_numpy(y), n_fft=512, hop_length=hop, win_length=512, window=torch.hann_window(512, periodic=True),
                       center=True, pad_mode="constant", return_complex=True).abs().numpy()
        assert m.shape[0] == 257 and m.shape[1] <= Y.shape[1]
        assert np.abs(m - Y[:, : m.shape[1]]).max() <= 1e-5 * np.abs(Y).max()


def test_float_audio_shim_stays_inside_the_frontend_tolerance():
    """`get_spectrogram_from_audio(float chunk)` hands the feature kernels int16 samples + a divisor
    (`audio_to_pcm16_peak`).  On the oracle: features of that representation vs features of the float chunk itself."""
    from oracle import bn_features_oracle as FO

    from birdnet_stm32.audio.spectrogram import audio_to_pcm16_peak

    rng = np.random.default_rng(12)
    T, sr = 66150, 22050
    t = np.arange(T) / sr
    for amp in (1.0, 0.03, 7.5):
        a = (amp * (0.6 * np.sin(2 * np.pi * (900 + 2500 * t) * t) + 0.1 * rng.standard_normal(T))).astype(np.float32)
        pcm, peak = audio_to_pcm16_peak(a)
        assert pcm.dtype == np.int16 and np.abs(pcm).max() == 32767
        rec = pcm.astype(np.float32) / np.float32(32768.0) / peak
        assert np.abs(rec - a).max() <= np.abs(a).max() / 65000.0
        # (dB scaling turns the 16-bit rounding of quiet bins into visible differences: 5e-4 there, 1e-4 elsewhere)
        for mode, mag, tol in (("mel", "none", 1e-4), ("mel", "pwl", 1e-4), ("mel", "pcen", 2e-4), ("mel", "db", 5e-4), ("log_mel", "none", 1e-4),
                               ("mfcc", "none", 1e-4)):
            want = FO.get_spectrogram_from_audio(a, sr, 512, 64, 256, mag, mode)
            got = FO.features_from_pcm16(pcm[None, :], np.array([peak], np.float32), sample_rate=sr, n_fft=512, mel_bins=64, spec_width=256,
                                         mag_scale=mag, mode=mode)[0]
            assert got.shape == want.shape and np.abs(got - want).max() <= tol, (amp, mode, mag, float(np.abs(got - want).max()))
    z, pk = audio_to_pcm16_peak(np.zeros(100, np.float32))
    assert not z.any() and pk == 0
