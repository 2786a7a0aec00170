"""Device ingest: decode -> channel mean -> resample_poly -> peak normalise -> chunks.

CPU tests: the oracle (`oracle/bn_ingest_oracle.py`, numpy + the real scipy.signal.resample_poly) reproduces the
outputs of the REAL reference `load_audio_window` / `split_audio_into_chunks` stored in
`tests/golden/ingest_reference.npz` bit for bit; the filter designed by the C++ host code equals scipy's firwin
tap for tap; length / chunk-count formulas; the RIFF parser.
GPU tests: `bn_ingest_*` against the golden reference outputs and against the oracle on larger seeded inputs, and
the float32-waveform entry of the engine (`bn_infer_wave_f32`) against the oracle frontend + graph.
Tolerance: resampled samples |d| <= 2e-6 of full scale (float32 products and sums in scipy's order; identical
results are expected and reported), normalised windows the same; everything without resampling is bit-exact.
"""

import ctypes as C
import json
import os
import struct

import numpy as np
import pytest

from conftest import CONFIG, GOLDEN, TFLITE

from birdnet_stm32 import _lib as L
from birdnet_stm32.audio import io as bio
from birdnet_stm32.audio.ingest import chunk_step, resample_ratio
from oracle import bn_ingest_oracle as O

KINDS = ["s16", "s24", "s32", "f32", "u8"]
TOL = 2e-6


def golden_cases():
    z = np.load(os.path.join(GOLDEN, "ingest_reference.npz"))
    for i in range(int(z["n_cases"])):
        sr_in, sr_out, ch, k = (int(v) for v in z[f"meta_{i}"])
        cd, ov, maxd = (float(v) for v in z[f"par_{i}"])
        yield dict(i=i, raw=z[f"raw_{i}"], sr_in=sr_in, sr_out=sr_out, ch=ch, kind=KINDS[k], cd=cd, ov=ov, maxd=maxd,
                   window=z[f"window_{i}"], chunks=z[f"chunks_{i}"])


def limit_frames(c):
    per = (3 if c["kind"] == "s24" else 1) * c["ch"]
    n = c["raw"].size // per
    n = int(min(n, c["maxd"] * c["sr_in"]))
    return c["raw"][: n * per], n


def write_wav(path, raw, kind, ch, sr, extensible=False):
    tag, bits = {"u8": (1, 8), "s16": (1, 16), "s24": (1, 24), "s32": (1, 32), "f32": (3, 32)}[kind]
    data = np.ascontiguousarray(raw).tobytes()
    bps = bits // 8
    with open(path, "wb") as fh:
        if extensible:
            fmt = struct.pack("<HHIIHHHHIH14s", 0xFFFE, ch, sr, sr * ch * bps, ch * bps, bits, 22, bits, 0, tag, b"\x00" * 14)
        else:
            fmt = struct.pack("<HHIIHH", tag, ch, sr, sr * ch * bps, ch * bps, bits)
        body = b"WAVE" + b"LIST" + struct.pack("<I", 4) + b"abcd" + b"fmt " + struct.pack("<I", len(fmt)) + fmt
        body += b"data" + struct.pack("<I", len(data)) + data
        fh.write(b"RIFF" + struct.pack("<I", len(body)) + body)


# ---------------------------------------------------------------------------------------------------------
# CPU: oracle pinned against the real reference, host arithmetic of the library
# ---------------------------------------------------------------------------------------------------------
def test_oracle_reproduces_reference_load_audio_window():
    n = 0
    for c in golden_cases():
        raw, _ = limit_frames(c)
        y, peak = O.load_window(raw, c["kind"], c["ch"], c["sr_in"], c["sr_out"])
        assert y.dtype == np.float32 and y.shape == c["window"].shape
        assert np.array_equal(y, c["window"]), f"case {c['i']}"
        assert np.array_equal(O.split_chunks(y, c["sr_out"], c["cd"], c["ov"]), c["chunks"]), f"case {c['i']}"
        n += 1
    assert n >= 9


def test_filter_taps_equal_scipy_firwin():
    lib = L.load()
    for sr_in, sr_out in ((48000, 22050), (44100, 22050), (44100, 24000), (32000, 24000), (16000, 22050), (48000, 24000),
                          (96000, 22050), (8000, 24000), (22050, 24000)):
        up, down = resample_ratio(sr_in, sr_out)
        n, pre = C.c_int(), C.c_int()
        assert lib.bn_ingest_filter(up, down, None, 0, C.byref(n), C.byref(pre)) == 0
        h = np.zeros(n.value, dtype=np.float32)
        assert lib.bn_ingest_filter(up, down, h.ctypes.data_as(C.c_void_p), n.value, C.byref(n), C.byref(pre)) == 0
        ref, pre_ref = O.resample_filter(up, down)
        assert pre.value == pre_ref
        assert h.size >= ref.size and np.all(h[ref.size:] == 0)
        assert np.array_equal(h[: ref.size], ref), f"{sr_in}->{sr_out}: {np.abs(h[:ref.size] - ref).max()}"
    # unreduced ratios are reduced first, 1:1 needs no filter
    assert lib.bn_ingest_filter(2, 2, None, 0, C.byref(n), C.byref(pre)) == 0 and n.value == 0


def test_length_and_chunk_count_formulas():
    lib = L.load()
    for c in golden_cases():
        _, n = limit_frames(c)
        assert lib.bn_ingest_out_len(n, c["sr_in"], c["sr_out"]) == c["window"].size
        T, step = chunk_step(c["sr_out"], c["cd"], c["ov"])
        assert T == c["chunks"].shape[1]
        assert lib.bn_ingest_num_chunks(c["window"].size, T, step) == c["chunks"].shape[0]
    # the reference's estimate_num_chunks, pinned in chunking_reference.json, through the C formula
    geo = json.load(open(os.path.join(GOLDEN, "chunking_reference.json")))
    for g in geo:
        T, step = chunk_step(g["sr"], g["cd"], g["overlap"])
        assert lib.bn_ingest_num_chunks(g["n"], T, step) == g["estimate"] == g["n_chunks"]
    assert lib.bn_ingest_out_len(0, 48000, 22050) == 0


def test_wav_parser_formats(tmp_path):
    rng = np.random.default_rng(3)
    for kind, ch, ext in (("s16", 1, False), ("s16", 2, True), ("s24", 3, False), ("s32", 1, False), ("f32", 2, True), ("u8", 1, False)):
        n = 1001
        if kind == "f32":
            raw = rng.standard_normal(n * ch).astype("<f4")
        elif kind == "s24" or kind == "u8":
            raw = rng.integers(0, 256, size=n * ch * (3 if kind == "s24" else 1), dtype=np.uint8)
        else:
            info = np.iinfo(np.int16 if kind == "s16" else np.int32)
            raw = rng.integers(info.min, info.max, size=n * ch).astype("<i2" if kind == "s16" else "<i4")
        p = str(tmp_path / f"{kind}_{ch}.wav")
        write_wav(p, raw, kind, ch, 44100, extensible=ext)
        got, k2, ch2, sr2 = bio.read_wav_frames(p)
        assert (k2, ch2, sr2) == (kind, ch, 44100) and np.array_equal(got, raw)
        got, *_ = bio.read_wav_frames(p, max_seconds=0.01)
        assert got.size == 441 * ch * (3 if kind == "s24" else 1)
    bad = tmp_path / "x.wav"
    bad.write_bytes(b"OggS" + b"\x00" * 64)
    with pytest.raises(bio.UnsupportedAudio):
        bio.read_wav_frames(str(bad))


# ---------------------------------------------------------------------------------------------------------
# GPU
# ---------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def ingest():
    from birdnet_stm32.audio.ingest import GpuIngest

    g = GpuIngest(0)
    yield g
    g.close()


@pytest.mark.gpu
def test_gpu_ingest_matches_reference_golden(ingest):
    worst = 0.0
    identical = 0
    for c in golden_cases():
        raw, n = limit_frames(c)
        y, peak = ingest.window(raw, c["kind"], c["ch"], c["sr_in"], c["sr_out"], normalize=True, return_peak=True)
        assert y.shape == c["window"].shape
        d = float(np.abs(y - c["window"]).max())
        worst = max(worst, d)
        identical += int(np.array_equal(y, c["window"]))
        assert d <= TOL, f"case {c['i']}: {d}"
        if c["sr_in"] == c["sr_out"]:
            assert np.array_equal(y, c["window"]), f"case {c['i']} (no resampling) must be bit-exact"
        T, step = chunk_step(c["sr_out"], c["cd"], c["ov"])
        ch = ingest.chunks(raw, c["kind"], c["ch"], c["sr_in"], c["sr_out"], T, step)
        assert ch.shape == c["chunks"].shape
        assert float(np.abs(ch - c["chunks"]).max()) <= TOL
        _, pk_ref = O.load_window(raw, c["kind"], c["ch"], c["sr_in"], c["sr_out"])
        assert abs(peak - pk_ref) <= TOL
    print(f"ingest vs reference golden: worst |d| = {worst:.3g}, bit-identical windows {identical}/9")
    assert ingest.launches > 0


@pytest.mark.gpu
def test_gpu_ingest_long_windows_vs_oracle(ingest):
    rng = np.random.default_rng(11)
    for sr_in, sr_out, ch, secs in ((48000, 22050, 2, 61.0), (44100, 24000, 1, 20.0), (96000, 22050, 1, 5.0), (8000, 24000, 2, 7.3),
                                    (22050, 24000, 1, 3.1), (22051, 22050, 1, 0.4), (44100, 22050, 2, 12.0)):
        # (22051 -> 22050 is coprime: up = 22050 phases do not fit a block, so the generic gather kernel runs)
        n = int(min(secs, 60) * sr_in)
        x = (rng.standard_normal((n, ch)) * 0.2).clip(-1, 1)
        x[:, 0] += 0.5 * np.sin(2 * np.pi * 1234.5 * np.arange(n) / sr_in)
        raw = np.round(32767 * x.clip(-1, 1)).astype(np.int16).reshape(-1)
        ref, pk = O.load_window(raw, "s16", ch, sr_in, sr_out)
        y = ingest.window(raw, "s16", ch, sr_in, sr_out)
        assert y.shape == ref.shape
        assert float(np.abs(y - ref).max()) <= TOL, f"{sr_in}->{sr_out}"
        T, step = chunk_step(sr_out, 3.0, 0.0)
        got = ingest.chunks(raw, "s16", ch, sr_in, sr_out, T, step)
        want = O.split_chunks(ref, sr_out, 3.0, 0.0)
        assert got.shape == want.shape and float(np.abs(got - want).max()) <= TOL
    # empty window
    assert ingest.window(np.zeros((0,), np.int16), "s16", 1, 48000, 22050).size == 0
    # silence: peak 0 -> not normalised
    z = ingest.window(np.zeros((4800,), np.int16), "s16", 1, 48000, 22050)
    assert z.shape == (2205,) and not z.any()


@pytest.mark.gpu
def test_gpu_wave_entry_matches_oracle_path(ingest):
    """48 kHz stereo file -> device ingest -> bn_infer_wave_f32 == oracle resample + frontend + int8 graph."""
    from birdnet_stm32.conversion.export_blob import export_blob
    from birdnet_stm32.evaluation.gpu_runner import GpuRunner
    from oracle import bn_oracle

    cfg = json.load(open(CONFIG))
    sr, T = int(cfg["sample_rate"]), int(cfg["sample_rate"] * cfg["chunk_duration"])
    blob = export_blob(TFLITE, cfg)
    rng = np.random.default_rng(5)
    n = int(10.4 * 48000)
    t = np.arange(n) / 48000
    x = np.stack([0.4 * np.sin(2 * np.pi * (800 + 900 * t) * t), 0.3 * np.sin(2 * np.pi * 3100 * t)], axis=1) + 0.05 * rng.standard_normal((n, 2))
    raw = np.round(32767 * x.clip(-1, 1)).astype(np.int16).reshape(-1)
    step = chunk_step(sr, cfg["chunk_duration"], 0.0)[1]
    chunks = ingest.chunks(raw, "s16", 2, 48000, sr, T, step)
    ref_wave, _ = O.load_window(raw, "s16", 2, 48000, sr)
    ref_chunks = O.split_chunks(ref_wave, sr, cfg["chunk_duration"], 0.0)
    assert chunks.shape == ref_chunks.shape == (4, T)
    runner = GpuRunner(blob, cfg)
    spec_ref = bn_oracle.frontend_hybrid_f32(ref_chunks, cfg["fft_length"], T // cfg["spec_width"], cfg["spec_width"])
    spec = runner.frontend_wave(chunks)
    rel = np.abs(spec - spec_ref).max()
    assert rel <= 1e-4, rel                                   # frontend tolerance (max-normalised spectrogram)
    model = bn_oracle.OracleModel(blob)
    want = model.predict(spec_ref)
    got = runner.predict_wave(chunks)
    assert np.abs(got - want).max() <= 1.0 / 256 + 1e-7        # dequantised scores within 1 LSB
    assert np.array_equal(got.argmax(axis=1), want.argmax(axis=1))
    offs = np.array([0, 1, 4], dtype=np.int32)
    pooled = runner.predict_pooled_wave(chunks, None, offs, pooling="lme", beta=10.0)
    for f in range(2):
        assert np.allclose(pooled[f], bn_oracle.pool_scores(got[offs[f]:offs[f + 1]], "lme", 10.0), atol=3e-6)
    # the float32 entry on PCM16-exact data equals the PCM16 entry bit for bit
    pcm = np.round(32767 * x[:, 0].clip(-1, 1)).astype(np.int16)[: 2 * T].reshape(2, T)
    a = runner.predict_pcm16(pcm, None)
    b = runner.predict_wave(pcm.astype(np.float32) / np.float32(32768.0), None)
    assert np.array_equal(a, b)
    runner.close()
