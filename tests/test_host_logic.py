"""Host-side mirror of the reference interface: chunk geometry, config, registry, exporter, ABI."""

import ctypes
import json
import os
import re

import numpy as np
import pytest

from conftest import GOLDEN, ROOT


def test_chunk_geometry_matches_reference():
    """split_audio_into_chunks / estimate_num_chunks against outputs of the reference's own functions."""
    from birdnet_stm32.audio import io

    cases = json.load(open(os.path.join(GOLDEN, "chunking_reference.json")))
    assert len(cases) > 200
    for c in cases:
        y = (np.arange(c["n"]) % 251).astype(np.float32)
        chunks = io.split_audio_into_chunks(y, c["sr"], c["cd"], c["overlap"])
        assert chunks.shape == (c["n_chunks"], c["shape1"]), c
        assert io.estimate_num_chunks(c["n"], c["sr"], c["cd"], c["overlap"]) == c["estimate"], c
        assert [int(ch[0]) for ch in chunks] == c["firsts"] and [int(ch[-1]) for ch in chunks] == c["lasts"], c
        # int16 input takes the same geometry
        if c["n"]:
            ch16 = io.split_audio_into_chunks(y.astype(np.int16), c["sr"], c["cd"], c["overlap"])
            assert ch16.dtype == np.int16 and np.array_equal(ch16, chunks.astype(np.int16))


def test_short_file_is_padded_not_truncated(tmp_path):
    """reference tests/test_audio_io.py:36-48"""
    from birdnet_stm32.audio import io

    sr = 16000
    audio = np.linspace(-1.0, 1.0, sr, dtype=np.float32)
    path = str(tmp_path / "short.wav")
    io.save_wav(audio, path, sr)
    chunks = io.load_audio_file(path, sample_rate=sr, chunk_duration=3.0)
    assert chunks.shape == (1, sr * 3)
    np.testing.assert_allclose(chunks[0, :sr], audio / np.max(np.abs(audio)), atol=1e-4)
    np.testing.assert_allclose(chunks[0, sr:], 0.0, atol=1e-7)
    pcm, peak = io.load_pcm16_chunks(path, sr, 3.0)
    assert pcm.dtype == np.int16 and pcm.shape == (1, sr * 3) and 0.99 < float(peak) <= 1.0
    assert io.load_audio_file(str(tmp_path / "missing.wav"), sample_rate=sr) == []
    with pytest.raises(io.UnsupportedAudio):
        io.load_pcm16_window(path, 22050)


def test_model_config_matches_reference_defaults(cfg):
    from birdnet_stm32.training.config import ModelConfig

    ref = json.load(open(os.path.join(GOLDEN, "config_reference.json")))
    assert ModelConfig().to_dict() == ref["default"]
    shipped = ModelConfig.from_dict(cfg)
    assert shipped.to_dict() == ref["shipped"]
    assert shipped.use_se is True and shipped.sample_rate == 22050      # legacy JSON: defaults filled in
    assert ModelConfig.from_dict({"sample_rate": 16000, "unknown_key": 1}).sample_rate == 16000
    with pytest.raises(ValueError, match="sample_rate"):
        ModelConfig(sample_rate=-1)
    with pytest.raises(ValueError, match="audio_frontend"):
        ModelConfig(audio_frontend="nope")
    with pytest.raises(ValueError, match="class_names"):
        ModelConfig(num_classes=2, class_names=["a"])


def test_config_round_trip(tmp_path):
    from birdnet_stm32.training.config import ModelConfig

    c = ModelConfig(num_classes=2, class_names=["x", "y"], sample_rate=16000)
    c.save(tmp_path / "sub" / "c.json")
    assert ModelConfig.load(tmp_path / "sub" / "c.json") == c


def test_frontend_registry_and_names():
    from birdnet_stm32.models import registry
    from birdnet_stm32.models.frontend import normalize_frontend_name

    assert registry.list_frontends() == ["hybrid", "librosa", "log_mel", "mfcc", "raw"]
    assert registry.get_frontend_info("hybrid").mode == "hybrid" and not registry.is_precomputed("hybrid")
    assert registry.is_precomputed("librosa") and registry.is_n6_compatible("raw")
    assert registry.has_gpu_frontend("hybrid") and registry.has_gpu_frontend("mfcc") and not registry.has_gpu_frontend("raw")
    with pytest.raises(KeyError, match="not registered"):
        registry.get_frontend_info("nope")
    with pytest.raises(ValueError, match="already registered"):
        registry.register_frontend(registry.FrontendInfo("hybrid", "hybrid", False, True))
    assert normalize_frontend_name("hybrid") == "hybrid"
    with pytest.warns(DeprecationWarning):
        assert normalize_frontend_name("tf") == "raw"
    with pytest.raises(ValueError, match="Invalid audio frontend"):
        normalize_frontend_name("stft")


def test_blob_layout_and_exporter_errors(graph, blob, cfg):
    import struct

    from birdnet_stm32.conversion import export_blob as eb
    from birdnet_stm32.conversion.tflite_reader import OpInfo, read_tflite

    magic, version, hb, n_t, n_o = struct.unpack_from("<8sIIII", blob, 0)
    assert magic == b"BNB200\0\0" and version == 2 and hb == 128
    fe, mag, sr, T, n_fft, hop, W, mels, C = struct.unpack_from("<9I", blob, 56)
    assert (fe, mag, sr, T, n_fft, hop, W, mels, C) == (1, 1, 22050, 66150, 512, 258, 256, 64, 100)
    assert n_o == 53            # 56 TFLite ops minus SHAPE / STRIDED_SLICE(shape) / PACK
    assert struct.unpack_from("<Q", blob, 48)[0] == len(blob)
    with pytest.raises(ValueError, match="TFL3"):
        read_tflite(b"\0" * 64)
    bad = type(graph)(tensors=graph.tensors, ops=list(graph.ops) + [OpInfo(99, "LOG", 1, [130], [130])], inputs=graph.inputs, outputs=graph.outputs)
    with pytest.raises(ValueError, match="unsupported operator LOG"):
        eb.export_blob(bad, cfg)
    # SAME padding asymmetry of stride-2 3x3 on even sizes: 0 before, 1 after
    assert eb.same_padding(64, 3, 2) == (32, 0) and eb.same_padding(64, 3, 1) == (64, 1) and eb.same_padding(256, 3, 2) == (128, 0)
    assert eb.activation_range("RELU6", 6 / 255, -128) == (-128, 127)
    assert eb.activation_range("RELU", 0.5, 3) == (3, 127)


def test_c_abi_library_loads_and_exports_every_declared_symbol():
    """No compute calls here (no GPU): the .so loads, exports what include/*.h declare, and
    creating an engine without a CUDA device fails loudly instead of falling back to the CPU."""
    from birdnet_stm32 import _lib as L

    lib = L.load()
    header = "".join(open(os.path.join(ROOT, "include", h)).read() for h in ("bn_engine.h", "bn_features.h", "bn_ingest.h", "bn_metrics.h", "bn_reader.h"))
    declared = set(re.findall(r"BN_API\s+[\w\s\*]+?\b(bn_\w+)\s*\(", header))
    assert declared == set(L.EXPORTS), declared ^ set(L.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.bn_version()
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if not has_gpu:
        from birdnet_stm32.evaluation.gpu_runner import GpuRunner

        from birdnet_stm32.conversion.export_blob import export_blob
        from conftest import TFLITE

        with pytest.raises(L.EngineError, match="no CPU fallback"):
            GpuRunner(export_blob(TFLITE, {}))
        from birdnet_stm32.audio.ingest import GpuIngest

        with pytest.raises(L.EngineError, match="no CPU fallback"):
            GpuIngest(0)


def test_oracle_has_not_drifted(oracle_model, pcm_batch):
    """Self-generated golden (oracle outputs at commit time) -- guards the checker itself."""
    from oracle import bn_oracle

    g = np.load(os.path.join(GOLDEN, "oracle_selfcheck.npz"))
    pcm, peak = pcm_batch
    spec = bn_oracle.frontend_hybrid(pcm, peak, 512, 66150 // 256, 256)
    np.testing.assert_allclose(spec[:, :, 100, 0], g["spec_col"], rtol=0, atol=1e-6)
    scores, t96 = oracle_model.run(spec, tap_id=96)
    np.testing.assert_array_equal(t96, g["t96"])
    np.testing.assert_array_equal(scores, g["scores"])


def test_prefetch_ordered_keeps_order_and_bounds_lookahead():
    import threading
    import time

    from birdnet_stm32.audio.io import prefetch_ordered

    started = []
    lock = threading.Lock()

    def work(i):
        with lock:
            started.append(i)
        time.sleep(0.002 * ((i * 7) % 5))
        return i * i

    got = []
    for k, v in enumerate(prefetch_ordered(work, range(40), workers=4, depth=6)):
        got.append(v)
        with lock:
            assert max(started) <= k + 6, "ran further ahead than `depth`"
    assert got == [i * i for i in range(40)]
    assert list(prefetch_ordered(work, range(3), workers=1)) == [0, 1, 4]
    assert list(prefetch_ordered(work, [], workers=4)) == []

    def boom(i):
        if i == 2:
            raise ValueError("x")
        return i

    with pytest.raises(ValueError):
        list(prefetch_ordered(boom, range(5), workers=3))


def test_native_batch_reader_matches_python_loader(tmp_path):
    """bn_read_pcm16_batch (C++ threads) against the Python loader that is itself pinned on the reference's chunk geometry:
    identical chunks, peaks and statuses for exact multiples, tails, short files, overlap, foreign formats, junk."""
    from test_ingest import write_wav

    from birdnet_stm32.audio import io, reader
    from birdnet_stm32.audio.ingest import chunk_step

    sr = 8000
    rng = np.random.default_rng(5)
    paths, kinds = [], []
    for i, secs in enumerate((3.0, 6.0, 7.37, 0.4, 9.0, 2.999875, 12.5, 61.0)):
        p = str(tmp_path / f"a{i}.wav")
        io.save_wav((rng.standard_normal(int(sr * secs)) * 9000).clip(-32768, 32767).astype(np.int16), p, sr)
        paths.append(p); kinds.append("ok")
    p = str(tmp_path / "stereo.wav"); write_wav(p, rng.integers(-3000, 3000, 2 * 4000).astype("<i2"), "s16", 2, sr); paths.insert(2, p); kinds.insert(2, "ingest")
    p = str(tmp_path / "rate.wav"); write_wav(p, rng.integers(-3000, 3000, 9000).astype("<i2"), "s16", 1, 16000, extensible=True); paths.insert(5, p); kinds.insert(5, "ingest")
    p = str(tmp_path / "float.wav"); write_wav(p, rng.standard_normal(5000).astype("<f4"), "f32", 1, sr); paths.append(p); kinds.append("ingest")
    p = str(tmp_path / "junk.wav"); open(p, "wb").write(b"RIFFxxxxWAVEjunk"); paths.append(p); kinds.append("bad")
    p = str(tmp_path / "empty.wav"); io.save_wav(np.zeros((0,), np.int16), p, sr); paths.append(p); kinds.append("bad")
    paths.append(str(tmp_path / "missing.wav")); kinds.append("bad")
    pr = reader.probe(paths[0], 60)
    assert (pr.status, pr.sample_rate, pr.channels, reader.FMT_NAMES[pr.fmt], pr.n_frames) == (reader.RD_NEEDS_INGEST, sr, 1, "s16", 3 * sr)
    for cd, ov in ((3.0, 0.0), (3.0, 1.0), (1.5, 0.4)):
        T, step = chunk_step(sr, cd, ov)
        for threads in (1, 4):
            out = np.full((400, T), 77, dtype=np.int16)
            n_files, used, info = reader.read_pcm16_batch(paths, sr, T, step, out, max_seconds=60, threads=threads)
            assert n_files == len(paths)
            row = 0
            for path, kind, fi in zip(paths, kinds, info):
                if kind == "ok":
                    want, peak = io.load_pcm16_chunks(path, sr, cd, ov, max_duration=60)
                    assert fi.status == reader.RD_OK and fi.n_chunks == want.shape[0], path
                    assert np.array_equal(out[row:row + fi.n_chunks], want), (path, cd, ov)
                    assert np.float32(fi.peak) == peak
                    row += fi.n_chunks
                elif kind == "ingest":
                    assert fi.status == reader.RD_NEEDS_INGEST and fi.n_chunks == 0
                else:
                    assert fi.status == reader.RD_UNREADABLE and fi.n_chunks == 0
            assert row == used and np.all(out[used:] == 77)
    # a full buffer stops at a file boundary
    T, step = chunk_step(sr, 3.0, 0.0)
    out = np.zeros((5, T), dtype=np.int16)
    n_files, used, info = reader.read_pcm16_batch(paths, sr, T, step, out, threads=3)
    assert n_files == 3 and used == 3                    # 1 + 2 chunks, the stereo file takes none, the 7.37 s file (3) does not fit
    assert reader.read_pcm16_batch(paths[3:4], sr, T, step, np.zeros((2, T), np.int16))[0] == 0   # first file alone does not fit
    with pytest.raises(ValueError):
        reader.read_pcm16_batch(paths, sr, T, step, np.zeros((4, T + 1), np.int16))


def test_native_raw_batch_reader_matches_python_wav_parser(tmp_path):
    from test_ingest import write_wav

    from birdnet_stm32.audio import io, reader

    rng = np.random.default_rng(8)
    paths = []
    for i, (kind, ch, sr, n) in enumerate((("s16", 2, 48000, 30001), ("f32", 1, 44100, 12345), ("s24", 3, 32000, 7777), ("u8", 1, 16000, 5000),
                                           ("s32", 2, 96000, 4001), ("s16", 1, 22050, 700000))):
        if kind == "f32":
            raw = rng.standard_normal(n * ch).astype("<f4")
        elif kind in ("s24", "u8"):
            raw = rng.integers(0, 256, size=n * ch * (3 if kind == "s24" else 1), dtype=np.uint8)
        else:
            info = np.iinfo(np.int16 if kind == "s16" else np.int32)
            raw = rng.integers(info.min, info.max, size=n * ch).astype("<i2" if kind == "s16" else "<i4")
        p = str(tmp_path / f"r{i}.wav")
        write_wav(p, raw, kind, ch, sr, extensible=(i % 2 == 1))
        paths.append(p)
    paths.insert(2, str(tmp_path / "missing.wav"))
    buf = np.zeros(8 << 20, dtype=np.uint8)
    for threads in (1, 4):
        for max_s in (60.0, 0.05):
            n_files, items = reader.read_raw_batch(paths, buf, max_seconds=max_s, threads=threads)
            assert n_files == len(paths) and items[2] is None
            for path, item in zip(paths, items):
                if item is None:
                    continue
                want = io.read_wav_frames(path, max_s)
                assert item[1:] == want[1:], path
                assert item[0].dtype == want[0].dtype and np.array_equal(item[0], want[0]), path
    # a small buffer stops at a file boundary
    n_files, items = reader.read_raw_batch(paths, np.zeros(200000, dtype=np.uint8), threads=2)
    assert n_files == 3 and items[2] is None             # 120,004 + 49,380 bytes fit, the missing file takes none, the 24-bit file does not fit
