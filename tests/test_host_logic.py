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
    header = "".join(open(os.path.join(ROOT, "include", h)).read() for h in ("bn_engine.h", "bn_features.h", "bn_ingest.h", "bn_metrics.h"))
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
