"""evaluate(): reference behaviour with a FakeRunner (tests/test_metrics.py:11-101 in the reference),
the file-sharded wrapper under gloo (world_size 2), and -- on the GPU box -- the device path against the
oracle path (top-1 and cmAP to 3 decimals)."""

import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import PKG, ROOT


class FakeRunner:
    """Fixed predictions, like the reference tests' FakeRunner."""

    def __init__(self, scores):
        self.scores = np.asarray(scores, dtype=np.float32)
        self._idx = 0

    def predict(self, x_batch):
        b = x_batch.shape[0]
        out = np.tile(self.scores[self._idx % len(self.scores)], (b, 1))
        self._idx += b
        return out.astype(np.float32)


def make_dataset(root, classes, n_per_class=3, sr=22050, seconds=(3.0, 7.4, 1.2)):
    from birdnet_stm32.audio import io

    files = []
    for ci, cls in enumerate(classes):
        os.makedirs(os.path.join(root, cls), exist_ok=True)
        for i in range(n_per_class):
            n = int(sr * seconds[i % len(seconds)])
            rng = np.random.default_rng(100 * ci + i)
            t = np.arange(n) / sr
            audio = 0.4 * np.sin(2 * np.pi * (500 + 700 * ci) * t) + 0.1 * rng.standard_normal(n)
            path = os.path.join(root, cls, f"sample_{i}.wav")
            io.save_wav(audio.astype(np.float32), path, sr)
            files.append(path)
    return files


RAW_CFG = {"sample_rate": 22050, "chunk_duration": 3, "num_mels": 64, "spec_width": 256, "fft_length": 512,
           "audio_frontend": "raw", "mag_scale": "none"}


def test_fake_runner_metrics_keys_and_shapes(tmp_path):
    from birdnet_stm32.evaluation.metrics import evaluate

    classes = ["bird_a", "bird_b"]
    files = make_dataset(str(tmp_path), classes)
    os.makedirs(tmp_path / "unknown_label")
    metrics, per_file, y_true, y_scores = evaluate(FakeRunner([[0.95, 0.05], [0.05, 0.95]]), files, classes, RAW_CFG,
                                                   pooling="avg", batch_size=64, measure_latency=True, profile_memory=True)
    for key in ("roc-auc", "f1", "precision", "recall", "cmAP", "mAP", "ap_per_class", "latency_mean_ms", "latency_p99_ms",
                "total_chunks", "peak_rss_mb"):
        assert key in metrics, key
    assert len(per_file) == len(files) and y_true.shape == (len(files), 2) and y_scores.shape == (len(files), 2)
    assert y_true.dtype == np.float32 and y_scores.dtype == np.float32
    assert metrics["total_chunks"] == 2 * (1 + 3 + 1)        # 3 s -> 1 chunk, 7.4 s -> 2 + tail, 1.2 s -> 1 padded
    assert metrics["f1"] > 0.0


def test_no_valid_files_raises():
    from birdnet_stm32.evaluation.metrics import evaluate

    with pytest.raises(RuntimeError, match="No valid test samples"):
        evaluate(FakeRunner([[0.5, 0.5]]), [], ["a", "b"], RAW_CFG, pooling="avg")


def test_hybrid_without_gpu_frontend_is_an_error_not_a_cpu_fallback(tmp_path):
    from birdnet_stm32.evaluation.metrics import evaluate

    files = make_dataset(str(tmp_path), ["a"], n_per_class=1)
    cfg = dict(RAW_CFG, audio_frontend="hybrid")
    with pytest.raises(RuntimeError, match="no CPU spectrogram"):
        evaluate(FakeRunner([[0.5]]), files, ["a"], cfg)


def test_shard_files_partitions_whole_files():
    from birdnet_stm32.evaluation.sharded import shard_files

    files = [f"f{i}" for i in range(11)]
    weights = [20, 1, 1, 7, 3, 3, 12, 1, 5, 9, 2]
    for world in (1, 2, 4, 8):
        parts = [shard_files(files, world, r, weights) for r in range(world)]
        assert sorted(sum(parts, [])) == sorted(files)
        loads = [sum(weights[files.index(f)] for f in p) for p in parts]
        assert max(loads) - min(loads) <= max(weights)
    assert shard_files(files, 2, 1) == files[1::2]


_WORKER = r"""
import os, sys, json
sys.path.insert(0, {root!r}); sys.path.insert(0, {pkg!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import numpy as np, torch.distributed as dist
from test_evaluate import FakeRunner, RAW_CFG
from birdnet_stm32.evaluation.sharded import evaluate_sharded
dist.init_process_group("gloo")
files = sorted(json.load(open({files!r})))
classes = ["bird_a", "bird_b", "bird_c"]
class ByLabel(FakeRunner):
    pass
m, per_file, yt, ys = evaluate_sharded(FakeRunner([[0.9, 0.2, 0.1]]), files, classes, RAW_CFG, pooling="lme")
if dist.get_rank() == 0:
    json.dump({{"n": int(yt.shape[0]), "labels": yt.argmax(1).tolist(), "cmAP": m["cmAP"], "local": len(per_file)}}, open({out!r}, "w"))
dist.destroy_process_group()
"""


def test_sharded_evaluate_gloo_world2(tmp_path):
    import json

    classes = ["bird_a", "bird_b", "bird_c"]
    files = make_dataset(str(tmp_path / "data"), classes, n_per_class=3)
    flist, out, script = str(tmp_path / "files.json"), str(tmp_path / "out.json"), str(tmp_path / "worker.py")
    json.dump(files, open(flist, "w"))
    open(script, "w").write(_WORKER.format(root=ROOT, pkg=PKG, files=flist, out=out))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29531", script], capture_output=True, text=True, env=env, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    got = json.load(open(out))
    assert got["n"] == len(files) and got["local"] < len(files)
    assert sorted(got["labels"]) == sorted([0] * 3 + [1] * 3 + [2] * 3)


@pytest.mark.gpu
def test_gpu_evaluate_matches_oracle_path(tmp_path, blob, cfg, oracle_model):
    """Device path (PCM16 batched across files, pooled on the GPU) vs the oracle path on the same files:
    top-1 per file identical, cmAP / ROC-AUC equal to 3 decimals (BASELINE.md parity bar)."""
    from birdnet_stm32.audio import io
    from birdnet_stm32.evaluation.gpu_runner import GpuRunner
    from birdnet_stm32.evaluation.metrics import _metrics_from_scores, evaluate
    from oracle import bn_oracle

    classes = cfg["class_names"]
    files = make_dataset(str(tmp_path), classes[:6], n_per_class=3, sr=22050)
    runner = GpuRunner(blob, cfg)
    try:
        metrics, per_file, y_true, y_scores = evaluate(runner, files, classes, cfg, pooling="lme", device_batch_chunks=7)
        ref_scores = []
        for path in files:
            pcm, peak = io.load_pcm16_chunks(path, 22050, 3.0)
            spec = bn_oracle.frontend_hybrid(pcm, np.full(len(pcm), peak, np.float32), 512, 66150 // 256, 256)
            ref_scores.append(bn_oracle.pool_scores(oracle_model.predict(spec), "lme", 10.0))
        ref_scores = np.asarray(ref_scores, dtype=np.float32)
        assert y_scores.shape == ref_scores.shape == (len(files), 100)
        np.testing.assert_array_equal(y_scores.argmax(1), ref_scores.argmax(1))
        assert np.abs(y_scores - ref_scores).max() <= 1 / 256 + 1e-5
        ref_m = _metrics_from_scores(y_true, ref_scores)
        assert round(metrics["cmAP"], 3) == round(ref_m["cmAP"], 3)
        assert round(metrics["roc-auc"], 3) == round(ref_m["roc-auc"], 3)
        # protocol path (per-file predict() calls on GPU spectrograms + host pooling) agrees with the device path
        class ProtocolOnly:
            predict = staticmethod(runner.predict)
            frontend = staticmethod(runner.frontend)

        _, _, _, ys2 = evaluate(ProtocolOnly(), files, classes, cfg, pooling="lme", batch_size=2)
        np.testing.assert_allclose(ys2, y_scores, rtol=0, atol=3e-6)
    finally:
        runner.close()


@pytest.mark.gpu
def test_gpu_evaluate_mixed_rates_and_formats(tmp_path, blob, cfg, oracle_model):
    """Files the reference would resample / mix on the host (48 kHz stereo, 44.1 kHz float32, 32 kHz 24-bit) next to
    native 22.05 kHz PCM16 files: device ingest + float32 waveform entry vs the oracle (real scipy resample_poly +
    numpy + C frontend + int8 graph).  Same bars as the PCM16 path: scores within 1 LSB, top-1 identical."""
    from test_ingest import write_wav

    from birdnet_stm32.audio import io
    from birdnet_stm32.evaluation.gpu_runner import GpuRunner
    from birdnet_stm32.evaluation.metrics import evaluate
    from oracle import bn_ingest_oracle as O
    from oracle import bn_oracle

    classes = cfg["class_names"]
    rng = np.random.default_rng(77)
    specs = [(48000, 2, "s16", 7.0), (22050, 1, "s16", 4.1), (44100, 1, "f32", 3.0), (32000, 3, "s24", 9.5), (22050, 2, "s16", 2.0),
             (48000, 1, "s16", 61.5)]
    files, raws = [], []
    for i, (sr0, ch, kind, secs) in enumerate(specs):
        n = int(sr0 * secs)
        t = np.arange(n) / sr0
        x = np.stack([0.35 * np.sin(2 * np.pi * (600 + 450 * i + 200 * c) * t) for c in range(ch)], axis=1) + 0.08 * rng.standard_normal((n, ch))
        x = x.clip(-1, 1)
        if kind == "s16":
            raw = np.round(32767 * x).astype("<i2").reshape(-1)
        elif kind == "f32":
            raw = x.astype("<f4").reshape(-1)
        else:
            v = np.round(8388607 * x).astype(np.int32).reshape(-1)
            raw = np.stack([v & 0xFF, (v >> 8) & 0xFF, (v >> 16) & 0xFF], axis=1).astype(np.uint8).reshape(-1)
        d = tmp_path / classes[i]
        d.mkdir()
        path = str(d / "f.wav")
        write_wav(path, raw, kind, ch, sr0)
        files.append(path)
        raws.append((raw, kind, ch, sr0))
    runner = GpuRunner(blob, cfg)
    try:
        metrics, per_file, y_true, y_scores = evaluate(runner, files, classes, cfg, pooling="lme", device_batch_chunks=9)
        assert y_scores.shape == (len(files), 100) and "skipped_files" not in metrics
        assert [f["file"] for f in per_file] == files
        ref = []
        for raw, kind, ch, sr0 in raws:
            per = (3 if kind == "s24" else 1) * ch
            nfr = int(min(raw.size // per, 60 * sr0))
            wave, _ = O.load_window(raw[: nfr * per], kind, ch, sr0, 22050)
            chunks = O.split_chunks(wave, 22050, 3.0, 0.0)
            spec = bn_oracle.frontend_hybrid_f32(chunks, 512, 66150 // 256, 256)
            ref.append(bn_oracle.pool_scores(oracle_model.predict(spec), "lme", 10.0))
        ref = np.asarray(ref, dtype=np.float32)
        np.testing.assert_array_equal(y_scores.argmax(1), ref.argmax(1))
        assert np.abs(y_scores - ref).max() <= 1 / 256 + 1e-5
        # the reference-compatible float view of a non-native file comes from the device too
        y = io.load_audio_window(files[0], sample_rate=22050, max_duration=60)
        want, _ = O.load_window(raws[0][0], "s16", 2, 48000, 22050)
        assert y.shape == want.shape and np.abs(y - want).max() <= 2e-6
        # protocol path (make_chunks_for_file + predict + host pooling)
        class ProtocolOnly:
            predict = staticmethod(runner.predict)
            frontend = staticmethod(runner.frontend)
            frontend_wave = staticmethod(runner.frontend_wave)
            device = 0

        _, _, _, ys2 = evaluate(ProtocolOnly(), files, classes, cfg, pooling="lme", batch_size=4)
        np.testing.assert_allclose(ys2, y_scores, rtol=0, atol=3e-6)
    finally:
        runner.close()


def _ref_bootstrap(y_true, y_scores, classes, n_bootstrap, confidence, seed):
    """The reference's formulation (`metrics.py:240-322`): materialise every resample and call scikit-learn on it."""
    from sklearn.metrics import average_precision_score

    rng = np.random.default_rng(seed)
    n = y_true.shape[0]
    alpha = (1 - confidence) / 2
    out = []
    for ci, name in enumerate(classes):
        ct, cs = y_true[:, ci], y_scores[:, ci]
        n_pos = int(ct.sum())
        ap = float(average_precision_score(ct, cs))
        if n_pos == 0 or n_pos == n:
            out.append((ap, ap, ap))
            continue
        boot = []
        for _ in range(n_bootstrap):
            idx = rng.integers(0, n, size=n)
            bt, bs = ct[idx], cs[idx]
            if bt.sum() == 0 or bt.sum() == len(bt):
                continue
            boot.append(float(average_precision_score(bt, bs)))
        out.append((ap, float(np.percentile(boot, 100 * alpha)), float(np.percentile(boot, 100 * (1 - alpha)))) if boot else (ap, ap, ap))
    return out


def test_metric_consumers_match_the_reference_formulations():
    import warnings

    from birdnet_stm32.evaluation.metrics import bootstrap_ap_ci, compute_det_curve, optimize_thresholds

    rng = np.random.default_rng(4)
    n, C = 90, 5
    y_true = np.zeros((n, C), dtype=np.float32)
    y_true[np.arange(n), rng.integers(0, 3, size=n)] = 1.0            # classes 3 and 4 have no positives
    y_true[:5, 1] = 1.0
    y_scores = (np.round(rng.random((n, C)) * 32) / 32).astype(np.float32)   # ties
    classes = [f"c{i}" for i in range(C)]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got = bootstrap_ap_ci(y_true, y_scores, classes, n_bootstrap=60, confidence=0.9, seed=7)
        want = _ref_bootstrap(y_true, y_scores, classes, 60, 0.9, 7)
    for g, w in zip(got, want):
        assert abs(g["ap"] - w[0]) <= 1e-12 and abs(g["ci_lower"] - w[1]) <= 1e-12 and abs(g["ci_upper"] - w[2]) <= 1e-12, (g, w)
    assert got[3]["n_positive"] == 0 and got[3]["ci_lower"] == got[3]["ap"]
    # DET curve: the reference's per-threshold rescan
    far, frr, thr = compute_det_curve(y_true, y_scores)
    y_t, y_s = y_true.ravel(), y_scores.ravel()
    uniq = np.unique(y_s)[::-1]
    P, N = y_t.sum(), len(y_t) - y_t.sum()
    np.testing.assert_array_equal(thr, uniq.astype(np.float64))
    for k in (0, len(uniq) // 2, len(uniq) - 1):
        pred = y_s >= uniq[k]
        assert far[k] == np.sum(1 - y_t[pred]) / N and frr[k] == (P - np.sum(y_t[pred])) / P
    assert compute_det_curve(np.zeros(4), np.ones(4))[2].tolist() == [0.5]
    th = optimize_thresholds(y_true, y_scores, classes)
    assert th["c3"] == 0.5 and 0.0 <= th["c0"] <= 1.0 and set(th) == set(classes)


def test_native_reader_path_equals_python_reader_path(tmp_path):
    """evaluate()'s device path with the C++ batch reader vs the Python reader threads, through a stub runner whose
    pooled scores are a checksum of the chunks it is handed: same files, same chunks, same order, same results."""
    import warnings

    from birdnet_stm32.evaluation.metrics import evaluate

    classes = ["bird_a", "bird_b", "bird_c"]
    files = make_dataset(str(tmp_path), classes, n_per_class=4, sr=22050, seconds=(3.0, 7.4, 1.2, 6.0))
    (tmp_path / "bird_a" / "broken.wav").write_bytes(b"RIFFxxxxWAVEnope")
    files.insert(3, str(tmp_path / "bird_a" / "broken.wav"))

    class Stub:
        device = 0

        def predict_pooled(self, pcm, peak, offs, pooling="avg", beta=10.0):
            out = np.zeros((len(offs) - 1, len(classes)), dtype=np.float32)
            for f in range(len(offs) - 1):
                seg = pcm[offs[f]:offs[f + 1]].astype(np.float64)
                out[f] = [abs(seg.sum()) % 97 / 97.0, (np.abs(seg).sum() % 89) / 89.0, float(peak[offs[f]])]
            return out

    cfg = dict(RAW_CFG, audio_frontend="hybrid")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        a = evaluate(Stub(), files, classes, cfg, pooling="avg", device_batch_chunks=5, native_reader=True, io_workers=3)
        b = evaluate(Stub(), files, classes, cfg, pooling="avg", device_batch_chunks=5, native_reader=False, io_workers=3)
    assert [f["file"] for f in a[1]] == [f["file"] for f in b[1]] and len(a[1]) == 12
    np.testing.assert_array_equal(a[3], b[3])
    np.testing.assert_array_equal(a[2], b[2])
    assert a[0]["skipped_files"] == b[0]["skipped_files"] == 1


def test_read_foreign_native_and_fallback(tmp_path):
    from test_ingest import write_wav

    from birdnet_stm32.audio import io
    from birdnet_stm32.evaluation.metrics import _read_foreign

    rng = np.random.default_rng(2)
    paths = []
    for i, (kind, ch, sr, n) in enumerate((("s16", 2, 48000, 50000), ("f32", 1, 44100, 30000), ("s24", 2, 32000, 20000))):
        raw = rng.standard_normal(n * ch).astype("<f4") if kind == "f32" else (
            rng.integers(0, 256, size=n * ch * 3, dtype=np.uint8) if kind == "s24" else rng.integers(-9000, 9000, size=n * ch).astype("<i2"))
        p = str(tmp_path / f"x{i}.wav")
        write_wav(p, raw, kind, ch, sr)
        paths.append(p)
    paths.insert(1, str(tmp_path / "gone.wav"))
    for stage in (np.zeros(1 << 20, np.uint8), np.zeros(250000, np.uint8), None):      # all fit / only the first fits / Python parser only
        items = _read_foreign(paths, stage, 3)
        assert len(items) == 4 and items[1] is None
        for path, item in zip(paths, items):
            if item is None:
                continue
            want = io.read_wav_frames(path, 60)
            assert item[1:] == want[1:] and np.array_equal(item[0], want[0])


def test_device_path_control_flow_with_foreign_files_simulated_on_cpu(tmp_path, monkeypatch):
    """The device path of evaluate() -- native PCM16 batches, files routed through the ingest, ordering, skips -- driven on
    the CPU with stand-ins for the CUDA pieces: torch tensors live in host memory, GpuIngest is replaced by the ingest
    oracle writing through the same pointer interface, and the runner returns checksums of the chunks it is handed.
    Both reader paths (C++ batch reader / Python threads) must produce identical results in file order."""
    import ctypes
    import warnings

    import torch

    from test_ingest import write_wav

    from birdnet_stm32.audio import ingest as ingest_mod
    from birdnet_stm32.evaluation.metrics import evaluate
    from oracle import bn_ingest_oracle as O

    classes = ["bird_a", "bird_b"]
    sr, cd = 8000, 1.0
    T = int(sr * cd)
    files = make_dataset(str(tmp_path), classes, n_per_class=3, sr=sr, seconds=(1.0, 2.6, 0.4))
    rng = np.random.default_rng(3)
    for i, (kind, ch, sr0, n) in enumerate((("s16", 2, 16000, 30000), ("f32", 1, 11025, 9000), ("s16", 1, 4000, 7000))):
        raw = rng.standard_normal(n * ch).astype("<f4") * 0.2 if kind == "f32" else rng.integers(-9000, 9000, size=n * ch).astype("<i2")
        p = str(tmp_path / classes[i % 2] / f"foreign{i}.wav")
        write_wav(p, raw, kind, ch, sr0)
        files.insert(2 * i + 1, p)
    (tmp_path / "bird_b" / "junk.wav").write_bytes(b"not a wav file at all")
    files.append(str(tmp_path / "bird_b" / "junk.wav"))

    class FakeIngest:
        def __init__(self, device=0):
            self.device = device

        def chunks_to_ptr(self, raw, kind, channels, sr_in, sr_out, chunk_len, step, out_ptr, max_chunks):
            wave, _ = O.load_window(np.asarray(raw), kind, channels, sr_in, sr_out)
            ch = O.split_chunks(wave, sr_out, chunk_len / sr_out, 0.0)
            assert ch.shape[0] <= max_chunks and ch.shape[1] == chunk_len
            ctypes.memmove(out_ptr, np.ascontiguousarray(ch, dtype=np.float32).ctypes.data, ch.nbytes)
            return ch.shape[0]

        def close(self):
            pass

    real_device = torch.device
    monkeypatch.setattr(ingest_mod, "GpuIngest", FakeIngest)
    monkeypatch.setattr(torch, "device", lambda *a, **k: real_device("cpu"))
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)

    class Stub:
        device = 0

        def predict_pooled(self, pcm, peak, offs, pooling="avg", beta=10.0):
            out = np.zeros((len(offs) - 1, len(classes)), dtype=np.float32)
            for f in range(len(offs) - 1):
                seg = pcm[offs[f]:offs[f + 1]].astype(np.float64)
                out[f] = [abs(seg.sum()) % 97 / 97.0, float(peak[offs[f]])]
            return out

        def infer_pool_wave_ptr(self, wave_ptr, peak_ptr, offs_ptr, F, pooling, beta, out_ptr, stream=None):
            offs = np.ctypeslib.as_array(ctypes.cast(offs_ptr, ctypes.POINTER(ctypes.c_int32)), shape=(F + 1,))
            wave = np.ctypeslib.as_array(ctypes.cast(wave_ptr, ctypes.POINTER(ctypes.c_float)), shape=(int(offs[F]), T))
            out = np.ctypeslib.as_array(ctypes.cast(out_ptr, ctypes.POINTER(ctypes.c_float)), shape=(F, len(classes)))
            for f in range(F):
                seg = wave[offs[f]:offs[f + 1]].astype(np.float64)
                out[f] = [abs(seg.sum()) % 1.0, float(np.abs(seg).max())]

    cfg = dict(RAW_CFG, audio_frontend="hybrid", sample_rate=sr, chunk_duration=cd)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        a = evaluate(Stub(), files, classes, cfg, pooling="avg", device_batch_chunks=4, native_reader=True, io_workers=3)
        b = evaluate(Stub(), files, classes, cfg, pooling="avg", device_batch_chunks=4, native_reader=False, io_workers=3)
    want_files = [f for f in files if not f.endswith("junk.wav")]
    assert [f["file"] for f in a[1]] == want_files == [f["file"] for f in b[1]]
    np.testing.assert_array_equal(a[3], b[3])
    np.testing.assert_array_equal(a[2], b[2])
    assert a[0]["skipped_files"] == b[0]["skipped_files"] == 1
    # the foreign files really went through the ingest stand-in: their second score is max|normalised wave| = 1
    for row, f in zip(a[3], want_files):
        if "foreign" in f:
            assert abs(row[1] - 1.0) < 1e-6
