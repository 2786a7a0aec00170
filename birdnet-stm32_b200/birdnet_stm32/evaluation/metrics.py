"""`evaluate()`: file-level evaluation, twin of the reference `evaluation/metrics.py:75-207`.

Same signature, same 4-tuple `(metrics, per_file, y_true, y_scores)`, same metric keys and the same
sklearn calls for ROC-AUC / AP / cmAP / F1, same skip rules (unknown label folder, unreadable file) and the
same `RuntimeError` when nothing is left.  What changes is the hot loop: with a `GpuRunner` the chunks
of *many* files are sent to the device together as PCM16 and come back already pooled per file
(`bn_infer_pool`), instead of one `predict()` per <= batch_size chunks of one file.  Any other object with
`predict(x_batch)` (e.g. the reference tests' FakeRunner) still works through the per-file protocol path.
"""

from __future__ import annotations

import math
import os
import resource
import time

import numpy as np

from birdnet_stm32.audio.io import load_pcm16_chunks
from birdnet_stm32.evaluation.pooling import pool_scores
from birdnet_stm32.models.frontend import normalize_frontend_name


def make_chunks_for_file(path: str, cfg: dict, frontend: str, mag_scale: str, n_fft: int, chunk_overlap: float,
                         frontend_runner=None) -> list[np.ndarray]:
    """Model-ready inputs for one file (reference `metrics.py:18-72`).

    `raw`: `x / (max|x| + 1e-6)` as `[T, 1]` (host, trivial).  `hybrid`: `[fft_bins, spec_width, 1]`
    spectrograms computed by the B200 frontend kernel of `frontend_runner` (a GpuRunner); there is no
    host STFT in this package.  Unreadable files give `[]`.
    """
    sr, cd = int(cfg["sample_rate"]), float(cfg["chunk_duration"])
    try:
        pcm, peak = load_pcm16_chunks(path, sr, cd, chunk_overlap, max_duration=60)
    except Exception:
        return []
    if pcm.shape[0] == 0:
        return []
    if frontend == "raw":
        x = pcm.astype(np.float32) / np.float32(32768.0)
        if peak > 0:
            x = x / np.float32(peak)
        out = []
        for ch in x:
            out.append((ch / (np.max(np.abs(ch)) + 1e-6))[:, None].astype(np.float32))
        return out
    if frontend == "hybrid":
        if frontend_runner is None:
            raise RuntimeError("hybrid inputs are computed on the GPU: pass frontend_runner=GpuRunner(...) "
                               "(this package has no CPU spectrogram path)")
        spec = frontend_runner.frontend(pcm, np.full((pcm.shape[0],), peak, dtype=np.float32))
        return [s for s in spec]
    if frontend == "librosa":
        # precomputed mel spectrograms (reference `metrics.py:49-54`): computed for all chunks of the file in one
        # call by the CUDA feature kernels; `frontend_runner` may carry a cached FeatureExtractor
        from birdnet_stm32.audio.spectrogram import FeatureExtractor

        key = (sr, pcm.shape[1], n_fft, int(cfg["num_mels"]), int(cfg["spec_width"]), mag_scale)
        cache = getattr(make_chunks_for_file, "_fx", None)
        if cache is None or cache[0] != key:
            if cache is not None:
                cache[1].close()
            cache = (key, FeatureExtractor(sr, pcm.shape[1], n_fft, int(cfg["num_mels"]), int(cfg["spec_width"]), mag_scale))
            make_chunks_for_file._fx = cache
        S = cache[1](pcm, np.full((pcm.shape[0],), peak, dtype=np.float32))
        return [s[:, :, None] for s in S]
    raise ValueError(f"Invalid audio_frontend for the B200 path: {frontend}")


def _metrics_from_scores(y_true_arr: np.ndarray, y_scores_arr: np.ndarray) -> dict:
    """ROC-AUC (micro), F1/precision/recall @0.5, per-class AP, cmAP, mAP -- reference `metrics.py:152-190`."""
    from sklearn.metrics import average_precision_score, roc_auc_score

    metrics: dict = {}
    try:
        metrics["roc-auc"] = float(roc_auc_score(y_true_arr, y_scores_arr, average="micro"))
    except Exception:
        metrics["roc-auc"] = float("nan")
    hit = (y_scores_arr >= 0.5).astype(np.float32)
    tp = np.sum(y_true_arr * hit)
    fp = np.sum((1 - y_true_arr) * hit)
    fn = np.sum(y_true_arr * (1 - hit))
    precision = tp / (tp + fp + 1e-12)
    recall = tp / (tp + fn + 1e-12)
    metrics["f1"] = float(2 * (precision * recall) / (precision + recall)) if precision + recall > 0 else 0.0
    metrics["precision"] = float(precision)
    metrics["recall"] = float(recall)
    aps: list[float] = []
    for ci in range(y_true_arr.shape[1]):
        try:
            aps.append(average_precision_score(y_true_arr[:, ci], y_scores_arr[:, ci]))
        except Exception:
            aps.append(np.nan)
    valid = [a for a in aps if not (a is None or (isinstance(a, float) and math.isnan(a)))]
    metrics["ap_per_class"] = aps
    metrics["cmAP"] = float(np.mean(valid)) if valid else float("nan")
    try:
        metrics["mAP"] = float(average_precision_score(y_true_arr, y_scores_arr, average="micro"))
    except Exception:
        metrics["mAP"] = float("nan")
    return metrics


def evaluate(model_runner, files: list[str], classes: list[str], cfg: dict, pooling: str = "average",
             batch_size: int = 64, overlap: float = 0.0, mep_beta: float = 10.0, measure_latency: bool = False,
             profile_memory: bool = False, device_batch_chunks: int = 4096,
             frontend_runner=None) -> tuple[dict, list[dict], np.ndarray, np.ndarray]:
    """Run inference per chunk, pool to file level and compute metrics (see module docstring)."""
    frontend = normalize_frontend_name(cfg["audio_frontend"])
    mag_scale = cfg.get("mag_scale", "none")
    n_fft = int(cfg["fft_length"])
    sr, cd = int(cfg["sample_rate"]), float(cfg["chunk_duration"])
    num_classes = len(classes)
    class_index = {c: i for i, c in enumerate(classes)}

    y_true: list[np.ndarray] = []
    y_scores: list[np.ndarray] = []
    per_file: list[dict] = []
    latencies_ms: list[float] = []
    total_chunks = 0
    skipped = 0
    rss_before_kb = resource.getrusage(resource.RUSAGE_SELF).ru_maxrss if profile_memory else 0

    def target_for(label: str) -> np.ndarray:
        t = np.zeros((num_classes,), dtype=np.float32)
        t[class_index[label]] = 1.0
        return t

    device_path = hasattr(model_runner, "predict_pooled") and frontend == "hybrid"
    if device_path:
        pend_pcm: list[np.ndarray] = []
        pend_peak: list[np.ndarray] = []
        pend_meta: list[tuple[str, str]] = []
        pend_counts: list[int] = []

        def flush():
            nonlocal total_chunks
            if not pend_meta:
                return
            pcm = np.concatenate(pend_pcm, axis=0)
            peak = np.concatenate(pend_peak, axis=0)
            offs = np.zeros(len(pend_counts) + 1, dtype=np.int32)
            offs[1:] = np.cumsum(pend_counts)
            t0 = time.perf_counter()
            pooled = model_runner.predict_pooled(pcm, peak, offs, pooling=pooling, beta=mep_beta)
            if measure_latency:
                per = (time.perf_counter() - t0) * 1000 / max(pcm.shape[0], 1)
                latencies_ms.extend([per] * pcm.shape[0])
            total_chunks += pcm.shape[0]
            for (path, label), row in zip(pend_meta, pooled):
                y_true.append(target_for(label))
                y_scores.append(row)
                per_file.append({"file": path, "label": label, "scores": row.tolist()})
            pend_pcm.clear(); pend_peak.clear(); pend_meta.clear(); pend_counts.clear()

        for path in files:
            label = os.path.basename(os.path.dirname(path))
            if label not in class_index:
                continue
            try:
                pcm, peak = load_pcm16_chunks(path, sr, cd, overlap, max_duration=60)
            except Exception:
                skipped += 1
                continue
            if pcm.shape[0] == 0:
                skipped += 1
                continue
            pend_pcm.append(pcm)
            pend_peak.append(np.full((pcm.shape[0],), peak, dtype=np.float32))
            pend_meta.append((path, label))
            pend_counts.append(pcm.shape[0])
            if sum(pend_counts) >= device_batch_chunks:
                flush()
        flush()
    else:
        fr = frontend_runner if frontend_runner is not None else (model_runner if hasattr(model_runner, "frontend") else None)
        for path in files:
            label = os.path.basename(os.path.dirname(path))
            if label not in class_index:
                continue
            chunks = make_chunks_for_file(path, cfg, frontend, mag_scale, n_fft, overlap, frontend_runner=fr)
            if len(chunks) == 0:
                skipped += 1
                continue
            preds = []
            for i in range(0, len(chunks), batch_size):
                batch = np.stack(chunks[i : i + batch_size], axis=0)
                t0 = time.perf_counter()
                p = model_runner.predict(batch)
                if measure_latency:
                    per = (time.perf_counter() - t0) * 1000 / batch.shape[0]
                    latencies_ms.extend([per] * batch.shape[0])
                preds.append(p)
                total_chunks += batch.shape[0]
            pooled = pool_scores(np.concatenate(preds, axis=0), method=pooling, beta=mep_beta)
            y_true.append(target_for(label))
            y_scores.append(pooled)
            per_file.append({"file": path, "label": label, "scores": pooled.tolist()})

    if len(y_true) == 0:
        raise RuntimeError("No valid test samples found for the provided class set.")

    y_true_arr = np.asarray(y_true, dtype=np.float32)
    y_scores_arr = np.asarray(y_scores, dtype=np.float32)
    metrics = _metrics_from_scores(y_true_arr, y_scores_arr)

    if measure_latency and latencies_ms:
        lat = np.array(latencies_ms)
        metrics["latency_mean_ms"] = float(np.mean(lat))
        metrics["latency_median_ms"] = float(np.median(lat))
        metrics["latency_p95_ms"] = float(np.percentile(lat, 95))
        metrics["latency_p99_ms"] = float(np.percentile(lat, 99))
        metrics["total_chunks"] = total_chunks
    if profile_memory:
        rss_after_kb = resource.getrusage(resource.RUSAGE_SELF).ru_maxrss
        metrics["peak_rss_mb"] = round(rss_after_kb / 1024, 1)
        metrics["rss_delta_mb"] = round((rss_after_kb - rss_before_kb) / 1024, 1)
    if skipped:
        metrics["skipped_files"] = skipped
    return metrics, per_file, y_true_arr, y_scores_arr
