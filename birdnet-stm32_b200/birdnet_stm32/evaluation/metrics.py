"""`evaluate()`: file-level evaluation, twin of the reference `evaluation/metrics.py:75-207`.

Same signature, same 4-tuple `(metrics, per_file, y_true, y_scores)`, same metric keys and the same
sklearn calls for ROC-AUC / AP / cmAP / F1, same skip rules (unknown label folder, unreadable file) and the
same `RuntimeError` when nothing is left.  What changes is the hot loop: with a `GpuRunner` the chunks
of *many* files are sent to the device together as PCM16 and come back already pooled per file
(`bn_infer_pool`), instead of one `predict()` per <= batch_size chunks of one file -- that loop lives in
`evaluation/device_eval.py` (reader / batcher / collector).  Any other object with `predict(x_batch)` (e.g. the
reference tests' FakeRunner) still works through the per-file protocol path below.
"""

from __future__ import annotations

import math
import os
import resource
import time

import numpy as np

from birdnet_stm32.audio.io import UnsupportedAudio, load_pcm16_chunks, read_wav_frames
from birdnet_stm32.evaluation.device_eval import read_foreign as _read_foreign  # noqa: F401  (re-exported for the tests)
from birdnet_stm32.evaluation.device_eval import run_device_path, skip_reason
from birdnet_stm32.evaluation.pooling import pool_scores
from birdnet_stm32.models.frontend import normalize_frontend_name


class NoValidSamples(RuntimeError):
    """`evaluate()` found no file it could score (the reference raises a plain RuntimeError with the same message)."""


def _foreign_float_chunks(path: str, cfg: dict, chunk_overlap: float, device: int) -> np.ndarray | None:
    """Float32 chunks `[n, T]` of a file that is not mono <= 16-bit PCM at the model rate: decoded, mixed, resampled and
    peak-normalised on the device (`audio/ingest.py`), i.e. what `load_audio_file` returns in the reference."""
    from birdnet_stm32.audio.ingest import chunk_step, shared_ingest

    try:
        raw, kind, ch, sr0 = read_wav_frames(path, 60)
    except Exception:
        return None
    T, step = chunk_step(int(cfg["sample_rate"]), float(cfg["chunk_duration"]), chunk_overlap)
    return shared_ingest(device).chunks(raw, kind, ch, sr0, int(cfg["sample_rate"]), T, step)


def make_chunks_for_file(path: str, cfg: dict, frontend: str, mag_scale: str, n_fft: int, chunk_overlap: float,
                         frontend_runner=None) -> list[np.ndarray]:
    """Model-ready inputs for one file (reference `metrics.py:18-72`).

    `raw`: `x / (max|x| + 1e-6)` as `[T, 1]` (host, trivial).  `hybrid`: `[fft_bins, spec_width, 1]`
    spectrograms computed by the B200 frontend kernel of `frontend_runner` (a GpuRunner); there is no
    host STFT in this package.  `librosa`: mel features from the CUDA feature kernels.  Files that are not mono
    <= 16-bit PCM at the model rate go through the device ingest first, for every frontend.  Unreadable files give `[]`.
    """
    sr, cd = int(cfg["sample_rate"]), float(cfg["chunk_duration"])
    device = int(getattr(frontend_runner, "device", 0) or 0)
    wave = None                       # float32 chunks of a foreign file
    try:
        pcm, peak = load_pcm16_chunks(path, sr, cd, chunk_overlap, max_duration=60)
    except UnsupportedAudio:
        wave = _foreign_float_chunks(path, cfg, chunk_overlap, device)
        if wave is None or wave.shape[0] == 0:
            return []
        pcm, peak = None, np.float32(0)
    except Exception:
        return []
    if wave is None and pcm.shape[0] == 0:
        return []
    if frontend == "raw":
        if wave is None:
            wave = pcm.astype(np.float32) / np.float32(32768.0)
            if peak > 0:
                wave = wave / np.float32(peak)
        return [(ch / (np.max(np.abs(ch)) + 1e-6))[:, None].astype(np.float32) for ch in wave]
    if frontend == "hybrid":
        if frontend_runner is None or (wave is not None and not hasattr(frontend_runner, "frontend_wave")):
            raise RuntimeError("hybrid inputs are computed on the GPU: pass frontend_runner=GpuRunner(...) "
                               "(this package has no CPU spectrogram path)")
        if wave is not None:
            return [s for s in frontend_runner.frontend_wave(wave)]
        spec = frontend_runner.frontend(pcm, np.full((pcm.shape[0],), peak, dtype=np.float32))
        return [s for s in spec]
    if frontend == "librosa":
        # precomputed mel spectrograms (reference `metrics.py:49-54`): computed for all chunks of the file in one
        # call by the CUDA feature kernels
        from birdnet_stm32.audio.spectrogram import FeatureExtractor, audio_to_pcm16_peak

        if wave is not None:
            # the feature kernels take PCM16 + peak: a resampled float window is re-expressed as int16 codes of its own
            # peak-normalised range (quantisation step 2^-15 of the peak, below the 1e-4 feature tolerance)
            conv = [audio_to_pcm16_peak(ch) for ch in wave]
            pcm = np.stack([c[0] for c in conv])
            peaks = np.array([c[1] for c in conv], dtype=np.float32)
        else:
            peaks = np.full((pcm.shape[0],), peak, dtype=np.float32)
        key = (sr, pcm.shape[1], n_fft, int(cfg["num_mels"]), int(cfg["spec_width"]), mag_scale)
        cache = getattr(make_chunks_for_file, "_fx", None)
        if cache is None or cache[0] != key:
            if cache is not None:
                cache[1].close()
            cache = (key, FeatureExtractor(sr, pcm.shape[1], n_fft, int(cfg["num_mels"]), int(cfg["spec_width"]), mag_scale))
            make_chunks_for_file._fx = cache
        S = cache[1](pcm, peaks)
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
             frontend_runner=None, metrics_backend: str = "sklearn", io_workers: int = 8,
             native_reader: bool = True, strict_files: bool = False) -> tuple[dict, list[dict], np.ndarray, np.ndarray]:
    """Run inference per chunk, pool to file level and compute metrics (see module docstring).

    metrics_backend: "sklearn" (the reference's own calls, host) or "device" (`evaluation/device_metrics.py`: the same
    definitions evaluated by `bn_metrics_compute` on the GPU -- for evaluations with millions of (file, class) cells).
    io_workers: reader threads of the device path (files are read and cut into chunks ahead of the GPU calls, in order).
    native_reader: device path only -- read the files with the C++ thread pool of `bn_read_pcm16_batch` straight into
    pinned batch buffers (double-buffered against the GPU calls) instead of the Python reader threads.
    strict_files: raise instead of skipping when a file cannot be decoded (the reference skips silently; `metrics`
    reports `skipped_files` and `skipped_by_reason` either way)."""
    if metrics_backend not in ("sklearn", "device"):
        raise ValueError(f"Unsupported metrics backend: {metrics_backend}")
    frontend = normalize_frontend_name(cfg["audio_frontend"])
    mag_scale = cfg.get("mag_scale", "none")
    n_fft = int(cfg["fft_length"])
    sr, cd = int(cfg["sample_rate"]), float(cfg["chunk_duration"])
    class_index = {c: i for i, c in enumerate(classes)}
    rss_before_kb = resource.getrusage(resource.RUSAGE_SELF).ru_maxrss if profile_memory else 0
    latencies_ms: list[float] = []
    total_chunks = 0
    skipped: dict[str, int] = {}
    skipped_paths: list[tuple[str, str]] = []

    todo = [p for p in files if os.path.basename(os.path.dirname(p)) in class_index]
    if hasattr(model_runner, "predict_pooled") and frontend == "hybrid":
        col, total_chunks, latencies_ms = run_device_path(model_runner, todo, classes, sr, cd, overlap, pooling, mep_beta,
                                                          device_batch_chunks, io_workers, native_reader, measure_latency)
        y_true, y_scores, per_file = col.y_true, col.y_scores, col.per_file
        skipped, skipped_paths = col.skipped, col.skipped_paths
    else:
        y_true, y_scores, per_file = [], [], []
        fr = frontend_runner if frontend_runner is not None else (model_runner if hasattr(model_runner, "frontend") else None)
        for path in todo:
            label = os.path.basename(os.path.dirname(path))
            chunks = make_chunks_for_file(path, cfg, frontend, mag_scale, n_fft, overlap, frontend_runner=fr)
            if len(chunks) == 0:
                why = skip_reason(path)
                skipped[why] = skipped.get(why, 0) + 1
                skipped_paths.append((path, why))
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
            target = np.zeros((len(classes),), dtype=np.float32)
            target[class_index[label]] = 1.0
            y_true.append(target)
            y_scores.append(pooled)
            per_file.append({"file": path, "label": label, "scores": pooled.tolist()})

    n_skipped = sum(skipped.values())
    if strict_files and n_skipped:
        raise RuntimeError(f"{n_skipped} file(s) could not be decoded {skipped}: " + ", ".join(p for p, _ in skipped_paths[:5]))
    if len(y_true) == 0:
        raise NoValidSamples("No valid test samples found for the provided class set.")

    y_true_arr = np.asarray(y_true, dtype=np.float32)
    y_scores_arr = np.asarray(y_scores, dtype=np.float32)
    if metrics_backend == "device":
        from birdnet_stm32.evaluation.device_metrics import metrics_from_scores_device

        metrics = metrics_from_scores_device(y_true_arr, y_scores_arr, int(getattr(model_runner, "device", 0)))
    else:
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
    if n_skipped:
        metrics["skipped_files"] = n_skipped
        metrics["skipped_by_reason"] = dict(skipped)
    return metrics, per_file, y_true_arr, y_scores_arr


# ---------------------------------------------------------------------------------------------------------
# consumers of evaluate()'s (y_true, y_scores): threshold tuning, bootstrap intervals, DET curve
# (reference `evaluation/metrics.py:210-372`; same results, sort-once formulations instead of per-threshold /
# per-resample rescans so that they stay usable on the 100k-file evaluations the GPU path makes routine)
# ---------------------------------------------------------------------------------------------------------
def optimize_thresholds(y_true: np.ndarray, y_scores: np.ndarray, classes: list[str]) -> dict[str, float]:
    """Score threshold with the best F1 per class (reference `metrics.py:210-237`: argmax of F1 along scikit-learn's
    precision-recall curve, 0.5 for a class without positives).  Computed here from one descending sort per class:
    at the k-th distinct score, TP = positives at or above it and F1 = 2 TP / (n_at_or_above + P); ties resolve to the
    lowest such threshold, which is where `precision_recall_curve`'s ascending thresholds put the first maximum."""
    best: dict[str, float] = {}
    for ci, name in enumerate(classes):
        t = np.asarray(y_true[:, ci]) != 0
        n_pos = int(t.sum())
        if n_pos == 0:
            best[name] = 0.5
            continue
        s = np.asarray(y_scores[:, ci])
        order = np.argsort(-s, kind="stable")
        ss = s[order]
        last = np.flatnonzero(np.r_[ss[1:] != ss[:-1], True])         # last index of every run of equal scores
        tp = np.cumsum(t[order])[last].astype(np.float64)
        precision, recall = tp / (last + 1.0), tp / n_pos
        # scikit-learn drops the thresholds below the one that first reaches full recall; F1 with its 1e-12 guard
        full = int(np.argmax(tp == n_pos))
        f1 = 2.0 * precision[: full + 1] * recall[: full + 1] / (precision[: full + 1] + recall[: full + 1] + 1e-12)
        # ascending-threshold order = reversed; argmax takes the first maximum there
        k = full - int(np.argmax(f1[::-1]))
        best[name] = float(ss[last[k]])
    return best


def _weighted_ap_sorted(labels_sorted: np.ndarray, group_end: np.ndarray, weights: np.ndarray) -> np.ndarray:
    """Average precision of resamples given as per-sample multiplicities.

    labels_sorted: [n] 0/1 in descending-score order; group_end: indices of the last element of every run of equal
    scores; weights: [R, n] multiplicities in the same order.  Equals sklearn's average_precision_score on the
    materialised resample: thresholds are the distinct scores, AP = sum (R_k - R_{k-1}) P_k."""
    w = weights.astype(np.float64)
    tp = np.cumsum(w * labels_sorted[None, :], axis=1)[:, group_end]
    cnt = np.cumsum(w, axis=1)[:, group_end]
    P = tp[:, -1:]
    with np.errstate(divide="ignore", invalid="ignore"):
        prec = np.where(cnt > 0, tp / cnt, 0.0)
        rec = tp / P
    drec = np.diff(np.concatenate([np.zeros((w.shape[0], 1)), rec], axis=1), axis=1)
    return np.sum(drec * prec, axis=1)


def bootstrap_ap_ci(y_true: np.ndarray, y_scores: np.ndarray, classes: list[str], n_bootstrap: int = 1000,
                    confidence: float = 0.95, seed: int = 42) -> list[dict]:
    """Per-class AP with bootstrap confidence intervals (reference `metrics.py:240-322`).

    Same random stream as the reference (`default_rng(seed)`, one `integers(0, n, size=n)` draw per resample, classes in
    order, degenerate classes skipped before drawing), same skipping of single-class resamples and the same percentiles.
    Each class is sorted once; a resample is a vector of multiplicities, so its AP costs two cumulative sums instead of a
    sort inside scikit-learn."""
    from sklearn.metrics import average_precision_score

    rng = np.random.default_rng(seed)
    n = y_true.shape[0]
    alpha = (1 - confidence) / 2
    results: list[dict] = []
    for ci, name in enumerate(classes):
        col_true, col_scores = y_true[:, ci], y_scores[:, ci]
        n_pos = int(col_true.sum())
        try:
            ap = float(average_precision_score(col_true, col_scores))
        except Exception:
            ap = float("nan")
        if n_pos == 0 or n_pos == n:
            results.append({"class": name, "ap": ap, "ci_lower": ap, "ci_upper": ap, "n_positive": n_pos, "n_total": n})
            continue
        order = np.argsort(-col_scores, kind="stable")
        s_sorted = col_scores[order]
        lab_sorted = (col_true[order] != 0).astype(np.float64)
        group_end = np.flatnonzero(np.r_[s_sorted[1:] != s_sorted[:-1], True])
        inv = np.empty(n, dtype=np.int64)
        inv[order] = np.arange(n)                       # original index -> position in sorted order
        boot: list[float] = []
        block = max(1, min(n_bootstrap, (1 << 22) // max(n, 1)))
        for b0 in range(0, n_bootstrap, block):
            nb = min(block, n_bootstrap - b0)
            weights = np.zeros((nb, n), dtype=np.int32)
            for r in range(nb):
                idx = rng.integers(0, n, size=n)
                weights[r] = np.bincount(inv[idx], minlength=n)
            pos = weights @ lab_sorted
            keep = (pos > 0) & (pos < n)
            if keep.any():
                boot.extend(_weighted_ap_sorted(lab_sorted, group_end, weights[keep]).tolist())
        if boot:
            lo, hi = float(np.percentile(boot, 100 * alpha)), float(np.percentile(boot, 100 * (1 - alpha)))
        else:
            lo = hi = ap
        results.append({"class": name, "ap": ap, "ci_lower": lo, "ci_upper": hi, "n_positive": n_pos, "n_total": n})
    return results


def compute_det_curve(y_true: np.ndarray, y_scores: np.ndarray) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    """DET curve points (far, frr, thresholds), one per distinct score in descending order (reference `metrics.py:325-372`)."""
    y_t = np.asarray(y_true).ravel()
    y_s = np.asarray(y_scores).ravel()
    total_pos = float(y_t.sum())
    total_neg = float(len(y_t) - total_pos)
    if total_pos == 0 or total_neg == 0:
        return np.array([0.0]), np.array([0.0]), np.array([0.5])
    order = np.argsort(-y_s, kind="stable")
    s_sorted = y_s[order]
    end = np.flatnonzero(np.r_[s_sorted[1:] != s_sorted[:-1], True])    # last index of each run: "score >= threshold"
    tp = np.cumsum(y_t[order].astype(np.float64))[end]
    fp = (end + 1).astype(np.float64) - tp
    out_dt = y_t.dtype if np.issubdtype(y_t.dtype, np.floating) else np.float64
    far = (fp / total_neg).astype(out_dt)
    frr = ((total_pos - tp) / total_pos).astype(out_dt)
    return far, frr, s_sorted[end].astype(np.float64)
