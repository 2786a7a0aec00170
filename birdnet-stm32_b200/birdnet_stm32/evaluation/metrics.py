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

from birdnet_stm32.audio.io import UnsupportedAudio, load_pcm16_chunks, prefetch_ordered, read_wav_frames
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
    except UnsupportedAudio:
        # other rate / channels / sample format: float32 chunks from the device ingest (audio/ingest.py)
        if frontend != "hybrid":
            return []
        if frontend_runner is None or not hasattr(frontend_runner, "frontend_wave"):
            raise RuntimeError("this file needs the device ingest and frontend: pass frontend_runner=GpuRunner(...)")
        try:
            raw, kind, ch, sr0 = read_wav_frames(path, 60)
        except Exception:
            return []
        from birdnet_stm32.audio.ingest import chunk_step, shared_ingest

        T, step = chunk_step(sr, cd, chunk_overlap)
        chunks = shared_ingest(int(getattr(frontend_runner, "device", 0))).chunks(raw, kind, ch, sr0, sr, T, step)
        return [s for s in frontend_runner.frontend_wave(chunks)] if chunks.shape[0] else []
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


def _read_foreign(paths: list[str], raw_stage: np.ndarray | None, threads: int) -> list:
    """`read_wav_frames(path, 60)` for every path (None where it fails), read by the native thread pool into `raw_stage`
    group by group; a file that does not fit the stage, or any reader error, falls back to the Python parser.  The
    returned arrays of one call are views into `raw_stage`: consume them before the next call."""
    out: list = [None] * len(paths)
    done = 0
    if raw_stage is not None and paths:
        try:
            from birdnet_stm32.audio import reader as _rd

            n, items = _rd.read_raw_batch(paths, raw_stage, max_seconds=60, threads=max(1, threads))
            out[:n] = items
            done = n
        except Exception:
            done = 0
    for i in range(done, len(paths)):
        try:
            out[i] = read_wav_frames(paths[i], 60)
        except Exception:
            out[i] = None
    return out


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
             native_reader: bool = True) -> tuple[dict, list[dict], np.ndarray, np.ndarray]:
    """Run inference per chunk, pool to file level and compute metrics (see module docstring).

    metrics_backend: "sklearn" (the reference's own calls, host) or "device" (`evaluation/device_metrics.py`: the same
    definitions evaluated by `bn_metrics_compute` on the GPU -- for evaluations with millions of (file, class) cells).
    io_workers: reader threads of the device path (files are read and cut into chunks ahead of the GPU calls, in order).
    native_reader: device path only -- read the files with the C++ thread pool of `bn_read_pcm16_batch` straight into
    pinned batch buffers (double-buffered against the GPU calls) instead of the Python reader threads."""
    if metrics_backend not in ("sklearn", "device"):
        raise ValueError(f"Unsupported metrics backend: {metrics_backend}")
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
        # Files are batched across file boundaries.  Mono 16-bit files at the model rate travel as PCM16; every other
        # WAV (other rate, several channels, other sample format) is decoded / mixed / resampled / peak-normalised /
        # chunked on the device by bn_ingest_chunks straight into a float32 chunk buffer in HBM (reference: the host
        # side of load_audio_window, audio/io.py:112-128) and classified from there.
        pend: list[dict] = []          # {"path", "label", "kind": "pcm" | "wave", "n", ...} in file order
        wave_state: dict = {}

        def wave_buffer():
            if not wave_state:
                import torch

                from birdnet_stm32.audio.ingest import GpuIngest, chunk_step

                dev = torch.device("cuda", int(getattr(model_runner, "device", 0)))
                T, step = chunk_step(sr, cd, overlap)
                cap = int(device_batch_chunks) + int(60 * sr / step) + 8
                wave_state.update(torch=torch, ingest=GpuIngest(dev.index), T=T, step=step, cap=cap, used=0,
                                  buf=torch.empty((cap, T), dtype=torch.float32, device=dev), dev=dev)
            return wave_state

        def flush():
            nonlocal total_chunks
            if not pend:
                return
            rows: dict[int, np.ndarray] = {}
            t0 = time.perf_counter()
            pcm_ids = [i for i, f in enumerate(pend) if f["kind"] == "pcm"]
            if pcm_ids:
                pcm = np.concatenate([pend[i]["pcm"] for i in pcm_ids], axis=0)
                peak = np.concatenate([pend[i]["peak"] for i in pcm_ids], axis=0)
                offs = np.zeros(len(pcm_ids) + 1, dtype=np.int32)
                offs[1:] = np.cumsum([pend[i]["n"] for i in pcm_ids])
                pooled = model_runner.predict_pooled(pcm, peak, offs, pooling=pooling, beta=mep_beta)
                rows.update(zip(pcm_ids, pooled))
            wav_ids = [i for i, f in enumerate(pend) if f["kind"] == "wave"]
            if wav_ids:
                ws = wave_state
                torch = ws["torch"]
                offs = np.zeros(len(wav_ids) + 1, dtype=np.int32)
                offs[1:] = np.cumsum([pend[i]["n"] for i in wav_ids])
                d_offs = torch.from_numpy(offs).to(ws["dev"])
                d_out = torch.empty((len(wav_ids), num_classes), dtype=torch.float32, device=ws["dev"])
                torch.cuda.synchronize(ws["dev"])
                model_runner.infer_pool_wave_ptr(ws["buf"].data_ptr(), None, d_offs.data_ptr(), len(wav_ids), pooling, mep_beta,
                                                 d_out.data_ptr(), None)
                rows.update(zip(wav_ids, d_out.cpu().numpy()))
                ws["used"] = 0
            n_all = sum(f["n"] for f in pend)
            if measure_latency:
                per = (time.perf_counter() - t0) * 1000 / max(n_all, 1)
                latencies_ms.extend([per] * n_all)
            total_chunks += n_all
            for i, f in enumerate(pend):
                y_true.append(target_for(f["label"]))
                y_scores.append(rows[i])
                per_file.append({"file": f["path"], "label": f["label"], "scores": rows[i].tolist()})
            pend.clear()

        def read_one(path: str):
            """Runs on a reader thread: container parse + chunk cut only (numpy / file I/O, no CUDA calls)."""
            try:
                pcm, peak = load_pcm16_chunks(path, sr, cd, overlap, max_duration=60)
                return ("pcm", pcm, peak) if pcm.shape[0] else ("skip",)
            except UnsupportedAudio:
                try:
                    return ("frames",) + read_wav_frames(path, 60)
                except Exception:
                    return ("skip",)
            except Exception:
                return ("skip",)

        todo = [p for p in files if os.path.basename(os.path.dirname(p)) in class_index]
        if native_reader and todo:
            # ---- native reader: batches of files -> pinned int16 chunk buffers, read of batch k + 1 under the GPU call of k ----
            from concurrent.futures import ThreadPoolExecutor

            from birdnet_stm32 import _lib as _L
            from birdnet_stm32.audio import reader as _rd
            from birdnet_stm32.audio.ingest import chunk_step as _chunk_step

            T, step = _chunk_step(sr, cd, overlap)
            cap = int(device_batch_chunks) + int(60 * sr / step) + 4
            stage = []
            for _ in range(2):
                try:
                    from birdnet_stm32.evaluation.gpu_runner import PinnedArray

                    pa = PinnedArray((cap, T), np.int16)
                    stage.append((pa, pa.array))
                except Exception:                      # no CUDA runtime (stub runners in the CPU tests): ordinary memory
                    stage.append((None, np.empty((cap, T), dtype=np.int16)))
            window = 2048                              # paths offered to one reader call
            raw_stage = None                           # raw frames of the files that need the ingest (allocated on first use)

            def read_batch(start: int, slot: int):
                paths = todo[start:start + window]
                n_files, used, info = _rd.read_pcm16_batch(paths, sr, T, step, stage[slot][1], max_seconds=60, threads=max(1, io_workers))
                return start, n_files, used, info

            with ThreadPoolExecutor(max_workers=1, thread_name_prefix="bn-read") as pool:
                fut = pool.submit(read_batch, 0, 0)
                slot = 0
                while fut is not None:
                    start, n_files, used, info = fut.result()
                    if n_files == 0:
                        raise RuntimeError(f"{todo[start]}: more chunks than the batch buffer holds ({cap})")
                    nxt = start + n_files
                    fut = pool.submit(read_batch, nxt, slot ^ 1) if nxt < len(todo) else None
                    buf = stage[slot][1]
                    rows: dict[int, np.ndarray] = {}
                    t0 = time.perf_counter()
                    ok = [i for i in range(n_files) if info[i].status == _rd.RD_OK]
                    if ok:
                        counts = np.array([info[i].n_chunks for i in ok], dtype=np.int64)
                        offs = np.zeros(len(ok) + 1, dtype=np.int32)
                        offs[1:] = np.cumsum(counts)
                        peak = np.repeat(np.array([info[i].peak for i in ok], dtype=np.float32), counts)
                        pooled = model_runner.predict_pooled(buf[:used], peak, offs, pooling=pooling, beta=mep_beta)
                        rows.update(zip(ok, pooled))
                    n_batch = used
                    foreign = [i for i in range(n_files) if info[i].status == _rd.RD_NEEDS_INGEST]
                    if foreign and raw_stage is None:
                        raw_stage = np.empty(256 << 20, dtype=np.uint8)
                    # the arrays are views into raw_stage, valid until the next group is read
                    for g0 in range(0, len(foreign), 16):
                        group = foreign[g0:g0 + 16]
                        f_items = _read_foreign([todo[start + j] for j in group], raw_stage, io_workers)
                        for i, item in zip(group, f_items):
                            if item is None:
                                continue
                            raw, kind, ch, sr0 = item
                            ws = wave_buffer()
                            n = ws["ingest"].chunks_to_ptr(raw, kind, ch, sr0, sr, ws["T"], ws["step"], ws["buf"].data_ptr(), ws["cap"])
                            if n == 0:
                                continue
                            torch = ws["torch"]
                            d_offs = torch.tensor([0, n], dtype=torch.int32, device=ws["dev"])
                            d_out = torch.empty((1, num_classes), dtype=torch.float32, device=ws["dev"])
                            torch.cuda.synchronize(ws["dev"])
                            model_runner.infer_pool_wave_ptr(ws["buf"].data_ptr(), None, d_offs.data_ptr(), 1, pooling, mep_beta, d_out.data_ptr(), None)
                            rows[i] = d_out.cpu().numpy()[0]
                            n_batch += n
                    if measure_latency and n_batch:
                        per = (time.perf_counter() - t0) * 1000 / n_batch
                        latencies_ms.extend([per] * n_batch)
                    total_chunks += n_batch
                    for i in range(n_files):
                        path = todo[start + i]
                        if i not in rows:
                            skipped += 1
                            continue
                        label = os.path.basename(os.path.dirname(path))
                        y_true.append(target_for(label))
                        y_scores.append(rows[i])
                        per_file.append({"file": path, "label": label, "scores": rows[i].tolist()})
                    slot ^= 1
            for pa, _ in stage:
                if pa is not None:
                    pa.free()
            todo = []
        for path, item in zip(todo, prefetch_ordered(read_one, todo, workers=io_workers)):
            label = os.path.basename(os.path.dirname(path))
            if item[0] == "skip":
                skipped += 1
                continue
            if item[0] == "pcm":
                _, pcm, peak = item
                pend.append({"path": path, "label": label, "kind": "pcm", "n": pcm.shape[0], "pcm": pcm,
                             "peak": np.full((pcm.shape[0],), peak, dtype=np.float32)})
            else:
                _, raw, kind, ch, sr0 = item
                ws = wave_buffer()
                n = ws["ingest"].chunks_to_ptr(raw, kind, ch, sr0, sr, ws["T"], ws["step"],
                                               ws["buf"].data_ptr() + 4 * ws["used"] * ws["T"], ws["cap"] - ws["used"])
                if n == 0:
                    skipped += 1
                    continue
                ws["used"] += n
                pend.append({"path": path, "label": label, "kind": "wave", "n": n})
            if sum(f["n"] for f in pend) >= device_batch_chunks:
                flush()
        flush()
        if wave_state:
            wave_state["ingest"].close()
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
    if skipped:
        metrics["skipped_files"] = skipped
    return metrics, per_file, y_true_arr, y_scores_arr


# ---------------------------------------------------------------------------------------------------------
# consumers of evaluate()'s (y_true, y_scores): threshold tuning, bootstrap intervals, DET curve
# (reference `evaluation/metrics.py:210-372`; same results, sort-once formulations instead of per-threshold /
# per-resample rescans so that they stay usable on the 100k-file evaluations the GPU path makes routine)
# ---------------------------------------------------------------------------------------------------------
def optimize_thresholds(y_true: np.ndarray, y_scores: np.ndarray, classes: list[str]) -> dict[str, float]:
    """Per-class threshold maximising F1 on the precision-recall curve (reference `metrics.py:210-237`)."""
    from sklearn.metrics import precision_recall_curve

    optimal: dict[str, float] = {}
    for ci, name in enumerate(classes):
        col_true, col_scores = y_true[:, ci], y_scores[:, ci]
        if col_true.sum() == 0:
            optimal[name] = 0.5
            continue
        prec, rec, thresholds = precision_recall_curve(col_true, col_scores)
        f1 = 2 * prec[:-1] * rec[:-1] / (prec[:-1] + rec[:-1] + 1e-12)
        optimal[name] = float(thresholds[int(np.argmax(f1))])
    return optimal


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
