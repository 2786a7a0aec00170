"""Metric tail on the GPU: the device twin of `_metrics_from_scores` (reference `evaluation/metrics.py:152-190`).

Same dictionary keys and the same definitions as the scikit-learn calls the reference makes (`roc_auc_score`
micro, `average_precision_score` per class and micro, precision / recall / F1 at 0.5); the sorting, counting and
float64 accumulation run in `bn_metrics_compute` (`include/bn_metrics.h`, `csrc/bn_metrics.cu`).  Meant for the
large file-sharded evaluations (BASELINE config 5: 100k files x 100 classes = 10M cells, where sklearn spends
seconds sorting on one core); `evaluate(..., metrics_backend="device")` selects it.
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from birdnet_stm32 import _lib as L


def metrics_from_scores_device(y_true: np.ndarray, y_scores: np.ndarray, device: int = 0) -> dict:
    yt = np.ascontiguousarray(y_true, dtype=np.float32)
    ys = np.ascontiguousarray(y_scores, dtype=np.float32)
    if yt.ndim != 2 or yt.shape != ys.shape:
        raise ValueError("y_true and y_scores must both be [n_files, n_classes]")
    F, Cn = yt.shape
    res = L.BnMetricsResult()
    aps = np.zeros((Cn,), dtype=np.float64)
    lib = L.load()
    L.check(lib.bn_metrics_compute(yt.ctypes.data_as(C.c_void_p), ys.ctypes.data_as(C.c_void_p), F, Cn, int(device), C.byref(res),
                                   aps.ctypes.data_as(C.c_void_p)))
    return {
        "roc-auc": float(res.roc_auc_micro), "f1": float(res.f1), "precision": float(res.precision), "recall": float(res.recall),
        "ap_per_class": [float(a) for a in aps], "cmAP": float(res.cmap), "mAP": float(res.map_micro),
    }


def metrics_from_device_ptrs(y_true_ptr: int, y_scores_ptr: int, n_files: int, n_classes: int, device: int = 0):
    """Same on device-resident float32 matrices (e.g. the all-gathered score tensor); returns (result struct, ap_per_class)."""
    res = L.BnMetricsResult()
    aps = np.zeros((n_classes,), dtype=np.float64)
    L.check(L.load().bn_metrics_compute(C.c_void_p(y_true_ptr), C.c_void_p(y_scores_ptr), int(n_files), int(n_classes), int(device),
                                        C.byref(res), aps.ctypes.data_as(C.c_void_p)))
    return res, aps


def bootstrap_ap_samples_device(y_true: np.ndarray, y_scores: np.ndarray, n_bootstrap: int = 1000, seed: int = 42, device: int = 0,
                                multiplicities: np.ndarray | None = None) -> np.ndarray:
    """AP of `n_bootstrap` resamples of the files for every class: float64 `[C, n_bootstrap]`, NaN where a resample holds a
    single class.  `multiplicities` (int32 `[n_bootstrap, F]`, how often each file is drawn) replaces the device generator."""
    yt = np.ascontiguousarray(y_true, dtype=np.float32)
    ys = np.ascontiguousarray(y_scores, dtype=np.float32)
    if yt.ndim != 2 or yt.shape != ys.shape:
        raise ValueError("y_true and y_scores must both be [n_files, n_classes]")
    F, Cn = yt.shape
    out = np.empty((Cn, int(n_bootstrap)), dtype=np.float64)
    mult = None
    if multiplicities is not None:
        mult = np.ascontiguousarray(multiplicities, dtype=np.int32)
        if mult.shape != (int(n_bootstrap), F):
            raise ValueError("multiplicities must be int32 [n_bootstrap, n_files]")
    L.check(L.load().bn_metrics_bootstrap_ap(yt.ctypes.data_as(C.c_void_p), ys.ctypes.data_as(C.c_void_p), F, Cn, int(n_bootstrap),
                                             C.c_uint64(int(seed) & 0xFFFFFFFFFFFFFFFF), mult.ctypes.data_as(C.c_void_p) if mult is not None else None,
                                             out.ctypes.data_as(C.c_void_p), int(device)))
    return out


def bootstrap_ap_ci_device(y_true: np.ndarray, y_scores: np.ndarray, classes: list[str], n_bootstrap: int = 1000, confidence: float = 0.95,
                           seed: int = 42, device: int = 0) -> list[dict]:
    """Device twin of `bootstrap_ap_ci` (reference `evaluation/metrics.py:240-322`): same list of dicts (`class`, `ap`, `ci_lower`,
    `ci_upper`, `n_positive`, `n_total`), same skipping of degenerate classes and single-class resamples, same percentiles.
    The resamples come from a counter-based generator on the GPU instead of numpy's PCG64 stream, so the interval bounds agree
    with the reference's statistically (Monte-Carlo error of 1000 resamples), not digit for digit; the point estimates `ap` are
    scikit-learn's to 1e-15 (`bn_metrics_compute`)."""
    yt = np.ascontiguousarray(y_true, dtype=np.float32)
    ys = np.ascontiguousarray(y_scores, dtype=np.float32)
    n = yt.shape[0]
    aps = metrics_from_scores_device(yt, ys, device)["ap_per_class"]
    samples = bootstrap_ap_samples_device(yt, ys, n_bootstrap, seed, device)
    alpha = (1 - confidence) / 2
    out: list[dict] = []
    for ci, name in enumerate(classes):
        n_pos = int((yt[:, ci] != 0).sum())
        ap = float(aps[ci])
        lo = hi = ap
        if 0 < n_pos < n:
            boot = samples[ci][~np.isnan(samples[ci])]
            if boot.size:
                lo, hi = float(np.percentile(boot, 100 * alpha)), float(np.percentile(boot, 100 * (1 - alpha)))
        out.append({"class": name, "ap": ap, "ci_lower": lo, "ci_upper": hi, "n_positive": n_pos, "n_total": n})
    return out
