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
