"""Chunk-score pooling (API of the reference `evaluation/pooling.py:6-47`).

`pool_scores` / `lme_pooling` keep the reference's signatures, dtype behaviour and errors for callers
that hold chunk scores in numpy (e.g. a custom runner).  With the GPU runner the same three methods run
on the device inside `bn_infer_pool` (`k_pool` in csrc/bn_generic.cu) and never pass through here.
"""

from __future__ import annotations

import numpy as np

_AVG = ("avg", "mean", "average")
_LME = ("lme", "log_mean_exp", "log_mean_exponential")


def lme_pooling(scores: np.ndarray, beta: float = 10.0) -> np.ndarray:
    """log(mean(exp(beta * s))) / beta over chunks (axis 0), max-shifted for stability."""
    if scores.size == 0:
        return scores
    scaled = beta * scores
    peak = scaled.max(axis=0, keepdims=True)
    mean_exp = np.exp(scaled - peak).mean(axis=0, keepdims=True)
    return ((peak + np.log(mean_exp + 1e-12)) / beta).ravel()


def pool_scores(chunk_scores: np.ndarray, method: str = "average", beta: float = 10.0) -> np.ndarray:
    """[N_chunks, C] -> [C].  Empty input gives float32 zeros; bad method / rank raise ValueError."""
    key = method.lower()
    if chunk_scores.ndim != 2:
        raise ValueError("chunk_scores must be [N_chunks, C]")
    if chunk_scores.shape[0] == 0:
        return np.zeros((chunk_scores.shape[1],), dtype=np.float32)
    if key in _AVG:
        return chunk_scores.mean(axis=0)
    if key == "max":
        return chunk_scores.max(axis=0)
    if key in _LME:
        return lme_pooling(chunk_scores, beta=beta)
    raise ValueError(f"Unsupported pooling method: {method}")
