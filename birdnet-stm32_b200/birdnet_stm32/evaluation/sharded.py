"""File-sharded multi-GPU evaluation: one process per GPU, no collective on the hot
path, ONE all-gather of the pooled per-file scores (+ label ids) at the end; rank 0 computes the metrics.

The reference has no distributed path at all (`evaluation/metrics.py:117-147` is a serial loop); this
is the new scale-out wrapper around `evaluate()`.  It is backend agnostic: `nccl` over NVLink on the
B200 box, `gloo` in the CPU tests.
"""

from __future__ import annotations

import os

import numpy as np


def shard_files(files: list[str], world_size: int, rank: int, weights: list[int] | None = None) -> list[str]:
    """Deal whole files to ranks so pooling stays local.  With `weights` (chunks per file) the files are
    assigned greedily, heaviest first, to the least loaded rank; otherwise round-robin."""
    if world_size <= 1:
        return list(files)
    if weights is None:
        return list(files[rank::world_size])
    order = sorted(range(len(files)), key=lambda i: (-weights[i], i))
    load = [0] * world_size
    mine = []
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        load[r] += weights[i]
        if r == rank:
            mine.append(i)
    return [files[i] for i in sorted(mine)]


def gather_scores(local_scores: np.ndarray, local_labels: np.ndarray, device=None):
    """All-gather `[F_r, C]` float32 scores and `[F_r]` int32 labels from every rank (padded to the
    largest shard); returns the concatenation in rank order on every rank."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local_scores, local_labels
    world = dist.get_world_size()
    dev = device if device is not None else (torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu"))
    C = int(local_scores.shape[1])
    n = torch.tensor([local_scores.shape[0]], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n)
    counts = [int(c.item()) for c in counts]
    fmax = max(max(counts), 1)
    buf = torch.zeros((fmax, C + 1), dtype=torch.float32, device=dev)
    if local_scores.shape[0]:
        buf[: local_scores.shape[0], :C] = torch.from_numpy(np.ascontiguousarray(local_scores, dtype=np.float32)).to(dev)
        buf[: local_scores.shape[0], C] = torch.from_numpy(local_labels.astype(np.float32)).to(dev)
    out = torch.empty((world * fmax, C + 1), dtype=torch.float32, device=dev)
    dist.all_gather_into_tensor(out, buf)
    out = out.cpu().numpy().reshape(world, fmax, C + 1)
    scores = np.concatenate([out[r, : counts[r], :C] for r in range(world)], axis=0)
    labels = np.concatenate([out[r, : counts[r], C] for r in range(world)], axis=0).astype(np.int32)
    return scores, labels


def evaluate_sharded(model_runner, files: list[str], classes: list[str], cfg: dict, **kw):
    """`evaluate()` over this rank's shard, then one all-gather; every rank returns the global result.

    Returns `(metrics, per_file_local, y_true_global, y_scores_global)`; metrics are computed from the
    gathered arrays with the same code as the single-process path.
    """
    import torch.distributed as dist

    from birdnet_stm32.evaluation.metrics import _metrics_from_scores, evaluate

    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    mine = shard_files(files, world, rank)
    try:
        _, per_file, y_true, y_scores = evaluate(model_runner, mine, classes, cfg, **kw)
        labels = y_true.argmax(axis=1).astype(np.int32)
    except RuntimeError:
        per_file, y_scores, labels = [], np.zeros((0, len(classes)), np.float32), np.zeros((0,), np.int32)
    scores_g, labels_g = gather_scores(y_scores, labels)
    if scores_g.shape[0] == 0:
        raise RuntimeError("No valid test samples found for the provided class set.")
    y_true_g = np.zeros((labels_g.shape[0], len(classes)), dtype=np.float32)
    y_true_g[np.arange(labels_g.shape[0]), labels_g] = 1.0
    if kw.get("metrics_backend", "sklearn") == "device":
        # the gathered [F, C] matrix goes back to this rank's GPU once: sort / count / accumulate there (bn_metrics_compute)
        from birdnet_stm32.evaluation.device_metrics import metrics_from_scores_device

        return metrics_from_scores_device(y_true_g, scores_g, int(getattr(model_runner, "device", 0))), per_file, y_true_g, scores_g
    return _metrics_from_scores(y_true_g, scores_g), per_file, y_true_g, scores_g
