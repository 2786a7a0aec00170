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
    from birdnet_stm32.evaluation.metrics import NoValidSamples

    local_metrics: dict = {}
    try:
        local_metrics, per_file, y_true, y_scores = evaluate(model_runner, mine, classes, cfg, **kw)
        labels = y_true.argmax(axis=1).astype(np.int32)
    except NoValidSamples:
        # only "this rank's shard holds no scorable file" is tolerated; engine / CUDA / reader errors propagate so that every
        # rank aborts instead of silently computing the metrics of a partial dataset
        per_file, y_scores, labels = [], np.zeros((0, len(classes)), np.float32), np.zeros((0,), np.int32)
    scores_g, labels_g = gather_scores(y_scores, labels)
    if scores_g.shape[0] == 0:
        raise NoValidSamples("No valid test samples found for the provided class set.")
    y_true_g = np.zeros((labels_g.shape[0], len(classes)), dtype=np.float32)
    y_true_g[np.arange(labels_g.shape[0]), labels_g] = 1.0
    if kw.get("metrics_backend", "sklearn") == "device":
        # the gathered [F, C] matrix goes back to this rank's GPU once: sort / count / accumulate there (bn_metrics_compute)
        from birdnet_stm32.evaluation.device_metrics import metrics_from_scores_device

        metrics = metrics_from_scores_device(y_true_g, scores_g, int(getattr(model_runner, "device", 0)))
    else:
        metrics = _metrics_from_scores(y_true_g, scores_g)
    metrics.update(reduce_run_stats(local_metrics))
    return metrics, per_file, y_true_g, scores_g


def reduce_run_stats(local: dict) -> dict:
    """Skipped-file counts, chunk totals and latency statistics of all ranks: sums for the counters, chunk-weighted mean
    for the mean latency, maximum for the percentiles and the memory figures (an upper bound: the per-chunk samples stay
    on their rank)."""
    import torch
    import torch.distributed as dist

    keys_sum = ["skipped_files", "total_chunks"]
    keys_max = ["latency_median_ms", "latency_p95_ms", "latency_p99_ms", "peak_rss_mb", "rss_delta_mb"]
    reasons = sorted((local.get("skipped_by_reason") or {}).items())
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return {k: local[k] for k in keys_sum + keys_max + ["latency_mean_ms", "skipped_by_reason"] if k in local}
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    chunks = float(local.get("total_chunks", 0))
    vec = [float(local.get(k, 0.0)) for k in keys_sum] + [float(local.get("latency_mean_ms", 0.0)) * chunks]
    t_sum = torch.tensor(vec, dtype=torch.float64, device=dev)
    t_max = torch.tensor([float(local.get(k, 0.0)) for k in keys_max], dtype=torch.float64, device=dev)
    dist.all_reduce(t_sum)
    dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
    gathered: list = [None] * dist.get_world_size()
    dist.all_gather_object(gathered, reasons)
    out: dict = {}
    for k, v in zip(keys_sum, t_sum.tolist()):
        if v:
            out[k] = int(v)
    if out.get("total_chunks") and "latency_mean_ms" in local or t_sum[2].item() > 0:
        tot = t_sum[1].item()
        if tot > 0 and t_sum[2].item() > 0:
            out["latency_mean_ms"] = t_sum[2].item() / tot
    for k, v in zip(keys_max, t_max.tolist()):
        if v:
            out[k] = v
    merged: dict = {}
    for r in gathered:
        for why, n in (r or []):
            merged[why] = merged.get(why, 0) + int(n)
    if merged:
        out["skipped_by_reason"] = merged
    return out


def gather_per_file(per_file: list[dict]) -> list[dict]:
    """Rank-local `per_file` rows of every rank, concatenated in rank order (for the predictions CSV)."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return per_file
    parts: list = [None] * dist.get_world_size()
    dist.all_gather_object(parts, per_file)
    return [row for part in parts for row in (part or [])]
