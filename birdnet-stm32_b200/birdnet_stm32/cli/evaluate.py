"""`python -m birdnet_stm32 evaluate`: the reference CLI (`cli/evaluate.py:29-207`) on the B200 engine.

Flags of the hot path are unchanged (`--model_path --model_config --data_path_test --max_files
--batch_size --overlap --pooling --save_csv --benchmark --benchmark_latency --profile_memory`); new:
`--device` and, under `torchrun`, file-sharded multi-GPU evaluation.  Report writers other than the
predictions CSV / benchmark JSON (HTML, plots, DET, bootstrap CI) are out of scope of this package: they
consume the returned `(metrics, per_file, y_true, y_scores)` unchanged in the reference.
"""

from __future__ import annotations

import argparse
import json
import os


def get_args(argv=None) -> argparse.Namespace:
    p = argparse.ArgumentParser(description="Evaluate model on test audio (file-level pooling) on B200.")
    p.add_argument("--model_path", type=str, required=True, help="Path to .tflite or .b200blob model")
    p.add_argument("--model_config", type=str, default="", help="Path to model config JSON")
    p.add_argument("--data_path_test", type=str, required=True, help="Path to test dataset root")
    p.add_argument("--max_files", type=int, default=-1, help="Max test files per class")
    p.add_argument("--batch_size", type=int, default=16, help="Batch size for the per-file protocol path")
    p.add_argument("--overlap", type=float, default=0.0, help="Chunk overlap (seconds)")
    p.add_argument("--pooling", type=str, default="avg", choices=["avg", "max", "lme"])
    p.add_argument("--save_csv", type=str, default="", help="Optional path to save predictions CSV")
    p.add_argument("--benchmark", type=str, default="", help="Save structured JSON benchmark report to this path")
    p.add_argument("--benchmark_latency", action="store_true", default=False)
    p.add_argument("--profile_memory", action="store_true", default=False)
    p.add_argument("--device", type=int, default=None, help="CUDA device (default: LOCAL_RANK or 0)")
    p.add_argument("--device_batch_chunks", type=int, default=4096, help="Chunks sent to the device per call")
    p.add_argument("--metrics_backend", type=str, default="sklearn", choices=["sklearn", "device"],
                   help="Metric tail: the reference's scikit-learn calls on the host, or the same definitions on the GPU")
    p.add_argument("--strict_files", action="store_true", default=False,
                   help="Fail instead of skipping when a file cannot be decoded (MP3 / OGG / M4A have no decoder here)")
    p.add_argument("--seed", type=int, default=None,
                   help="Seed numpy's global RNG before file discovery (--max_files subsampling and the shuffle use it)")
    return p.parse_args(argv)


def save_predictions_csv(per_file: list[dict], classes: list[str], out_path: str) -> None:
    """Predictions in the reference's column layout (`evaluation/reporting.py:53-78`):
    file, label, top1_label, top1_score, then one score column per class, three decimals."""
    os.makedirs(os.path.dirname(out_path) or ".", exist_ok=True)
    with open(out_path, "w") as fh:
        fh.write(",".join(["file", "label", "top1_label", "top1_score"] + list(classes)) + "\n")
        for row in per_file:
            sc = row["scores"]
            best = max(range(len(sc)), key=sc.__getitem__)
            fh.write(",".join([row["file"], row["label"], classes[best], f"{sc[best]:.3f}"] + [f"{v:.3f}" for v in sc]) + "\n")


def save_benchmark_json(metrics: dict, classes: list[str], model_path: str, out_path: str, config: dict | None = None,
                        species_data: list | None = None) -> None:
    """Benchmark report with the reference's keys (`evaluation/reporting.py:192-236`): model_path, num_classes, num_files
    (the reference stores `total_chunks` there), metrics without the per-class array, floats rounded to 6 places, optional
    species and config."""
    core = {k: (round(v, 6) if isinstance(v, float) else v) for k, v in metrics.items() if k != "ap_per_class"}
    report = {"model_path": model_path, "num_classes": len(classes), "num_files": metrics.get("total_chunks", 0), "metrics": core}
    if species_data:
        report["species"] = species_data
    if config:
        report["config"] = config
    os.makedirs(os.path.dirname(out_path) or ".", exist_ok=True)
    with open(out_path, "w") as fh:
        json.dump(report, fh, indent=2, default=str)


def discover_files(data_path: str, classes: list[str], max_files: int, world: int, rank: int, seed: int | None) -> list[str]:
    """File list of the evaluation.  `load_file_paths_from_directory` subsamples and shuffles with numpy's GLOBAL,
    unseeded RNG (like the reference, `data/dataset.py:88-94`), so under torchrun every rank would draw a different
    subset: rank 0 discovers and broadcasts, then every rank sorts the same list before sharding."""
    import numpy as np

    from birdnet_stm32.data.dataset import SUPPORTED_AUDIO_EXTS, load_file_paths_from_directory

    files: list[str] = []
    if rank == 0:
        if seed is not None:
            np.random.seed(seed)
        files, _ = load_file_paths_from_directory(data_path, classes=classes, exts=SUPPORTED_AUDIO_EXTS, max_samples=max_files)
    if world > 1:
        import torch.distributed as dist

        box = [files]
        dist.broadcast_object_list(box, src=0)
        files = sorted(box[0])
    return files


def main(argv=None):
    args = get_args(argv)
    cfg_path = args.model_config or os.path.splitext(args.model_path)[0] + "_model_config.json"
    if not os.path.isfile(cfg_path):
        raise FileNotFoundError(f"Model config JSON not found: {cfg_path}")
    from birdnet_stm32.evaluation.metrics import evaluate
    from birdnet_stm32.evaluation.sharded import evaluate_sharded, gather_per_file
    from birdnet_stm32.models.runners import load_model_runner
    from birdnet_stm32.training.config import ModelConfig

    cfg = ModelConfig.load(cfg_path).to_dict()
    classes = cfg.get("class_names", [])
    if not classes:
        raise ValueError("class_names missing in model config.")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    device = args.device if args.device is not None else local
    rank = 0
    if world > 1:
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(device)
        dist.init_process_group("nccl")
        rank = dist.get_rank()
    files = discover_files(args.data_path_test, classes, args.max_files, world, rank, args.seed)
    if not files:
        raise RuntimeError(f"No test audio found in {args.data_path_test}")

    runner = load_model_runner(args.model_path, model_config=cfg, device=device)
    kw = dict(pooling=args.pooling, batch_size=args.batch_size, overlap=max(0.0, min(cfg["chunk_duration"] - 0.1, args.overlap)),
              measure_latency=args.benchmark_latency or bool(args.benchmark), profile_memory=args.profile_memory,
              device_batch_chunks=args.device_batch_chunks, metrics_backend=args.metrics_backend, strict_files=args.strict_files)
    if world > 1:
        metrics, per_file, y_true, y_scores = evaluate_sharded(runner, files, classes, cfg, **kw)
        if args.save_csv:
            per_file = gather_per_file(per_file)       # every rank's rows, not just rank 0's shard
    else:
        metrics, per_file, y_true, y_scores = evaluate(runner, files, classes, cfg, **kw)
    is_main = rank == 0

    if is_main:
        print(f"Files evaluated: {y_true.shape[0]}")
        for key in ("roc-auc", "cmAP", "mAP", "f1", "precision", "recall"):
            print(f"{key}: {metrics[key]:.4f}")
        for key in ("latency_mean_ms", "latency_median_ms", "latency_p95_ms", "latency_p99_ms", "total_chunks", "peak_rss_mb",
                    "skipped_files", "skipped_by_reason"):
            if key in metrics:
                print(f"{key}: {metrics[key]}")
        if args.save_csv:
            save_predictions_csv(per_file, classes, args.save_csv)
            print(f"Predictions saved to {args.save_csv}")
        if args.benchmark:
            save_benchmark_json(metrics, classes, args.model_path, args.benchmark, config=cfg)
            print(f"Benchmark report saved to {args.benchmark}")
    if world > 1:
        import torch.distributed as dist

        dist.destroy_process_group()
    return metrics


if __name__ == "__main__":
    main()
