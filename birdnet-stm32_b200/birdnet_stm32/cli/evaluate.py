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
    return p.parse_args(argv)


def main(argv=None):
    args = get_args(argv)
    cfg_path = args.model_config or os.path.splitext(args.model_path)[0] + "_model_config.json"
    if not os.path.isfile(cfg_path):
        raise FileNotFoundError(f"Model config JSON not found: {cfg_path}")
    from birdnet_stm32.data.dataset import SUPPORTED_AUDIO_EXTS, load_file_paths_from_directory
    from birdnet_stm32.evaluation.metrics import evaluate
    from birdnet_stm32.evaluation.sharded import evaluate_sharded
    from birdnet_stm32.models.runners import load_model_runner
    from birdnet_stm32.training.config import ModelConfig

    cfg = ModelConfig.load(cfg_path).to_dict()
    classes = cfg.get("class_names", [])
    if not classes:
        raise ValueError("class_names missing in model config.")
    files, _ = load_file_paths_from_directory(args.data_path_test, classes=classes, exts=SUPPORTED_AUDIO_EXTS, max_samples=args.max_files)
    if not files:
        raise RuntimeError(f"No test audio found in {args.data_path_test}")

    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    device = args.device if args.device is not None else local
    runner = load_model_runner(args.model_path, model_config=cfg, device=device)
    kw = dict(pooling=args.pooling, batch_size=args.batch_size, overlap=max(0.0, min(cfg["chunk_duration"] - 0.1, args.overlap)),
              measure_latency=args.benchmark_latency, profile_memory=args.profile_memory, device_batch_chunks=args.device_batch_chunks,
              metrics_backend=args.metrics_backend)
    if world > 1:
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(device)
        dist.init_process_group("nccl")
        files = sorted(files)      # every rank must see the same order before sharding
        metrics, per_file, y_true, y_scores = evaluate_sharded(runner, files, classes, cfg, **kw)
        is_main = dist.get_rank() == 0
    else:
        metrics, per_file, y_true, y_scores = evaluate(runner, files, classes, cfg, **kw)
        is_main = True

    if is_main:
        print(f"Files evaluated: {y_true.shape[0]}")
        for key in ("roc-auc", "cmAP", "mAP", "f1", "precision", "recall"):
            print(f"{key}: {metrics[key]:.4f}")
        for key in ("latency_mean_ms", "latency_median_ms", "latency_p95_ms", "latency_p99_ms", "total_chunks", "peak_rss_mb", "skipped_files"):
            if key in metrics:
                print(f"{key}: {metrics[key]}")
        if args.save_csv:
            with open(args.save_csv, "w") as fh:
                fh.write("file,label," + ",".join(c.replace(",", " ") for c in classes) + "\n")
                for row in per_file:
                    fh.write(f"{row['file']},{row['label']}," + ",".join(f"{v:.3f}" for v in row["scores"]) + "\n")
        if args.benchmark:
            slim = {k: v for k, v in metrics.items() if k != "ap_per_class"}
            with open(args.benchmark, "w") as fh:
                json.dump({"model": args.model_path, "files": int(y_true.shape[0]), "pooling": args.pooling, "metrics": slim}, fh, indent=2)
    if world > 1:
        import torch.distributed as dist

        dist.destroy_process_group()
    return metrics


if __name__ == "__main__":
    main()
