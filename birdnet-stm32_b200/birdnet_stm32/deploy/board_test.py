"""Board-test report reader and host comparison (reference `deploy/board_test.py:406-512`).

The reference's board test flashes a firmware that classifies the FIRST chunk of every `.wav` on the SD card
(`firmware/Src/main.c:304-355`) and prints, per file, the top-k classes over UART:

    [3/10] song_sparrow_01.wav
      [WAV] 22050 Hz, 16-bit, 1 ch, 66150 samples
      [BENCH] read=12ms stft=48ms npu=9ms total=69ms
      song_sparrow_01.wav:
        [1] Melospiza melodia_Song Sparrow: 87.3%
        [2] Passerella iliaca_Fox Sparrow: 4.1%
    === DONE ===
    Processed: 10 / 10 files (0 errors)
    Benchmark: ... (avg read=12ms stft=48ms npu=9ms total=69ms)

`parse_serial_output` turns those lines into the same dictionary the reference returns (keys `results`, `processed`,
`total`, `errors`, `benchmark`, `raw_lines`; per file `file`, `detections` [{label, score}], `bench`).  What the reference
leaves to the reader's eye -- is the board right? -- is `compare_with_engine`: the same first chunks go through the B200
engine (`GpuRunner.predict_pcm16`) and every board detection is matched against the engine's score for that class.
The board computes its STFT itself (symmetric Hann window, un-centred frames, `firmware/Src/audio_stft.c:21,41-48`) and
runs the NPU's own int8 kernels, so agreement is a tolerance, not an identity.
"""

from __future__ import annotations

import os
import re

import numpy as np

DONE_MARKER = "=== DONE ==="

_FILE = re.compile(r"^\[(\d+)/(\d+)\]\s+(.+)$")
_DETECTION = re.compile(r"^\s+\[(\d+)\]\s+(.+?):\s+([\d.]+)%$")
_PROCESSED = re.compile(r"^Processed:\s+(\d+)\s*/\s*(\d+)\s+files\s+\((\d+)\s+errors\)")
_BENCH = re.compile(r"^\s+\[BENCH\]\s+read=(\d+)ms\s+stft=(\d+)ms\s+npu=(\d+)ms\s+total=(\d+)ms$")
_BENCH_AVG = re.compile(r"^Benchmark:.*avg read=(\d+)ms\s+stft=(\d+)ms\s+npu=(\d+)ms\s+total=(\d+)ms\)$")


def parse_serial_output(lines: list[str], labels: list[str], top_k: int = 5, threshold: float = 0.01) -> dict:
    """Firmware UART lines -> structured report (same keys and filtering as the reference: at most `top_k` detections per
    file, scores as fractions, detections below `threshold` dropped)."""
    files: list[dict] = []
    report = {"results": files, "processed": 0, "total": 0, "errors": 0, "benchmark": None, "raw_lines": lines}
    for line in lines:
        hit = _FILE.match(line)
        if hit:
            files.append({"file": hit.group(3), "detections": [], "bench": None})
            continue
        hit = _DETECTION.match(line)
        if hit:
            if files:
                score = float(hit.group(3)) / 100.0
                if score >= threshold and len(files[-1]["detections"]) < top_k:
                    files[-1]["detections"].append({"label": hit.group(2), "score": score})
            continue
        hit = _BENCH.match(line)
        if hit:
            if files:
                files[-1]["bench"] = dict(zip(("read_ms", "stft_ms", "npu_ms", "total_ms"), map(int, hit.groups())))
            continue
        hit = _BENCH_AVG.match(line)
        if hit:
            report["benchmark"] = dict(zip(("avg_read_ms", "avg_stft_ms", "avg_npu_ms", "avg_total_ms"), map(int, hit.groups())))
            continue
        hit = _PROCESSED.match(line)
        if hit:
            report["processed"], report["total"], report["errors"] = (int(v) for v in hit.groups())
    return report


def first_chunk_pcm16(path: str, sample_rate: int, chunk_len: int) -> np.ndarray | None:
    """What the firmware reads: the first `chunk_len` samples of a mono 16-bit WAV at the model rate, zero-padded
    (`firmware/Src/main.c:296-306`); None for files the firmware skips (other rate / format)."""
    from birdnet_stm32.audio.io import UnsupportedAudio, read_wav_pcm16

    try:
        pcm, sr = read_wav_pcm16(path, chunk_len)
    except (UnsupportedAudio, OSError):
        return None
    if sr != sample_rate:
        return None
    out = np.zeros((chunk_len,), dtype=np.int16)
    out[: pcm.size] = pcm[:chunk_len]
    return out


def compare_with_engine(report: dict, audio_dir: str, runner, labels: list[str], cfg: dict, score_tolerance: float = 0.15) -> dict:
    """Match the board's detections against the B200 engine on the same first chunks.

    Returns per-file rows (`top1_board`, `top1_engine`, `top1_match`, `max_abs_delta` over the board's listed classes) and
    the aggregates `files_compared`, `top1_agreement`, `detections_within_tolerance`, `mean_abs_delta`.
    Files the log lists but the host cannot find / decode are reported under `missing`.
    The engine sees the raw int16 samples without peak normalisation, like the firmware (`wav_read_chunk_f32` scales by
    1/32768 only)."""
    sr = int(cfg["sample_rate"])
    T = int(sr * float(cfg["chunk_duration"]))
    index = {name: i for i, name in enumerate(labels)}
    rows, missing, pcm_rows, row_files = [], [], [], []
    for entry in report["results"]:
        path = os.path.join(audio_dir, entry["file"])
        pcm = first_chunk_pcm16(path, sr, T) if os.path.exists(path) else None
        if pcm is None:
            missing.append(entry["file"])
            continue
        pcm_rows.append(pcm)
        row_files.append(entry)
    if pcm_rows:
        scores = runner.predict_pcm16(np.stack(pcm_rows), None)
    n_det = n_ok = 0
    deltas: list[float] = []
    for entry, s in zip(row_files, scores if pcm_rows else []):
        det = entry["detections"]
        top_engine = labels[int(np.argmax(s))]
        top_board = det[0]["label"] if det else None
        worst = 0.0
        for d in det:
            if d["label"] not in index:
                continue
            delta = abs(float(s[index[d["label"]]]) - d["score"])
            deltas.append(delta)
            worst = max(worst, delta)
            n_det += 1
            n_ok += delta <= score_tolerance
        rows.append({"file": entry["file"], "top1_board": top_board, "top1_engine": top_engine, "top1_match": top_board == top_engine,
                     "max_abs_delta": worst, "engine_top1_score": float(np.max(s))})
    with_det = [r for r in rows if r["top1_board"] is not None]
    return {
        "files": rows, "missing": missing, "files_compared": len(rows),
        "top1_agreement": (sum(r["top1_match"] for r in with_det) / len(with_det)) if with_det else float("nan"),
        "detections_within_tolerance": (n_ok / n_det) if n_det else float("nan"),
        "mean_abs_delta": float(np.mean(deltas)) if deltas else float("nan"),
        "score_tolerance": score_tolerance,
    }
