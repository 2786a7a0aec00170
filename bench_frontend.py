"""BASELINE config 2: frontend-only sweep on synthetic 3 s chunks, 1 B200 vs host-core numpy.

For every frontend variant -- the hybrid model input (K1 STFT magnitudes + min-max normalise) and the
precomputed-spectrogram features mel + {none, pwl, pcen, db}, log_mel, mfcc (csrc/bn_features.cu) -- this times
`--chunks` device-resident PCM16 chunks with CUDA events on the launching stream (after warm-up; inputs are larger
than L2) and reports chunks/s and the PCM bytes read per second against the measured HBM peak.  The CPU column is
the oracle restatement of the reference's librosa path (`oracle/bn_features_oracle.py`, 1 core, bounded sample).

usage: python bench_frontend.py [--chunks 10000] [--reps 5] [--cpu-chunks 24] [--out profiles/r1/frontend_sweep.json]
"""

from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "birdnet-stm32_b200"))

import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chunks", type=int, default=10000)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--cpu-chunks", type=int, default=24)
    ap.add_argument("--out", default="")
    args = ap.parse_args()

    import torch

    import bench
    from birdnet_stm32.audio import synth
    from birdnet_stm32.audio.spectrogram import FeatureExtractor
    from birdnet_stm32.evaluation.gpu_runner import GpuRunner

    if not torch.cuda.is_available():
        raise SystemExit("bench_frontend.py needs a CUDA device (there is no CPU fallback)")
    dev = torch.device("cuda", 0)
    sr, T, W = 24000, 72000, 256
    n = args.chunks
    pcm = bench.synth_device_pcm(torch, n, T, sr, 11, dev)
    peak = (pcm.abs().amax(dim=1).float() / 32768.0).contiguous()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    st = torch.cuda.current_stream().cuda_stream

    def time_it(fn):
        for _ in range(args.warmup):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.reps

    rows = []
    # hybrid: K1 + normalise into the graph-input layout [B, 257, 256]
    cfg = bench.load_cfg_24k()
    runner = GpuRunner(bench.make_blob(cfg), cfg, wave=2048)
    spec = torch.empty((n, 257, W), dtype=torch.float32, device=dev)
    ms = time_it(lambda: runner.frontend_ptr(pcm.data_ptr(), peak.data_ptr(), n, spec.data_ptr(), st))
    rows.append({"frontend": "hybrid (linear |STFT| + normalize)", "ms": ms})
    del spec
    runner.close()
    variants = [("mel", "none"), ("mel", "pwl"), ("mel", "pcen"), ("mel", "db"), ("log_mel", "none"), ("mfcc", "none")]
    for mode, mag in variants:
        fx = FeatureExtractor(sr, T, 512, 64, W, mag, mode, 20)
        out = torch.empty((n, fx.rows, W), dtype=torch.float32, device=dev)
        ms = time_it(lambda: fx.run_device(pcm.data_ptr(), peak.data_ptr(), n, out.data_ptr(), st))
        rows.append({"frontend": f"librosa-style {mode} + {mag}", "ms": ms})
        fx.close()
        del out
    for r in rows:
        r["chunks_per_s"] = n / (r["ms"] * 1e-3)
        r["pcm_gbs"] = n * T * 2 / (r["ms"] * 1e-3) / 1e9
        r["frac_of_hbm_peak"] = r["pcm_gbs"] / hbm

    # CPU column: numpy/scipy restatement of the reference path, one core, bounded sample
    from oracle import bn_features_oracle as fo

    hp = synth.synth_pcm16(args.cpu_chunks, T, sr, seed=3)
    hk = synth.file_peaks(hp)
    cpu = {}
    t0 = time.perf_counter()
    fo.features_from_pcm16(hp, hk, sample_rate=sr, n_fft=512, mel_bins=-1, spec_width=W)
    cpu["hybrid (linear |STFT| + normalize)"] = args.cpu_chunks / (time.perf_counter() - t0)
    for mode, mag in variants:
        t0 = time.perf_counter()
        fo.features_from_pcm16(hp, hk, sample_rate=sr, n_fft=512, mel_bins=64, spec_width=W, mag_scale=mag, mode=mode, n_mfcc=20)
        cpu[f"librosa-style {mode} + {mag}"] = args.cpu_chunks / (time.perf_counter() - t0)
    for r in rows:
        r["cpu_chunks_per_s_1core"] = cpu[r["frontend"]]

    res = {"config": "BASELINE config 2: frontend-only sweep, synthetic 3 s / 24 kHz PCM16 chunks, device-resident",
           "chunks": n, "reps": args.reps, "hbm_peak_gbs": hbm, "peak_source": "measured" if "hbm_gbs" in peaks else "fallback",
           "cpu": {"kind": "port (numpy/scipy restatement of the librosa path)", "cores": 1, "sample_chunks": args.cpu_chunks},
           "rows": rows}
    line = json.dumps(res)
    print(line)
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        with open(args.out, "w") as fh:
            fh.write(line + "\n")


if __name__ == "__main__":
    main()
