"""BASELINE configs 3 and 4 as throughput cases: chunks/s of the synthesised graphs on one B200.

  config 3  raw-waveform frontend DS-CNN (learned filterbank conv, T = 48,000), random-init, own PTQ per channel
  config 4  wide DS-CNN alpha 1.0 / depth multiplier 2 with SE + attention pooling, plain-DS and inverted-residual forms,
            per-channel vs per-tensor int8 weights

The graphs are produced by this repo's TensorFlow-free builder + PTQ (`birdnet_stm32/conversion/ptq.py`; the reference needs
TensorFlow for `convert_to_tflite`, conversion/quantize.py:111-168) and run on the engine's one-kernel-per-op plan (these
topologies are outside the fused plan of the shipped checkpoint: SE gates, inverted residuals and attention pooling have no
fused kernels yet), bit-exact against the oracle (tests/test_ptq.py).  Timed with CUDA events on device-resident model inputs
(`bn_infer_spec_f32` entry); int8 MACs per chunk are counted from the op list for the tensor-pipe figure.

usage: python bench_configs.py [--chunks 2048] [--reps 5] [--out profiles/r2/configs34.json]
"""

from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "birdnet-stm32_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np


def macs_per_chunk(g) -> tuple[int, int]:
    """(all int8 MACs, MACs in 1x1 conv / FC = tensor-core eligible) from the parsed graph."""
    total = pw = 0
    for op in g.ops:
        if op.kind not in ("CONV_2D", "DEPTHWISE_CONV_2D", "FULLY_CONNECTED"):
            continue
        w = g.tensor(op.inputs[1])
        y = g.tensor(op.outputs[0])
        out_elems = int(np.prod(y.shape[1:]))
        if op.kind == "DEPTHWISE_CONV_2D":
            m = out_elems * int(w.shape[1] * w.shape[2])
        elif op.kind == "CONV_2D":
            m = out_elems * int(w.shape[1] * w.shape[2] * w.shape[3])
            if w.shape[1] == 1 and w.shape[2] == 1:
                pw += m
        else:
            m = out_elems * int(w.shape[1])
            pw += m
        total += m
    return total, pw


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chunks", type=int, default=2048)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--out", default="")
    args = ap.parse_args()

    import torch

    from birdnet_stm32.conversion import ptq
    from birdnet_stm32.evaluation.gpu_runner import GpuRunner
    from test_ptq import CASES, _case

    if not torch.cuda.is_available():
        raise SystemExit("bench_configs.py needs a CUDA device (there is no CPU fallback)")
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    int8_peak = 2.0 * float(peaks.get("bf16_tflops_sustained", 1400.0)) / 2.0      # TMAC/s: int8 rate = 2 x bf16, 2 ops per MAC
    rows = {}
    for name in CASES:
        fg, _, g, blob = _case(name)
        _, cfg, per_channel = CASES[name]
        x1 = ptq.synth_calibration(fg, 16, seed=5).astype(np.float32)
        reps = (args.chunks + 15) // 16
        x = torch.from_numpy(np.tile(x1, (reps,) + (1,) * (x1.ndim - 1))[: args.chunks]).cuda().contiguous()
        n = x.shape[0]
        r = GpuRunner(blob, cfg, wave=min(n, 2048))
        out = torch.empty((n, r.num_classes), dtype=torch.float32, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        for _ in range(args.warmup):
            r.infer_spec_ptr(x.data_ptr(), n, out.data_ptr(), st)
        torch.cuda.synchronize()
        l0 = r.launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            r.infer_spec_ptr(x.data_ptr(), n, out.data_ptr(), st)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.reps
        macs, pw = macs_per_chunk(g)
        cps = n / (ms / 1e3)
        rows[name] = {"chunks_per_s": cps, "ms_per_pass": ms, "chunks": n, "launches_per_pass": (r.launches - l0) // args.reps,
                      "fast_path": int(r.query().fast_path), "per_channel": bool(per_channel), "ops": len(g.ops),
                      "int8_macs_per_chunk": macs, "pointwise_macs_per_chunk": pw, "tmacs_achieved": cps * macs / 1e12,
                      "int8_tensor_pipe_frac_of_2x_bf16_sustained": cps * pw / 1e12 / int8_peak,
                      "input_shape": list(x1.shape[1:])}
        r.close()
        print(name, json.dumps(rows[name]), flush=True)
    if args.out:
        os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
        json.dump({"what": "BASELINE configs 3 / 4, one B200, generic (one kernel per op) plan, device-resident inputs, CUDA events",
                   "int8_peak_tmacs": int8_peak, "cases": rows}, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
