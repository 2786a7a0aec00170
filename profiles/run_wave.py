"""One engine wave, repeated: the short command the ncu captures in profiles/ are taken on.

usage: python profiles/run_wave.py [chunks] [repeats] [BN_OPT_FUSION mask]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "birdnet-stm32_b200"))

import numpy as np
import torch

import bench
from birdnet_stm32.evaluation.gpu_runner import GpuRunner

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cfg = bench.load_cfg_24k()
blob = bench.make_blob(cfg)
dev = torch.device("cuda", 0)
T = 72000
pcm = bench.synth_device_pcm(torch, n, T, 24000, 7, dev)
peak = (pcm.abs().amax(dim=1).float() / 32768.0).contiguous()
out = torch.empty((n, 100), dtype=torch.float32, device=dev)
r = GpuRunner(blob, cfg, wave=n)
if len(sys.argv) > 3:
    from birdnet_stm32 import _lib as L
    r.set_option(L.BN_OPT_FUSION, int(sys.argv[3]))
torch.cuda.synchronize()
for _ in range(reps):
    r.infer_pcm16_ptr(pcm.data_ptr(), peak.data_ptr(), n, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
print("launches", r.launches, "top1", out.argmax(dim=1)[:8].tolist())
