"""Row f1 measurement: device ingest (decode + mix + resample_poly + peak + chunks) vs scipy on the host.

usage: python scripts/bench_ingest.py [files]   -> one JSON line (files of 60 s, 48 kHz stereo int16 -> 22.05 kHz chunks)
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "birdnet-stm32_b200"))

import numpy as np
import torch

from birdnet_stm32.audio.ingest import GpuIngest, chunk_step
from oracle import bn_ingest_oracle as O

n_files = int(sys.argv[1]) if len(sys.argv) > 1 else 32
sr_in, sr_out, ch, secs = 48000, 22050, 2, 60.0
n = int(sr_in * secs)
rng = np.random.default_rng(1)
raw = np.round(8000 * rng.standard_normal(n * ch)).clip(-32768, 32767).astype(np.int16)
T, step = chunk_step(sr_out, 3.0, 0.0)
g = GpuIngest(0)
nc = g.num_chunks(g.out_len(n, sr_in, sr_out), T, step)
dev = torch.device("cuda", 0)
buf = torch.empty((nc, T), dtype=torch.float32, device=dev)
pinned = torch.from_numpy(raw).pin_memory().numpy()
d_raw = torch.from_numpy(raw).to(dev)

import ctypes as C
from birdnet_stm32 import _lib as L
lib = L.load()
got = C.c_int()
def run_dev():
    L.check(lib.bn_ingest_chunks(g._h, C.c_void_p(d_raw.data_ptr()), 0, n, ch, sr_in, sr_out, T, step, C.c_void_p(buf.data_ptr()), nc,
                                 C.byref(got), None, None))
def run_host():
    g.chunks_to_ptr(pinned, "s16", ch, sr_in, sr_out, T, step, buf.data_ptr(), nc)

for f in (run_dev, run_host):
    for _ in range(3):
        f()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n_files):
    run_dev()
e1.record()
torch.cuda.synchronize()
ms_dev = e0.elapsed_time(e1) / n_files
t0 = time.perf_counter()
for _ in range(n_files):
    run_host()
torch.cuda.synchronize()
ms_host_in = (time.perf_counter() - t0) * 1000 / n_files

t0 = time.perf_counter()
reps = 2
for _ in range(reps):
    y, _ = O.load_window(raw, "s16", ch, sr_in, sr_out)
    O.split_chunks(y, sr_out, 3.0, 0.0)
ms_cpu = (time.perf_counter() - t0) * 1000 / reps
ref = O.split_chunks(O.load_window(raw, "s16", ch, sr_in, sr_out)[0], sr_out, 3.0, 0.0)
diff = float(np.abs(buf.cpu().numpy() - ref).max())
in_bytes = raw.nbytes
out_bytes = nc * T * 4
print(json.dumps({
    "workload": f"{secs:.0f} s {sr_in} Hz {ch}-channel int16 file -> {nc} float32 chunks of {T} at {sr_out} Hz",
    "gpu_device_resident_ms_per_file": round(ms_dev, 4), "gpu_host_frames_ms_per_file": round(ms_host_in, 4),
    "scipy_numpy_host_ms_per_file": round(ms_cpu, 2), "speedup_device_resident": round(ms_cpu / ms_dev, 1),
    "speedup_host_frames": round(ms_cpu / ms_host_in, 1), "chunks_per_s_device_resident": round(nc / ms_dev * 1000),
    "algorithmic_bytes_per_file": in_bytes + out_bytes, "achieved_gbs_device_resident": round((in_bytes + out_bytes) / ms_dev / 1e6, 1),
    "max_abs_diff_vs_scipy": diff, "launches_per_file": g.launches // (2 * (n_files + 3)),
}))
