"""End-to-end evaluate() on WAV files from the page cache: files/s and chunks/s through the public API, native-rate PCM16
files and files that need the device ingest (48 kHz stereo).  usage: python scripts/bench_evaluate_files.py [files]"""
import json, os, sys, tempfile, time, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "birdnet-stm32_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from birdnet_stm32.audio import io
from birdnet_stm32.evaluation.gpu_runner import GpuRunner
from birdnet_stm32.evaluation.metrics import evaluate
from test_ingest import write_wav

n_files = int(sys.argv[1]) if len(sys.argv) > 1 else 192
FIX = os.path.join(ROOT, "tests", "fixtures")
cfg = json.load(open(os.path.join(FIX, "birdnet_stm32n6_100_model_config.json")))
classes = cfg["class_names"]
rng = np.random.default_rng(0)
root = tempfile.mkdtemp()
native, foreign = [], []
for i in range(n_files):
    d = os.path.join(root, classes[i % len(classes)]); os.makedirs(d, exist_ok=True)
    secs = float(rng.uniform(5, 60))
    p = os.path.join(d, f"n{i}.wav")
    io.save_wav((rng.standard_normal(int(22050 * secs)) * 3000).astype(np.int16), p, 22050); native.append(p)
    if i % 3 == 0:
        q = os.path.join(d, f"s{i}.wav")
        write_wav(q, (rng.standard_normal(int(48000 * secs) * 2) * 3000).astype("<i2"), "s16", 2, 48000); foreign.append(q)
runner = GpuRunner(os.path.join(FIX, "birdnet_stm32n6_100.tflite"), cfg)
out = {"files_native": len(native), "files_48k_stereo": len(foreign), "host_cores": len(os.sched_getaffinity(0))}
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    evaluate(runner, native[:8] + foreign[:4], classes, cfg, pooling="lme")          # warm-up (workspace, filters)
    for name, files in (("native_22k_mono_pcm16", native), ("ingest_48k_stereo", foreign), ("mixed", native + foreign)):
        for w in (1, 8):
            t0 = time.perf_counter()
            m, pf, yt, ys = evaluate(runner, files, classes, cfg, pooling="lme", io_workers=w, measure_latency=True)
            dt = time.perf_counter() - t0
            out[f"{name}_workers{w}"] = {"files_per_s": round(len(files) / dt, 1), "chunks_per_s": round(m["total_chunks"] / dt), "seconds": round(dt, 3)}
print(json.dumps(out))
