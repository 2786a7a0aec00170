"""Row f2 measurement: metric tail of BASELINE config 5 (100k files x 100 classes) on the GPU vs scikit-learn on the host.
usage: python scripts/bench_metrics.py [files]  -> one JSON line"""
import json, os, sys, time, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "birdnet-stm32_b200"))
import numpy as np
import torch
from birdnet_stm32.evaluation.device_metrics import metrics_from_scores_device, metrics_from_device_ptrs
from birdnet_stm32.evaluation.metrics import _metrics_from_scores

F = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
C = 100
rng = np.random.default_rng(2024)
y_true = np.zeros((F, C), np.float32); y_true[np.arange(F), rng.integers(0, C, F)] = 1
ys = (rng.random((F, C)) * 0.6).astype(np.float32) + 0.4 * y_true * rng.random((F, C)).astype(np.float32)
ys = (np.round(ys * 4096) / 4096).astype(np.float32)
metrics_from_scores_device(y_true[:1000], ys[:1000])
t0 = time.perf_counter(); got = metrics_from_scores_device(y_true, ys); t_dev_host = time.perf_counter() - t0
d_t, d_s = torch.from_numpy(y_true).cuda(), torch.from_numpy(ys).cuda()
torch.cuda.synchronize()
t0 = time.perf_counter(); res, aps = metrics_from_device_ptrs(d_t.data_ptr(), d_s.data_ptr(), F, C); t_dev = time.perf_counter() - t0
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    t0 = time.perf_counter(); ref = _metrics_from_scores(y_true, ys); t_cpu = time.perf_counter() - t0
diff = max(abs(ref[k] - got[k]) for k in ("roc-auc", "cmAP", "mAP", "f1", "precision", "recall"))
print(json.dumps({"workload": f"{F} files x {C} classes (config 5 metric tail)", "device_resident_s": round(t_dev, 4),
                  "host_arrays_s": round(t_dev_host, 4), "sklearn_host_s": round(t_cpu, 3), "speedup_device_resident": round(t_cpu / t_dev, 1),
                  "max_abs_diff_vs_sklearn": diff, "launches": int(res.n_launches), "cmAP": got["cmAP"], "roc-auc": got["roc-auc"]}))
