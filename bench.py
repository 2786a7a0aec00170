#!/usr/bin/env python
"""bench.py -- chunk-classification throughput on N B200s (contract: see the task brief).

A *step* = one pass of the hot path over one batch of synthetic multi-chunk files:
PCM16 (24 kHz, 3 s chunks) -> STFT frontend -> int8 DS-CNN (shipped checkpoint graph) ->
LME pooling per file.  Workload = BASELINE.json configs[4] ("file-sharded evaluation of synthetic
multi-chunk files with LME pooling"), the configuration the headline metric (chunks/s at
1/2/4/8 B200) is quoted on; each rank owns `--files` whole files per step (weak scaling, no
collective on the hot path, one all-gather of the pooled scores per step).

  value  : chunks/s with the PCM already resident in HBM (device pointers, CUDA-event timed).
  e2e    : chunks/s through GpuRunner.predict_pooled with pinned HOST buffers (H2D + D2H inside).
  roofline / cpu_baseline / clocks / gpu_launches: see DESIGN.md "Measurement".

`--impl reference` times the CPU oracle (a port of the reference's TFLite + librosa path; the real
reference cannot be installed in this image) on the box's host cores for the same metric.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "birdnet-stm32_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

FIX = os.path.join(ROOT, "tests", "fixtures")
BYTES_PER_CHUNK_ALG = 144_400          # 2*72000 B PCM in + 400 B scores out
MACS_PER_CHUNK = 26_735_616
# both arms print the same workload string (the driver compares `config.workload` of the two lines)
WORKLOAD = ("config5: file-sharded evaluation of synthetic multi-chunk files (chunks/file ~ U{1..20}), "
            "shipped birdnet_stm32n6_100 graph, 3 s / 24 kHz PCM16 chunks, LME pooling beta=10")
# CUDA-core work per chunk that no tensor core or HBM bandwidth removes (SURVEY 8(d) "honest third bound"), counted as
# warp instructions of the shipped kernels' inner loops: 256 real 512-point FFTs + |z| + min/max (K1), 65,792 quantisations
# + 16,384 LUT epilogues (K2), 131,072 stem outputs x 9 taps (K3), 249,856 depthwise outputs, 311,296 pointwise
# requantisations of which 188,416 carry TFLite's three-multiplier residual ADD (K45), MEAN + FC (K6).
CUDA_CORE_WARP_INSTR_PER_CHUNK = 408_000      # ncu smsp__inst_executed.sum over one wave / chunks (profiles/r2/ncu_full_summary.md, commit 7620eb1)


def load_cfg_24k() -> dict:
    cfg = json.load(open(os.path.join(FIX, "birdnet_stm32n6_100_model_config.json")))
    cfg = dict(cfg)
    cfg["sample_rate"] = 24000         # BASELINE synthetic chunks: 3 s @ 24 kHz -> T = 72000, hop = 281
    cfg["hop_length"] = 281
    return cfg


def make_blob(cfg):
    from birdnet_stm32.conversion.export_blob import export_blob

    return export_blob(os.path.join(FIX, "birdnet_stm32n6_100.tflite"), cfg)


def file_layout(n_files: int, seed: int):
    """chunks per file ~ U{1..20} (<= 60 s rule, evaluation/metrics.py:44-46) -> offsets [F+1]."""
    rng = np.random.default_rng(seed)
    counts = rng.integers(1, 21, size=n_files)
    offs = np.zeros(n_files + 1, dtype=np.int32)
    offs[1:] = np.cumsum(counts)
    labels = rng.integers(0, 100, size=n_files)
    return offs, labels


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def synth_device_pcm(torch, n_chunks: int, T: int, sr: int, seed: int, device):
    """Synthetic chirp + noise chunks generated on the device (no host I/O in any timed region)."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    pcm = torch.empty((n_chunks, T), dtype=torch.int16, device=device)
    t = torch.arange(T, device=device, dtype=torch.float32) / sr
    step = 512
    for s in range(0, n_chunks, step):
        n = min(step, n_chunks - s)
        f0 = 300 + (0.45 * sr - 300) * torch.rand((n, 1), generator=g, device=device)
        f1 = 300 + (0.45 * sr - 300) * torch.rand((n, 1), generator=g, device=device)
        amp = 0.1 + 0.7 * torch.rand((n, 1), generator=g, device=device)
        x = amp * torch.sin(2 * np.pi * (f0 * t + (f1 - f0) * t * t / (2 * T / sr)))
        x += (0.05 + 0.25 * torch.rand((n, 1), generator=g, device=device)) * torch.randn((n, T), generator=g, device=device)
        pcm[s:s + n] = torch.round(32767 * x.clamp(-1, 1)).to(torch.int16)
        del x
    return pcm


def file_peaks_device(torch, pcm, offs_np):
    """File-level peak max|pcm/32768| broadcast to the file's chunks (audio/io.py:124-126)."""
    chunk_peak = (pcm.abs().amax(dim=1).to(torch.float32) / 32768.0)
    counts = torch.as_tensor(np.diff(offs_np), device=pcm.device, dtype=torch.long)
    file_id = torch.repeat_interleave(torch.arange(len(counts), device=pcm.device), counts)
    fpk = torch.zeros(len(counts), device=pcm.device).scatter_reduce(0, file_id, chunk_peak, reduce="amax")
    return fpk[file_id].contiguous()


# ---------------------------------------------------------------------------------------------------
def try_real_reference() -> str:
    """use the real TFLite / librosa if this machine has them.  Returns what was found."""
    found = []
    for mod in ("tensorflow", "tflite_runtime", "ai_edge_litert", "librosa"):
        try:
            __import__(mod)
            found.append(mod + " present")
        except Exception:
            found.append(mod + " absent")
    return ", ".join(found)


def cpu_port_throughput(cfg, blob, seconds_target: float, threads: int = 0):
    """Oracle (CPU port of the reference path; built -O3 -march=native on THIS machine) on a bounded sample;
    returns (chunks/s, sample description, threads used)."""
    from birdnet_stm32.audio import synth
    from oracle import bn_oracle

    bn_oracle.use_native_build()
    cores = len(os.sched_getaffinity(0))
    use = threads or cores
    T = int(cfg["sample_rate"] * cfg["chunk_duration"])
    model = bn_oracle.OracleModel(blob, threads=use)

    def run(n):
        pcm = synth.synth_pcm16(n, T, cfg["sample_rate"], seed=99)
        peak = synth.file_peaks(pcm)
        t0 = time.perf_counter()
        spec = bn_oracle.frontend_hybrid(pcm, peak, cfg["fft_length"], T // cfg["spec_width"], cfg["spec_width"], threads=use)
        scores = model.predict(spec)
        bn_oracle.pool_scores(scores, "lme", 10.0)
        return time.perf_counter() - t0

    n0 = max(2 * use, 8)
    dt = run(n0)                                  # calibration
    n = int(max(n0, min(4096, n0 * seconds_target / max(dt, 1e-3))))
    n = (n + use - 1) // use * use
    dt = run(n)
    return n / dt, f"{n} synthetic 3 s / {cfg['sample_rate']} Hz chunks, frontend + int8 graph + LME pooling, {dt:.1f} s, {use} thread(s)", use


def e2e_files_leg(runner, cfg, n_files: int, io_workers: int) -> dict:
    """The reference's real entry point on real files: `evaluate()` (evaluation/metrics.py:75-207 in the reference) over
    synthetic mono 16-bit WAV files of 1..20 chunks, served from the page cache, read by the native C++ reader straight
    into pinned batch buffers, classified and pooled on the device, metric tail included."""
    import shutil
    import tempfile
    import wave

    from birdnet_stm32.audio import synth
    from birdnet_stm32.evaluation.metrics import evaluate

    sr = int(cfg["sample_rate"])
    T = int(sr * cfg["chunk_duration"])
    classes = list(cfg.get("class_names") or [f"c{i}" for i in range(100)])[:100]
    rng = np.random.default_rng(7)
    root = tempfile.mkdtemp(prefix="bn_e2e_files_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    base = synth.synth_pcm16(24, T, sr, seed=5)                   # a pool of chunks; files are random runs of it
    files, n_chunks = [], 0
    try:
        for i in range(n_files):
            k = int(rng.integers(1, 21))
            label = classes[int(rng.integers(0, len(classes)))]
            os.makedirs(os.path.join(root, label), exist_ok=True)
            p = os.path.join(root, label, f"f{i:05d}.wav")
            pcm = base[rng.integers(0, len(base), size=k)].reshape(-1)[: k * T - int(rng.integers(0, T // 2))]
            with wave.open(p, "wb") as wf:
                wf.setnchannels(1)
                wf.setsampwidth(2)
                wf.setframerate(sr)
                wf.writeframes(pcm.tobytes())
            files.append(p)
            n_chunks += k
        ecfg = dict(cfg, class_names=classes)
        kw = dict(pooling="lme", io_workers=io_workers, metrics_backend="device")
        evaluate(runner, files[: min(64, n_files)], classes, ecfg, **kw)      # warm-up (pinned batch buffers, metric kernels)
        # three timed passes, median reported: the leg is host-side work (16 reader threads, page cache) and varies from box to
        # box and run to run far more than anything on the device
        runs = []
        for _ in range(3):
            t0 = time.perf_counter()
            metrics, per_file, _, _ = evaluate(runner, files, classes, ecfg, **kw)
            runs.append(time.perf_counter() - t0)
        dt = sorted(runs)[1]
        return {"value": n_chunks / dt, "unit": "chunks/s", "files": len(per_file), "chunks": n_chunks, "files_per_s": len(per_file) / dt,
                "seconds": dt, "seconds_all_passes": [round(x, 4) for x in runs], "reader": "native (bn_read_pcm16_batch)", "reader_threads": io_workers, "pooling": "lme", "metrics_backend": "device",
                "bytes_read": int(sum(os.path.getsize(f) for f in files)), "skipped_files": metrics.get("skipped_files", 0),
                "what": "evaluate() on synthetic mono PCM16 WAV files in the page cache: read + chunk + H2D + inference + pooling + metrics"}
    finally:
        shutil.rmtree(root, ignore_errors=True)


def h2d_peak_gbs(torch, dev, nbytes: int, dist=None) -> float:
    """Raw pinned host -> device copy rate of THIS box with all ranks copying at once: plain cudaMemcpyAsync of `nbytes`
    per rank, best of 3, max time over ranks.  The ceiling the e2e leg is held against."""
    n = int(min(nbytes, 1 << 30))
    h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    best = None
    for _ in range(4):
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        d.copy_(h, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if dist is not None:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms = float(ms.item())
        best = ms if best is None else min(best, ms)
    return n / (best / 1e3) / 1e9


def emit_line(line: dict) -> None:
    """The ONE JSON line on the real stdout (see the descriptor juggling around NCCL's banner in run_b200)."""
    sys.stdout.flush()
    fd = os.environ.pop("_BN_STDOUT_FD", None)
    if fd is not None:
        os.dup2(int(fd), 1)
        os.close(int(fd))
    print(json.dumps(line), flush=True)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = load_cfg_24k()
    blob = make_blob(cfg)
    per_step = []
    sample = ""
    cores = 0
    total_steps = args.steps + args.warmup
    budget = min(20.0, 120.0 / max(total_steps, 1))
    for i in range(total_steps):
        v, sample, cores = cpu_port_throughput(cfg, blob, seconds_target=budget)
        if i >= args.warmup:
            per_step.append(v)
    value = float(np.mean(per_step))
    one_thread, one_sample, _ = cpu_port_throughput(cfg, blob, seconds_target=8.0, threads=1)
    line = {
        "impl": "reference", "metric": "3s audio chunks/sec (STFT+int8 DS-CNN)", "value": value, "unit": "chunks/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": None,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int8 (fp64 FFT frontend)",
        "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "note": "CPU port (oracle/, gcc -O3 -march=native, OpenMP over chunks) of the reference TFLite + librosa path; "
                           "TensorFlow/librosa are not installable in this image (import attempts: " + try_real_reference() + ")"},
        "cpu_baseline": {"value": value, "unit": "chunks/s", "cores": cores, "kind": "port", "sample": sample,
                         "one_thread": {"value": one_thread, "sample": one_sample}},
        "e2e": {"value": value, "unit": "chunks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch

    from birdnet_stm32 import _lib as L
    from birdnet_stm32.evaluation.gpu_runner import GpuRunner, PinnedArray

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU port")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist_

        dist = dist_
        # stdout carries exactly one JSON line: NCCL prints its "NCCL version ..." banner from C to file descriptor 1 on
        # the first collective, so descriptor 1 points at stderr until the line is printed (emit_line)
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        if rank == 0 and "_BN_STDOUT_FD" not in os.environ:
            sys.stdout.flush()
            os.environ["_BN_STDOUT_FD"] = str(os.dup(1))
            os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)

    cfg = load_cfg_24k()
    blob = make_blob(cfg)
    T = int(cfg["sample_rate"] * cfg["chunk_duration"])
    C = 100
    F = args.files
    offs_np, labels = file_layout(F, seed=2024 + rank)
    n_chunks = int(offs_np[-1])

    runner = GpuRunner(blob, cfg, device=local, wave=args.wave)
    if args.host_wave:
        from birdnet_stm32 import _lib as _L

        runner.set_option(_L.BN_OPT_HOST_WAVE, args.host_wave)
    if args.fusion >= 0:
        runner.set_option(L.BN_OPT_FUSION, args.fusion)
    pcm = synth_device_pcm(torch, n_chunks, T, cfg["sample_rate"], seed=2024 + rank, device=dev)
    peak = file_peaks_device(torch, pcm, offs_np)
    offs = torch.as_tensor(offs_np, device=dev)
    out = torch.empty((F, C), dtype=torch.float32, device=dev)
    gathered = torch.empty((world * F, C), dtype=torch.float32, device=dev) if world > 1 else None
    stream = torch.cuda.current_stream().cuda_stream

    def step():
        runner.infer_pool_ptr(pcm.data_ptr(), peak.data_ptr(), offs.data_ptr(), F, "lme", 10.0, out.data_ptr(), stream)
        if dist is not None:
            dist.all_gather_into_tensor(gathered, out)      # the one collective: pooled scores -> metrics

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()

    # ---- timed region: device-resident inputs --------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = runner.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = runner.launches - l0
    # same K steps again with a CUDA-event pair around every kernel (per-kernel times for the roofline;
    # kept out of the headline timing because ~5k event records per step perturb it by a few percent)
    runner.profile(True, reset=True)
    for _ in range(args.steps):
        step()
    barrier()
    prof = runner.profile_read()
    runner.profile(False)
    clocks = sampler.stop() if rank == 0 else None
    t_ms = torch.tensor([ms], device=dev)
    if dist is not None:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_max = float(t_ms.item())
    total_chunks = torch.tensor([float(n_chunks)], device=dev)
    if dist is not None:
        dist.all_reduce(total_chunks)
    value = float(total_chunks.item()) * args.steps / (ms_max / 1e3)

    # ---- e2e: pinned host buffers through the public API -----------------------------------------
    h_pcm = PinnedArray((n_chunks, T), np.int16)
    h_peak = PinnedArray((n_chunks,), np.float32)
    h_pcm.array[...] = pcm.cpu().numpy()
    h_peak.array[...] = peak.cpu().numpy()
    for _ in range(2):
        host_scores = runner.predict_pooled(h_pcm.array, h_peak.array, offs_np, "lme", 10.0)
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(2, args.steps // 2)
    for _ in range(e2e_steps):
        host_scores = runner.predict_pooled(h_pcm.array, h_peak.array, offs_np, "lme", 10.0)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t_e = torch.tensor([dt], device=dev)
    if dist is not None:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_value = float(total_chunks.item()) * e2e_steps / float(t_e.item())
    same = bool(np.array_equal(host_scores, out.cpu().numpy()))
    h2d = n_chunks * T * 2 + n_chunks * 4 + (F + 1) * 4
    d2h = F * C * 4
    h2d_peak = h2d_peak_gbs(torch, dev, h2d, dist)           # per rank, all ranks copying concurrently
    e2e_gbs_per_rank = e2e_value / world * (T * 2 + 4) / 1e9

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (CUDA events over the timed region) ------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured" if "hbm_gbs" in peaks else "fallback"
    tot_ms = sum(v[0] for v in prof.values()) or 1.0
    dom = max(prof.items(), key=lambda kv: kv[1][0]) if prof else ("none", (1.0, 1))
    dom_ms_per_launch = dom[1][0] / max(dom[1][1], 1)
    # chunks one launch of that kernel processes (every kernel runs once per wave; the last wave of a step is partial)
    chunks_per_launch = n_chunks * args.steps / max(dom[1][1], 1)
    achieved = BYTES_PER_CHUNK_ALG * chunks_per_launch / (dom_ms_per_launch / 1e3) / 1e9
    # DRAM bytes per launch of the same kernel: NOT measured in this run (ncu cannot wrap a timed run).  Taken from the
    # committed `ncu --set full` capture of the same kernel at the bench's wave size (profiles/r2/ncu_traffic.json names
    # the commit it was captured on); null when no capture of this kernel exists.
    traffic = None
    traffic_src = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r2", "ncu_traffic.json")))
        ent = tr.get(dom[0])
        if ent:
            traffic = ent["dram_bytes_per_launch"] / ent["chunks_per_launch"] * chunks_per_launch
            traffic_src = f"ncu --set full capture, commit {tr.get('commit', '?')}, {ent['chunks_per_launch']} chunks per launch (profiles/r2/ncu_traffic.json); scaled by chunks per launch"
    except Exception:
        pass
    sm_clock = (clocks or {}).get("sm_mhz") or 1965.0
    issue_peak = 148 * 4 * sm_clock * 1e6                 # warp instructions / s at 1 per scheduler per clock
    third = {"what": "CUDA-core instruction issue (FP32 FFT, depthwise / stem taps, fixed-point requantisation): warp instructions "
                     "per chunk of the shipped kernels' inner loops / (148 SMs x 4 schedulers x SM clock)",
             "warp_instr_per_chunk": CUDA_CORE_WARP_INSTR_PER_CHUNK, "sm_mhz": sm_clock,
             "bound_chunks_per_s": issue_peak / CUDA_CORE_WARP_INSTR_PER_CHUNK,
             "frac": value / world / (issue_peak / CUDA_CORE_WARP_INSTR_PER_CHUNK)}
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm, "traffic": traffic, "traffic_source": traffic_src,
        "peak_source": peak_src, "kernel": dom[0], "kernel_share_of_step": dom[1][0] / tot_ms,
        "kernel_ms_per_launch": dom_ms_per_launch, "chunks_per_launch": chunks_per_launch,
        "path_achieved_gbs": value / world * BYTES_PER_CHUNK_ALG / 1e9,
        "path_frac": value / world * BYTES_PER_CHUNK_ALG / 1e9 / hbm,
        "int8_tmacs_achieved": value / world * MACS_PER_CHUNK / 1e12,
        "cuda_core_bound": third,
        "kernels_ms": {k: round(v[0] / max(args.steps, 1), 4) for k, v in sorted(prof.items())},
    }

    files_leg = None
    if world == 1 and not args.no_files:
        try:
            files_leg = e2e_files_leg(runner, cfg, args.eval_files, io_workers=min(16, len(os.sched_getaffinity(0))))
        except Exception as exc:                       # the headline numbers do not depend on this leg
            files_leg = {"error": f"{type(exc).__name__}: {exc}"}
    cpu_v, cpu_sample, cpu_cores = cpu_port_throughput(cfg, blob, seconds_target=12.0) if world == 1 and not args.no_cpu else (None, "skipped", 0)
    cpu_1, cpu_1_sample, _ = cpu_port_throughput(cfg, blob, seconds_target=6.0, threads=1) if cpu_v is not None else (None, "", 0)

    line = {
        "metric": "3s audio chunks/sec (STFT+int8 DS-CNN)", "value": value, "unit": "chunks/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int8 (int32 accumulate, fp32 STFT frontend)",
        "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "files_per_gpu_per_step": F, "chunks_per_gpu_per_step": n_chunks, "wave": runner.query().wave,
                   "fast_path": int(runner.query().fast_path),
                   "l2": f"inputs ({n_chunks * T * 2 / 1e9:.2f} GB PCM per step) are larger than L2, no flush needed",
                   "parallelism": f"file-sharded x{world}, one all-gather of pooled scores per step"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "chunks/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "matches_device_run": same,
                "h2d_peak_gbs": h2d_peak, "h2d_achieved_gbs": e2e_gbs_per_rank, "frac": e2e_gbs_per_rank / h2d_peak,
                "h2d_note": "per GPU; peak = plain pinned cudaMemcpyAsync of one step's input bytes on all ranks at once, best of 4"},
        "gpu_launches": int(launches),
        "roofline": roofline,
    }
    if files_leg is not None:
        line["e2e_files"] = files_leg
    if cpu_v is not None:
        line["cpu_baseline"] = {"value": cpu_v, "unit": "chunks/s", "cores": cpu_cores, "kind": "port", "sample": cpu_sample,
                                "one_thread": {"value": cpu_1, "sample": cpu_1_sample},
                                "real_reference": try_real_reference()}
    emit_line(line)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--files", type=int, default=2048, help="files per GPU per step")
    ap.add_argument("--wave", type=int, default=0, help="chunks per engine wave (0 = engine default)")
    ap.add_argument("--host-wave", type=int, default=0, help="chunks per wave for host-memory inputs, e2e leg (0 = engine default)")
    ap.add_argument("--fusion", type=int, default=-1, help="BN_OPT_FUSION bit mask (A/B runs; -1 = engine default)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-files", action="store_true", help="skip the e2e_files leg (evaluate() on WAV files)")
    ap.add_argument("--eval-files", type=int, default=2048, help="files of the e2e_files leg")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
