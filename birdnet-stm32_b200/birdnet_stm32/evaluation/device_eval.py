"""Device path of `evaluate()`: many files per GPU call, in three units.

  reader    `native_batches` / `python_batches`: files -> batches of PCM16 chunks (mono <= 16-bit files at the model rate,
            WAV or FLAC) plus the list of files that need the device ingest (other rates / channels / sample formats).
            The native reader fills pinned buffers with a C++ thread pool, the read of batch k + 1 running under the GPU
            call of batch k.
  batcher   `ForeignBatcher`: the files that need decoding / mixing / resampling are ingested on the device one after
            another into ONE float32 chunk buffer in HBM and classified with ONE `bn_infer_pool_wave_f32` call per batch
            (no per-file launch set, no per-file synchronisation).
  collector `Collector`: pooled rows back into file order, `y_true` / `y_scores` / `per_file`, skipped files by reason.

Reference: the per-file loop of `evaluation/metrics.py:117-147` (read, chunk, `predict` per <= batch_size chunks of one
file, `pool_scores`); results are the same 4-tuple.
"""

from __future__ import annotations

import os
import time
from dataclasses import dataclass, field

import numpy as np

from birdnet_stm32.audio.io import UnsupportedAudio, load_pcm16_chunks, prefetch_ordered, read_wav_frames

LOSSY_EXTS = (".mp3", ".ogg", ".m4a", ".aac", ".opus", ".wma")


def skip_reason(path: str) -> str:
    """Why a file that could not be read is skipped: no decoder for its (lossy) container, or unreadable / empty."""
    return "no_decoder" if path.lower().endswith(LOSSY_EXTS) else "unreadable"


@dataclass
class Batch:
    """Files [start, start + n_files) of the work list."""
    start: int
    n_files: int
    pcm: np.ndarray | None = None                 # int16 [rows, T]: the chunks of the PCM16 files, in file order
    pcm_files: list = field(default_factory=list)  # (index in batch, n_chunks, peak, ok) per file owning rows of `pcm`
    foreign: list = field(default_factory=list)    # (index in batch, loader) -- loader() -> (raw, kind, ch, sr0) or None
    skipped: list = field(default_factory=list)    # (index in batch, reason)


class Collector:
    def __init__(self, classes: list[str]):
        self.classes = classes
        self.index = {c: i for i, c in enumerate(classes)}
        self.y_true: list[np.ndarray] = []
        self.y_scores: list[np.ndarray] = []
        self.per_file: list[dict] = []
        self.skipped: dict[str, int] = {}
        self.skipped_paths: list[tuple[str, str]] = []

    def add(self, path: str, scores: np.ndarray):
        label = os.path.basename(os.path.dirname(path))
        t = np.zeros((len(self.classes),), dtype=np.float32)
        t[self.index[label]] = 1.0
        self.y_true.append(t)
        self.y_scores.append(scores)
        self.per_file.append({"file": path, "label": label, "scores": scores.tolist()})

    def skip(self, path: str, reason: str):
        self.skipped[reason] = self.skipped.get(reason, 0) + 1
        if len(self.skipped_paths) < 50:
            self.skipped_paths.append((path, reason))

    @property
    def n_skipped(self) -> int:
        return sum(self.skipped.values())


class ForeignBatcher:
    """float32 chunks of many files in one device buffer -> one pooled inference call per flush."""

    def __init__(self, runner, sr: int, cd: float, overlap: float, capacity_chunks: int, num_classes: int):
        import torch

        from birdnet_stm32.audio.ingest import GpuIngest, chunk_step

        self.torch = torch
        self.runner = runner
        self.dev = torch.device("cuda", int(getattr(runner, "device", 0)))
        self.T, self.step = chunk_step(sr, cd, overlap)
        self.sr = sr
        self.max_file_chunks = int(60 * sr / self.step) + 8
        self.cap = int(capacity_chunks) + self.max_file_chunks
        self.ingest = GpuIngest(self.dev.index if self.dev.index is not None else 0)
        self.buf = torch.empty((self.cap, self.T), dtype=torch.float32, device=self.dev)
        self.num_classes = num_classes
        self.used = 0
        self.keys: list = []
        self.counts: list[int] = []

    def room(self) -> bool:
        return self.cap - self.used >= self.max_file_chunks

    def add(self, key, raw, kind, ch, sr0) -> int:
        """Ingest one file behind the chunks already in the buffer; returns its chunk count (0 = nothing usable)."""
        n = self.ingest.chunks_to_ptr(raw, kind, ch, sr0, self.sr, self.T, self.step,
                                      self.buf.data_ptr() + 4 * self.used * self.T, self.cap - self.used)
        if n > 0:
            self.used += n
            self.keys.append(key)
            self.counts.append(n)
        return n

    def flush(self, pooling: str, beta: float) -> dict:
        if not self.keys:
            return {}
        torch = self.torch
        offs = np.zeros(len(self.keys) + 1, dtype=np.int32)
        offs[1:] = np.cumsum(self.counts)
        d_offs = torch.from_numpy(offs).to(self.dev)
        d_out = torch.empty((len(self.keys), self.num_classes), dtype=torch.float32, device=self.dev)
        torch.cuda.synchronize(self.dev)               # the ingest kernels of every file of this batch, once
        self.runner.infer_pool_wave_ptr(self.buf.data_ptr(), None, d_offs.data_ptr(), len(self.keys), pooling, beta, d_out.data_ptr(), None)
        rows = dict(zip(self.keys, d_out.cpu().numpy()))
        self.used, self.keys, self.counts = 0, [], []
        return rows

    def close(self):
        self.ingest.close()


def read_foreign(paths: list[str], raw_stage: np.ndarray | None, threads: int) -> list:
    """`read_wav_frames(path, 60)` for every path (None where it fails), read by the native thread pool into `raw_stage`
    group by group; a file that does not fit the stage, or any reader error, falls back to the Python parser.  The
    returned arrays of one call are views into `raw_stage`: consume them before the next call."""
    out: list = [None] * len(paths)
    done = 0
    if raw_stage is not None and paths:
        try:
            from birdnet_stm32.audio import reader as _rd

            n, items = _rd.read_raw_batch(paths, raw_stage, max_seconds=60, threads=max(1, threads))
            out[:n] = items
            done = n
        except Exception:
            done = 0
    for i in range(done, len(paths)):
        try:
            out[i] = read_wav_frames(paths[i], 60)
        except Exception:
            out[i] = None
    return out


# ---------------------------------------------------------------------------------------------------------------
# readers
# ---------------------------------------------------------------------------------------------------------------
def native_batches(todo: list[str], sr: int, T: int, step: int, cap_chunks: int, io_workers: int):
    """Yield `Batch`es from the C++ reader (`bn_read_pcm16_batch`): two pinned stages, batch k + 1 is read while the caller
    works on batch k.  The `pcm` array of a batch is a view into its stage: use it before asking for the next but one."""
    from concurrent.futures import ThreadPoolExecutor

    from birdnet_stm32.audio import reader as _rd

    stages = _pinned_stages(cap_chunks, T)
    window = 2048                                      # paths offered to one reader call
    raw_stage: list = [None]

    def read(start: int, slot: int):
        paths = todo[start:start + window]
        n_files, used, info = _rd.read_pcm16_batch(paths, sr, T, step, stages[slot][1], max_seconds=60, threads=max(1, io_workers))
        return start, n_files, used, info

    def foreign_loader(paths: list[str]):
        """One loader per group of <= 16 files: raw frames are read (natively) when the first of them is asked for."""
        cache: dict = {}

        def load(j: int):
            g0 = j - j % 16
            if cache.get("g0") != g0:
                if raw_stage[0] is None:
                    raw_stage[0] = np.empty(256 << 20, dtype=np.uint8)
                cache["g0"], cache["items"] = g0, read_foreign(paths[g0:g0 + 16], raw_stage[0], io_workers)
            return cache["items"][j - g0]

        return load

    try:
        with ThreadPoolExecutor(max_workers=1, thread_name_prefix="bn-read") as pool:
            fut = pool.submit(read, 0, 0)
            slot = 0
            while fut is not None:
                start, n_files, used, info = fut.result()
                if n_files == 0:
                    raise RuntimeError(f"{todo[start]}: more chunks than the batch buffer holds ({cap_chunks})")
                nxt = start + n_files
                fut = pool.submit(read, nxt, slot ^ 1) if nxt < len(todo) else None
                b = Batch(start, n_files, pcm=stages[slot][1][:used])
                fpaths = []
                for i in range(n_files):
                    fi = info[i]
                    if fi.n_chunks > 0:                # owns rows of the buffer, even if its data read failed afterwards
                        b.pcm_files.append((i, int(fi.n_chunks), float(fi.peak), fi.status == _rd.RD_OK))
                        if fi.status != _rd.RD_OK:
                            b.skipped.append((i, "unreadable"))
                    elif fi.status == _rd.RD_NEEDS_INGEST:
                        fpaths.append((i, todo[start + i]))
                    else:
                        b.skipped.append((i, skip_reason(todo[start + i])))
                if fpaths:
                    load = foreign_loader([p for _, p in fpaths])
                    b.foreign = [(i, (lambda j=j: load(j))) for j, (i, _) in enumerate(fpaths)]
                yield b
                slot ^= 1
    finally:
        pass                                           # the pinned stages stay cached for the next evaluate() call


_STAGE_CACHE: dict = {}


def _pinned_stages(cap_chunks: int, T: int) -> list:
    """Two page-locked int16 [cap, T] batch buffers, kept across `evaluate()` calls: allocating ~1 GB of pinned memory
    costs more than reading and classifying a few thousand chunks.  One geometry is cached at a time."""
    key = (int(cap_chunks), int(T))
    if _STAGE_CACHE.get("key") == key:
        return _STAGE_CACHE["stages"]
    for pa, _ in _STAGE_CACHE.get("stages", []):
        if pa is not None:
            pa.free()
    stages = []
    for _ in range(2):
        try:
            from birdnet_stm32.evaluation.gpu_runner import PinnedArray

            pa = PinnedArray((cap_chunks, T), np.int16)
            stages.append((pa, pa.array))
        except Exception:                              # no CUDA runtime (stub runners in the CPU tests): ordinary memory
            stages.append((None, np.empty((cap_chunks, T), dtype=np.int16)))
    _STAGE_CACHE["key"], _STAGE_CACHE["stages"] = key, stages
    return stages


def python_batches(todo: list[str], sr: int, cd: float, overlap: float, batch_chunks: int, io_workers: int):
    """The same batches from Python reader threads (container parse + chunk cut per file, in order)."""

    def read_one(path: str):
        try:
            pcm, peak = load_pcm16_chunks(path, sr, cd, overlap, max_duration=60)
            return ("pcm", pcm, float(peak)) if pcm.shape[0] else ("skip",)
        except UnsupportedAudio:
            try:
                return ("frames", read_wav_frames(path, 60))
            except Exception:
                return ("skip",)
        except Exception:
            return ("skip",)

    start, items = 0, []

    def make(start, items):
        b = Batch(start, len(items))
        rows = []
        for i, it in enumerate(items):
            if it[0] == "pcm":
                rows.append(it[1])
                b.pcm_files.append((i, it[1].shape[0], it[2], True))
            elif it[0] == "frames":
                b.foreign.append((i, (lambda fr=it[1]: fr)))
            else:
                b.skipped.append((i, skip_reason(todo[start + i])))
        b.pcm = np.concatenate(rows, axis=0) if rows else None
        return b

    n_chunks = 0
    for k, item in enumerate(prefetch_ordered(read_one, todo, workers=io_workers)):
        items.append(item)
        n_chunks += item[1].shape[0] if item[0] == "pcm" else (20 if item[0] == "frames" else 0)
        if n_chunks >= batch_chunks:
            yield make(start, items)
            start, items, n_chunks = k + 1, [], 0
    if items:
        yield make(start, items)


# ---------------------------------------------------------------------------------------------------------------
# the loop
# ---------------------------------------------------------------------------------------------------------------
def run_device_path(runner, todo: list[str], classes: list[str], sr: int, cd: float, overlap: float, pooling: str, beta: float,
                    device_batch_chunks: int, io_workers: int, native_reader: bool, measure_latency: bool):
    """-> (Collector, total_chunks, per-chunk latencies in ms)"""
    from birdnet_stm32.audio.ingest import chunk_step

    T, step = chunk_step(sr, cd, overlap)
    col = Collector(classes)
    latencies: list[float] = []
    total_chunks = 0
    batcher: ForeignBatcher | None = None
    cap = int(device_batch_chunks) + int(60 * sr / step) + 4
    batches = native_batches(todo, sr, T, step, cap, io_workers) if native_reader else \
        python_batches(todo, sr, cd, overlap, int(device_batch_chunks), io_workers)
    try:
        for b in batches:
            t0 = time.perf_counter()
            rows: dict[int, np.ndarray] = {}
            n_batch = 0
            if b.pcm_files:
                counts = np.array([n for _, n, _, _ in b.pcm_files], dtype=np.int64)
                offs = np.zeros(len(counts) + 1, dtype=np.int32)
                offs[1:] = np.cumsum(counts)
                peak = np.repeat(np.array([p for _, _, p, _ in b.pcm_files], dtype=np.float32), counts)
                pooled = runner.predict_pooled(b.pcm, peak, offs, pooling=pooling, beta=beta)
                for (i, n, _, ok), row in zip(b.pcm_files, pooled):
                    if ok:
                        rows[i] = row
                        n_batch += n
            for i, load in b.foreign:
                item = load()
                if item is None:
                    b.skipped.append((i, skip_reason(todo[b.start + i])))
                    continue
                if batcher is None:
                    batcher = ForeignBatcher(runner, sr, cd, overlap, device_batch_chunks, len(classes))
                if not batcher.room():
                    rows.update(batcher.flush(pooling, beta))
                n = batcher.add(i, *item)
                if n == 0:
                    b.skipped.append((i, "unreadable"))
                n_batch += n
            if batcher is not None:
                rows.update(batcher.flush(pooling, beta))
            if measure_latency and n_batch:
                latencies.extend([(time.perf_counter() - t0) * 1000 / n_batch] * n_batch)
            total_chunks += n_batch
            why = dict(b.skipped)
            for i in range(b.n_files):
                path = todo[b.start + i]
                if i in rows:
                    col.add(path, rows[i])
                else:
                    col.skip(path, why.get(i, "unreadable"))
    finally:
        if batcher is not None:
            batcher.close()
    return col, total_chunks, latencies
