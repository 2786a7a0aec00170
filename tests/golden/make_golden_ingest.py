"""Generates tests/golden/ingest_reference.npz.  Run ONCE in the build container (where /root/reference is
mounted); the tests only read the committed output.

The REAL `load_audio_window` and `split_audio_into_chunks` of /root/reference/birdnet_stm32/audio/io.py are run
on small synthetic WAV files.  `soundfile` is not installed here, so the module is imported with a stub
`soundfile` that only parses the RIFF container and converts samples to float32 the way libsndfile does (the
conversion list is in oracle/bn_ingest_oracle.py); the channel mean, `fast_resample` ->
`scipy.signal.resample_poly`, the peak normalisation and the chunking are the reference's own code and the real
scipy.  Stored per case: the raw interleaved samples, format, channels, rates, max_duration, the window the
reference returned and the chunks it cut.
"""

import importlib.util
import os
import struct
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/birdnet_stm32"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "birdnet-stm32_b200"))

from birdnet_stm32.audio.io import read_wav_frames  # container parser only
from oracle import bn_ingest_oracle as O


def write_wav(path, raw, kind, ch, sr):
    tag, bits = {"u8": (1, 8), "s16": (1, 16), "s24": (1, 24), "s32": (1, 32), "f32": (3, 32)}[kind]
    data = np.ascontiguousarray(raw).tobytes()
    bps = bits // 8
    with open(path, "wb") as fh:
        fh.write(b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVE")
        fh.write(b"fmt " + struct.pack("<IHHIIHH", 16, tag, ch, sr, sr * ch * bps, ch * bps, bits))
        fh.write(b"data" + struct.pack("<I", len(data)) + data)


def make_stub():
    sf = types.ModuleType("soundfile")

    class Info:
        def __init__(self, path):
            raw, kind, ch, sr = read_wav_frames(path, None)
            self.samplerate = sr
            self.frames = raw.size // ((3 if kind == "s24" else 1) * ch)

    class SoundFile:
        def __init__(self, path, mode="r"):
            self.raw, self.kind, self.ch, self.sr = read_wav_frames(path, None)
            self.pos = 0

        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

        def seek(self, frame):
            self.pos = int(frame)

        def read(self, frames, dtype="float32", always_2d=True):
            assert dtype == "float32" and always_2d
            x = O.decode(self.raw, self.kind, self.ch)
            return x[self.pos : self.pos + int(frames)]

    sf.info = Info
    sf.SoundFile = SoundFile
    return sf


def synth(rng, n, ch, sr):
    t = np.arange(n) / sr
    cols = []
    for c in range(ch):
        f0, f1 = rng.uniform(300, 0.45 * sr, size=2)
        ph = 2 * np.pi * (f0 * t + 0.5 * (f1 - f0) * t * t / max(t[-1], 1e-9))
        cols.append(rng.uniform(0.2, 0.7) * np.sin(ph) + 0.05 * rng.standard_normal(n))
    return np.clip(np.stack(cols, axis=1), -1.0, 1.0)


def encode(x, kind):
    if kind == "s16":
        return np.round(32767 * x).astype("<i2").reshape(-1)
    if kind == "s32":
        return np.round(2147483000 * x).astype("<i4").reshape(-1)
    if kind == "f32":
        return x.astype("<f4").reshape(-1)
    if kind == "u8":
        return (np.round(127 * x) + 128).astype(np.uint8).reshape(-1)
    v = np.round(8388607 * x).astype(np.int32).reshape(-1)
    b = np.empty((v.size, 3), dtype=np.uint8)
    b[:, 0] = v & 0xFF
    b[:, 1] = (v >> 8) & 0xFF
    b[:, 2] = (v >> 16) & 0xFF
    return b.reshape(-1)


def main():
    sys.modules["soundfile"] = make_stub()
    spec = importlib.util.spec_from_file_location("ref_audio_io", os.path.join(REF, "audio/io.py"))
    io = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(io)

    rng = np.random.default_rng(20261018)
    # (sr_in, sr_out, channels, format, seconds, chunk_duration, overlap, max_duration)
    cases = [
        (48000, 22050, 2, "s16", 0.40, 0.1, 0.0, 30),
        (44100, 22050, 1, "s16", 0.35, 0.15, 0.03, 30),
        (44100, 24000, 1, "f32", 0.30, 0.1, 0.0, 30),
        (32000, 24000, 3, "s24", 0.25, 0.15, 0.2, 30),         # overlap clamped to cd - 0.1
        (22050, 22050, 2, "s16", 0.30, 0.1, 0.0, 30),
        (16000, 22050, 1, "u8", 0.30, 0.1, 0.0, 30),
        (48000, 24000, 1, "s32", 0.25, 0.1, 0.0, 0.2),       # max_duration cuts the read
        (48000, 22050, 1, "s16", 0.05, 0.1, 0.0, 30),        # shorter than one chunk -> zero padded
        (24000, 24000, 1, "s24", 0.22, 0.1, 0.0, 30),
    ]
    out = {"n_cases": np.int64(len(cases))}
    with tempfile.TemporaryDirectory() as td:
        for i, (sr_in, sr_out, ch, kind, secs, cd, ov, maxd) in enumerate(cases):
            x = synth(rng, int(secs * sr_in), ch, sr_in)
            raw = encode(x, kind)
            path = os.path.join(td, f"c{i}.wav")
            write_wav(path, raw, kind, ch, sr_in)
            y = io.load_audio_window(path, sample_rate=sr_out, max_duration=maxd, chunk_duration=cd)
            chunks = io.split_audio_into_chunks(y, sample_rate=sr_out, chunk_duration=cd, chunk_overlap=ov)
            assert y.dtype == np.float32 and y.size > 0
            out[f"raw_{i}"] = raw
            out[f"meta_{i}"] = np.array([sr_in, sr_out, ch, ["s16", "s24", "s32", "f32", "u8"].index(kind)], dtype=np.int64)
            out[f"par_{i}"] = np.array([cd, ov, maxd], dtype=np.float64)
            out[f"window_{i}"] = y
            out[f"chunks_{i}"] = chunks
            print(i, kind, ch, sr_in, "->", sr_out, "window", y.shape, "chunks", chunks.shape)
    np.savez_compressed(os.path.join(HERE, "ingest_reference.npz"), **out)
    print("bytes:", os.path.getsize(os.path.join(HERE, "ingest_reference.npz")))


if __name__ == "__main__":
    main()
