"""Row a10 (`load_file_paths_from_directory`, reference `data/dataset.py:49-99`) and row f1's FLAC decoding
(`csrc/bn_flac.h` behind the native reader; the reference reads FLAC through libsndfile, `audio/io.py:90-116`)."""

import os

import numpy as np
import pytest

from flac_enc import encode


# ---------------------------------------------------------------------------------------------------------------
# file discovery
# ---------------------------------------------------------------------------------------------------------------
def _tree(root):
    layout = {"bird_a": ["1.wav", "2.WAV", "3.flac", "notes.txt"], "bird_b": ["x.mp3", "y.ogg"], "noise": ["n.wav"],
              "bird_c/nested": ["deep.wav"], "zz_unlisted": ["u.wav"]}
    for d, names in layout.items():
        os.makedirs(os.path.join(root, d), exist_ok=True)
        for n in names:
            open(os.path.join(root, d, n), "wb").write(b"x")
    return layout


def test_discovery_follows_the_reference_semantics(tmp_path):
    from birdnet_stm32.data.dataset import load_file_paths_from_directory

    root = str(tmp_path)
    _tree(root)
    files, classes = load_file_paths_from_directory(root)
    names = sorted(os.path.relpath(f, root) for f in files)
    # extension filter is case-insensitive; class = the PARENT directory name, so a nested folder is its own class
    assert names == sorted(["bird_a/1.wav", "bird_a/2.WAV", "bird_a/3.flac", "bird_b/x.mp3", "bird_b/y.ogg", "noise/n.wav",
                            "bird_c/nested/deep.wav", "zz_unlisted/u.wav"])
    # noise-like folders are left out of the class list but their files are kept (reference docstring, :66-67)
    assert classes == ["bird_a", "bird_b", "nested", "zz_unlisted"]
    # class restriction
    files2, classes2 = load_file_paths_from_directory(root, classes=["bird_a", "noise"])
    assert sorted(os.path.basename(f) for f in files2) == ["1.wav", "2.WAV", "3.flac", "n.wav"] and classes2 == ["bird_a"]
    # per-class cap: a uniform random subset of the class, drawn from numpy's global RNG
    np.random.seed(1)
    a, _ = load_file_paths_from_directory(root, classes=["bird_a"], max_samples=2)
    np.random.seed(1)
    b, _ = load_file_paths_from_directory(root, classes=["bird_a"], max_samples=2)
    assert a == b and len(a) == 2 and set(a) <= set(files)
    assert len(load_file_paths_from_directory(root, classes=["bird_a"], max_samples=0)[0]) == 3      # 0 / None = no cap
    assert load_file_paths_from_directory(root, classes=["nope"]) == ([], [])
    # custom extension tuple
    only_flac, _ = load_file_paths_from_directory(root, exts=(".flac",))
    assert [os.path.basename(f) for f in only_flac] == ["3.flac"]


def test_discovery_shuffles_with_the_global_rng(tmp_path):
    from birdnet_stm32.data.dataset import load_file_paths_from_directory

    root = str(tmp_path)
    os.makedirs(os.path.join(root, "c"))
    for i in range(40):
        open(os.path.join(root, "c", f"{i:02d}.wav"), "wb").write(b"x")
    np.random.seed(5)
    a, _ = load_file_paths_from_directory(root)
    np.random.seed(5)
    b, _ = load_file_paths_from_directory(root)
    np.random.seed(6)
    c, _ = load_file_paths_from_directory(root)
    assert a == b and a != c and sorted(a) == sorted(c) and a != sorted(a)


def test_cli_discovery_is_identical_on_every_rank(tmp_path, monkeypatch):
    """with --max_files every rank used to subsample with its own unseeded RNG.  Rank 0 discovers and
    broadcasts; here the broadcast is emulated and the other rank's RNG is deliberately different."""
    import torch.distributed as dist

    from birdnet_stm32.cli.evaluate import discover_files

    root = str(tmp_path)
    os.makedirs(os.path.join(root, "c"))
    for i in range(30):
        open(os.path.join(root, "c", f"{i:02d}.wav"), "wb").write(b"x")
    sent = {}

    def fake_broadcast(box, src=0):
        if box[0]:
            sent["files"] = list(box[0])
        else:
            box[0] = list(sent["files"])

    monkeypatch.setattr(dist, "broadcast_object_list", fake_broadcast)
    np.random.seed(123)
    r0 = discover_files(root, ["c"], 10, world=2, rank=0, seed=None)
    np.random.seed(999)
    r1 = discover_files(root, ["c"], 10, world=2, rank=1, seed=None)
    assert r0 == r1 and len(r0) == 10 and r0 == sorted(r0)


# ---------------------------------------------------------------------------------------------------------------
# FLAC
# ---------------------------------------------------------------------------------------------------------------
def _signal(n, ch, bps, kind, rng):
    t = np.arange(n)
    if kind == "tone":
        x = np.stack([0.6 * np.sin(2 * np.pi * (440 + 50 * c) * t / 22050) + 0.05 * rng.standard_normal(n) for c in range(ch)], 1)
    elif kind == "noise":
        x = rng.uniform(-1, 1, (n, ch))
    elif kind == "silence":
        x = np.zeros((n, ch))
    else:                                        # strongly correlated channels, low bits cleared (wasted bits)
        b = 0.5 * np.sin(2 * np.pi * 300 * t / 22050)
        x = np.stack([b + 0.01 * rng.standard_normal(n) for _ in range(ch)], 1)
    v = np.round(np.clip(x, -1, 1) * ((1 << (bps - 1)) - 1)).astype(np.int64)
    return (v >> 3) << 3 if kind == "corr" else v


def _decode(path, max_seconds=0.0):
    from birdnet_stm32.audio.io import read_wav_frames

    raw, kind, ch, sr = read_wav_frames(path, max_seconds)
    return raw.reshape(-1, ch), kind, sr


def test_rfc9639_example_stream(tmp_path):
    """The first decoding example of RFC 9639 (appendix D.1): one frame, two VERBATIM subframes with wasted bits.  The bytes
    carry their own check: the frame header CRC-8 and the frame CRC-16 must both verify for the decoder to accept them."""
    raw = bytes.fromhex("664c6143800000221000100000000f00000f0ac442f0000000013e84b41807dc690307586a3dad1a2e0f"
                        "fff869180000bf0358fd03128baa9a")
    p = tmp_path / "rfc.flac"
    p.write_bytes(raw)
    y, kind, sr = _decode(str(p))
    assert (kind, sr, y.shape) == ("s16", 44100, (1, 2))
    assert y[0, 0] == 25588 and y[0, 1] == 10416
    bad = bytearray(raw)
    bad[-4] ^= 1
    p.write_bytes(bytes(bad))
    from birdnet_stm32.audio.io import UnsupportedAudio

    with pytest.raises(UnsupportedAudio):
        _decode(str(p))


@pytest.mark.parametrize("bps", [8, 12, 16, 20, 24])
def test_flac_streams_of_the_test_encoder_round_trip(tmp_path, bps):
    """Every subframe type / residual method / stereo mode / block-size coding the encoder cycles through (tests/flac_enc.py)."""
    rng = np.random.default_rng(bps)
    n_ok = 0
    for ch in (1, 2, 3):
        for kind in ("tone", "noise", "silence", "corr"):
            for bs, n in ((4096, 30000), (1152, 5000), (255, 1000), (1000, 2500)):
                x = _signal(n, ch, bps, kind, rng)
                p = tmp_path / "t.flac"
                p.write_bytes(encode(x, 22050, bps, bs, first_frame_number=120 if bs == 255 else 0, id3=(ch == 2)))
                y, fmt, sr = _decode(str(p))
                lj = (16 - bps) if bps <= 16 else (32 - bps)
                assert fmt == ("s16" if bps <= 16 else "s32") and sr == 22050 and y.shape == x.shape
                assert np.array_equal(y.astype(np.int64) >> lj, x), (bps, ch, kind, bs)
                assert not np.any(y.astype(np.int64) & ((1 << lj) - 1))
                n_ok += 1
    assert n_ok == 48


def test_flac_window_limit_and_probe(tmp_path):
    from birdnet_stm32.audio import reader as rd

    rng = np.random.default_rng(0)
    x = _signal(50000, 2, 16, "tone", rng)
    p = tmp_path / "w.flac"
    p.write_bytes(encode(x, 22050, 16))
    info = rd.probe(str(p), 1.0)
    assert (info.status, info.container, info.channels, info.sample_rate, info.n_frames) == (rd.RD_NEEDS_INGEST, 1, 2, 22050, 22050)
    y, _, _ = _decode(str(p), 1.0)
    assert y.shape == (22050, 2) and np.array_equal(y, x[:22050])
    (tmp_path / "trunc.flac").write_bytes(p.read_bytes()[:9000])
    assert rd.probe(str(tmp_path / "trunc.flac"), 0.0).n_frames == 50000          # header is intact ...
    n, items = rd.read_raw_batch([str(tmp_path / "trunc.flac")], np.zeros(1 << 20, np.uint8))
    assert n == 1 and items[0] is None                                             # ... the data is not: unreadable, not garbage


def test_flac_copy_of_a_dataset_gives_identical_chunks_and_scores(tmp_path):
    """a FLAC-encoded copy of the WAV test dataset must give bit-identical `y_scores`.  The device path
    is exercised with a runner whose pooled scores are a checksum of the PCM it is handed (GPU-free); mono 16-bit files
    at the model rate take the PCM16 path for both containers, a stereo 48 kHz file goes through the raw-frame reader."""
    import warnings

    from test_evaluate import RAW_CFG, make_dataset

    from birdnet_stm32.audio import reader as rd
    from birdnet_stm32.audio.io import load_pcm16_chunks, read_wav_frames
    from birdnet_stm32.evaluation.metrics import evaluate

    classes = ["bird_a", "bird_b"]
    wav_root, flac_root = tmp_path / "wav", tmp_path / "flac"
    files = make_dataset(str(wav_root), classes, n_per_class=3, sr=22050, seconds=(3.0, 7.4, 1.2))
    flacs = []
    for f in files:
        pcm, _ = __import__("birdnet_stm32.audio.io", fromlist=["read_wav_pcm16"]).read_wav_pcm16(f)
        out = str(flac_root / os.path.relpath(f, str(wav_root))).replace(".wav", ".flac")
        os.makedirs(os.path.dirname(out), exist_ok=True)
        open(out, "wb").write(encode(pcm, 22050, 16))
        flacs.append(out)
        a, pa = load_pcm16_chunks(f, 22050, 3.0)
        b, pb = load_pcm16_chunks(out, 22050, 3.0)
        assert np.array_equal(a, b) and pa == pb

    class Stub:
        device = 0

        def predict_pooled(self, pcm, peak, offs, pooling="avg", beta=10.0):
            out = np.zeros((len(offs) - 1, len(classes)), dtype=np.float32)
            for f in range(len(offs) - 1):
                seg = pcm[offs[f]:offs[f + 1]].astype(np.float64)
                out[f] = [abs(seg.sum()) % 97 / 97.0, float(peak[offs[f]])]
            return out

    cfg = dict(RAW_CFG, audio_frontend="hybrid")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for native in (True, False):
            w = evaluate(Stub(), files, classes, cfg, pooling="avg", device_batch_chunks=5, native_reader=native, io_workers=2)
            fl = evaluate(Stub(), flacs, classes, cfg, pooling="avg", device_batch_chunks=5, native_reader=native, io_workers=2)
            np.testing.assert_array_equal(w[3], fl[3])
            np.testing.assert_array_equal(w[2], fl[2])
    # a file that needs the ingest: raw frames of the FLAC copy == raw frames of the WAV
    from test_ingest import write_wav

    st = np.random.default_rng(1).integers(-9000, 9000, size=(30000, 2)).astype("<i2")
    write_wav(str(tmp_path / "st.wav"), st.reshape(-1), "s16", 2, 48000)
    (tmp_path / "st.flac").write_bytes(encode(st, 48000, 16))
    a, b = read_wav_frames(str(tmp_path / "st.wav"), 60), read_wav_frames(str(tmp_path / "st.flac"), 60)
    assert a[1:] == b[1:] and np.array_equal(a[0], b[0])
    n, items = rd.read_raw_batch([str(tmp_path / "st.flac")], np.zeros(1 << 20, np.uint8))
    assert np.array_equal(items[0][0], a[0])


def test_lossy_containers_are_counted_by_reason_and_strict_mode_raises(tmp_path):
    from test_evaluate import RAW_CFG, make_dataset

    from birdnet_stm32.evaluation.metrics import evaluate

    classes = ["bird_a"]
    files = make_dataset(str(tmp_path), classes, n_per_class=2, sr=22050, seconds=(3.0, 1.0))
    (tmp_path / "bird_a" / "song.mp3").write_bytes(b"ID3\x03\x00\x00\x00\x00\x00\x00" + b"\xff\xfb" * 100)
    (tmp_path / "bird_a" / "broken.wav").write_bytes(b"RIFFxxxxWAVEnope")
    files += [str(tmp_path / "bird_a" / "song.mp3"), str(tmp_path / "bird_a" / "broken.wav")]

    class Stub:
        device = 0

        def predict_pooled(self, pcm, peak, offs, pooling="avg", beta=10.0):
            return np.full((len(offs) - 1, 1), 0.5, np.float32)

    cfg = dict(RAW_CFG, audio_frontend="hybrid")
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m, per_file, _, _ = evaluate(Stub(), files, classes, cfg, pooling="avg")
        assert m["skipped_files"] == 2 and m["skipped_by_reason"] == {"no_decoder": 1, "unreadable": 1} and len(per_file) == 2
        with pytest.raises(RuntimeError, match="could not be decoded"):
            evaluate(Stub(), files, classes, cfg, pooling="avg", strict_files=True)
