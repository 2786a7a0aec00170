"""GPU parity at the configuration `bench.py` measures.

The shipped graph at T = 72,000 (3 s / 24 kHz), hop 281, through the frame-major K1 + tensor-core head path, on LARGE
waves: the device-resident wave (>= 4096 chunks in one launch set) and the 592-chunk host waves, against the OpenMP CPU
oracle (`oracle/bn_oracle.c`, running on requantisation constants it derived itself, `oracle/tflite_quant.py`).

Bars (BASELINE.json north_star; written out here):
  * frontend: spectrogram error <= 1e-4 of the chunk maximum, >= 99.9 % identical int8 input codes;
  * int8 body on the ORACLE's spectrograms: scores bit-exact (every chunk);
  * full path PCM16 -> scores: |delta| <= 1 LSB (1/256); top-1 equal wherever the oracle's top-1 margin exceeds 1 LSB
    (a 1-LSB tie cannot be arbitrated by a float frontend that is allowed 0.1 % differing input codes);
  * pooled file scores: equal to pooling the oracle's chunk scores within 1 LSB.
SURVEY 8(d) config 1 sweep B in {1, 16, 256, 4096} and the section-7 gate "identical to the oracle for >= 10 k chunks".
"""

import json

import numpy as np
import pytest

from conftest import CONFIG, TFLITE

pytestmark = pytest.mark.gpu

T24, SR24, HOP24, W = 72000, 24000, 281, 256
LSB = 1.0 / 256


@pytest.fixture(scope="module")
def cfg24():
    cfg = dict(json.load(open(CONFIG)))
    cfg["sample_rate"] = SR24
    cfg["hop_length"] = HOP24
    return cfg


@pytest.fixture(scope="module")
def blob24(cfg24):
    from birdnet_stm32.conversion.export_blob import export_blob

    return export_blob(TFLITE, cfg24)


@pytest.fixture(scope="module")
def runner24(blob24, cfg24):
    from birdnet_stm32.evaluation.gpu_runner import GpuRunner

    r = GpuRunner(blob24, cfg24)
    assert r.query().fast_path == 1 and r.info.chunk_len == T24
    yield r
    r.close()


@pytest.fixture(scope="module")
def oracle24(blob24):
    from oracle import bn_oracle, tflite_quant

    bn_oracle.build()
    return bn_oracle.OracleModel(tflite_quant.patch_blob(blob24, tflite_quant.derive(TFLITE)))


def device_chunks(n: int, seed: int):
    """n synthetic 3 s / 24 kHz chunks (chirp + noise, generated on the GPU like bench.py's) with the SURVEY 8(d) edge
    cases in the last four rows; returns (int16 [n, T] host array, float32 [n] peaks)."""
    import torch

    from bench import synth_device_pcm
    from birdnet_stm32.audio import synth

    pcm = synth_device_pcm(torch, n, T24, SR24, seed=seed, device=torch.device("cuda", 0)).cpu().numpy()
    if n >= 8:
        pcm[-4:] = synth.synth_pcm16(8, T24, SR24, seed=seed, edge_cases=True)[-4:]
    return pcm, synth.file_peaks(pcm)


def oracle_scores(oracle24, pcm, peak, keep_spec=False):
    from oracle import bn_oracle

    out, specs = [], []
    for s in range(0, len(pcm), 512):
        spec = bn_oracle.frontend_hybrid(pcm[s:s + 512], peak[s:s + 512], 512, HOP24, W)
        out.append(oracle24.predict(spec))
        if keep_spec:
            specs.append(spec)
    return np.concatenate(out), (np.concatenate(specs) if keep_spec else None)


def check_scores(got, ref, what):
    assert got.shape == ref.shape and got.dtype == np.float32
    d = np.abs(got - ref)
    assert d.max() <= LSB + 1e-7, f"{what}: max |delta| = {d.max() * 256:.2f} LSB"
    srt = np.sort(ref, axis=1)
    decided = (srt[:, -1] - srt[:, -2]) > LSB + 1e-7          # oracle top-1 margin above one LSB
    assert np.array_equal(got.argmax(1)[decided], ref.argmax(1)[decided]), f"{what}: top-1 differs on a decided chunk"
    return float((d.max(axis=1) == 0).mean())


def test_frontend_24k_hop281_matches_oracle(runner24):
    from oracle import bn_oracle

    pcm, peak = device_chunks(16, seed=11)
    ref = bn_oracle.frontend_hybrid(pcm, peak, 512, HOP24, W)
    got = runner24.frontend(pcm, peak)
    err = np.abs(got - ref).reshape(len(pcm), -1).max(axis=1)      # spectrograms are max-normalised to [0, 1]
    assert err.max() <= 1e-4, err
    scale = np.float32(0.003921568859368563)
    same = (np.round(ref / scale) == np.round(got / scale)).mean()
    assert same >= 0.999, same


def test_fused_head_output_codes_24k(runner24, oracle24):
    """Tensor #96 (mel mixer + PWL, the first tensor the fused plan materialises) from PCM through K1 `<frame-major>` +
    `k_head_tc` vs the oracle's from its own spectrogram: >= 99.9 % identical codes."""
    pcm, peak = device_chunks(64, seed=12)
    runner24.predict_pcm16(pcm, peak)
    got = runner24.dump_tensor(96, 64 * 256 * len(pcm))
    _, spec = oracle_scores(oracle24, pcm, peak, keep_spec=True)
    _, ref = oracle24.run(spec, tap_id=96)
    same = (got == ref.reshape(-1)).mean()
    assert same >= 0.999, same


@pytest.mark.parametrize("B", [1, 16, 256, 4096])
def test_batch_sweep_device_wave_and_host_wave(runner24, oracle24, B):
    """SURVEY 8(d) config 1.  Device-resident input = one wave of B chunks; host input = 592-chunk waves."""
    import torch

    pcm, peak = device_chunks(B, seed=100 + B)
    ref, spec = oracle_scores(oracle24, pcm, peak, keep_spec=True)
    # int8 body on the oracle's spectrograms: bit-exact, every chunk
    body = np.concatenate([runner24.predict(spec[s:s + 1024]) for s in range(0, B, 1024)])
    np.testing.assert_array_equal(body, ref)
    # full path, host buffers (BN_OPT_HOST_WAVE = 592 chunks per wave)
    host = runner24.predict_pcm16(pcm, peak)
    exact = check_scores(host, ref, f"B={B} host")
    # full path, device buffers (one wave)
    d_pcm, d_peak = torch.as_tensor(pcm, device="cuda"), torch.as_tensor(peak, device="cuda")
    d_out = torch.empty((B, 100), dtype=torch.float32, device="cuda")
    runner24.infer_pcm16_ptr(d_pcm.data_ptr(), d_peak.data_ptr(), B, d_out.data_ptr())
    torch.cuda.synchronize()
    np.testing.assert_array_equal(d_out.cpu().numpy(), host)       # wave size is invisible
    assert exact >= 0.98, f"only {exact:.3f} of the chunks have scores identical to the oracle"


def test_ten_thousand_chunks_pooled(runner24, oracle24):
    """>= 10 k chunks in files of 1..20 chunks through `bn_infer_pool` (device wave and host waves) vs the oracle."""
    import torch

    from oracle import bn_oracle

    rng = np.random.default_rng(5)
    counts = rng.integers(1, 21, size=980)
    offs = np.zeros(len(counts) + 1, np.int32)
    offs[1:] = np.cumsum(counts)
    n = int(offs[-1])
    assert n >= 10_000
    pcm, _ = device_chunks(n, seed=31)
    chunk_peak = np.abs(pcm.astype(np.float32) / np.float32(32768.0)).max(axis=1)
    peak = np.concatenate([np.full(c, chunk_peak[a:a + c].max(), np.float32) for a, c in zip(offs[:-1], counts)])   # file peak
    ref, _ = oracle_scores(oracle24, pcm, peak)
    got_chunks = runner24.predict_pcm16(pcm, peak)
    exact = check_scores(got_chunks, ref, "10k chunks")
    assert exact >= 0.98, exact
    want = np.stack([bn_oracle.pool_scores(ref[a:b], "lme", 10.0) for a, b in zip(offs[:-1], offs[1:])])
    host = runner24.predict_pooled(pcm, peak, offs, "lme", 10.0)
    assert np.abs(host - want).max() <= LSB + 1e-6
    d_pcm, d_peak, d_offs = (torch.as_tensor(a, device="cuda") for a in (pcm, peak, offs))
    d_out = torch.empty((len(counts), 100), dtype=torch.float32, device="cuda")
    runner24.infer_pool_ptr(d_pcm.data_ptr(), d_peak.data_ptr(), d_offs.data_ptr(), len(counts), "lme", 10.0, d_out.data_ptr())
    torch.cuda.synchronize()
    np.testing.assert_array_equal(d_out.cpu().numpy(), host)
    # file-level ranking metric agrees to 3 decimals (cmAP over synthetic labels)
    from sklearn.metrics import average_precision_score

    y = np.zeros((len(counts), 100), np.int32)
    y[np.arange(len(counts)), rng.integers(0, 100, len(counts))] = 1
    keep = y.sum(0) > 0
    a = np.mean([average_precision_score(y[:, c], host[:, c]) for c in np.where(keep)[0]])
    b = np.mean([average_precision_score(y[:, c], want[:, c]) for c in np.where(keep)[0]])
    assert round(float(a), 3) == round(float(b), 3)
