"""GPU parity tests (run on the B200 box): CUDA engine through the C ABI vs the CPU oracle."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ACT_TENSORS = [11, 76, 77, 81, 82, 83, 84, 85, 86, 87, 88, 89, 90, 91, 92, 93, 94, 95, 96] + list(range(97, 130))


@pytest.fixture(scope="module")
def runner(blob, cfg):
    from birdnet_stm32.evaluation.gpu_runner import GpuRunner

    r = GpuRunner(blob, cfg)
    yield r
    r.close()


@pytest.fixture(scope="module")
def oracle_spec(pcm_batch):
    from oracle import bn_oracle

    pcm, peak = pcm_batch
    return bn_oracle.frontend_hybrid(pcm, peak, 512, 66150 // 256, 256)


def test_frontend_matches_oracle(runner, pcm_batch, oracle_spec):
    """Float frontend tolerance: max-normalised error <= 1e-4 and >= 99.9 % identical
    int8 input codes after the graph's QUANTIZE (scale 1/255, zp -128)."""
    pcm, peak = pcm_batch
    got = runner.frontend(pcm, peak)
    assert got.shape == oracle_spec.shape and got.dtype == np.float32
    err = np.abs(got - oracle_spec).reshape(len(pcm), -1).max(axis=1)
    assert err.max() <= 1e-4, err
    scale = np.float32(0.003921568859368563)
    q_ref = np.clip(np.round(oracle_spec / scale) - 128, -128, 127)
    q_got = np.clip(np.round(got / scale) - 128, -128, 127)
    same = (q_ref == q_got).mean()
    assert same >= 0.999, same
    # silence stays exactly zero
    np.testing.assert_array_equal(got[-1], 0.0)


def test_graph_every_tensor_bit_exact_generic(runner, graph, oracle_model, oracle_spec):
    """Int8 body: identical quantised input -> every tensor identical (generic plan, debug taps)."""
    from birdnet_stm32 import _lib as L

    runner.set_option(L.BN_OPT_FORCE_GENERIC, 1)
    try:
        B = oracle_spec.shape[0]
        got = runner.predict(oracle_spec)
        ref = oracle_model.predict(oracle_spec)
        np.testing.assert_array_equal(got, ref)
        # BN_OPT_FUSION = 0: no multi-op launches in the generic plan either (a 1x1 convolution + the ADD behind it would
        # otherwise leave the convolution's own output unwritten), so every op materialises its tensor
        runner.set_option(L.BN_OPT_FUSION, 0)
        np.testing.assert_array_equal(runner.predict(oracle_spec), ref)
        for tid in ACT_TENSORS:
            t = graph.tensor(tid)
            if t.is_const or t.dtype != np.int8:
                continue
            nb = int(np.prod(t.shape[1:]))
            g = runner.dump_tensor(tid, nb * B)
            _, o = oracle_model.run(oracle_spec, tap_id=tid)
            assert np.array_equal(g, o.reshape(-1)), f"tensor {tid} ({t.name}) differs in {(g != o.reshape(-1)).sum()} of {g.size}"
    finally:
        runner.set_option(L.BN_OPT_FUSION, 139)
        runner.set_option(L.BN_OPT_FORCE_GENERIC, 0)


def test_graph_bit_exact_default_plan(runner, oracle_model, oracle_spec):
    got = runner.predict(oracle_spec)
    ref = oracle_model.predict(oracle_spec)
    np.testing.assert_array_equal(got, ref)
    # dynamic batch, like tests/test_runners.py:73-81 in the reference
    one = runner.predict(oracle_spec[2:3])
    np.testing.assert_array_equal(one, ref[2:3])
    assert one.dtype == np.float32 and one.shape == (1, 100)


BLOCK_OUT_TAPS = [96, 97, 99, 102, 104, 107, 110, 112, 115, 118, 121, 123, 126]
DW_OUT_TAPS = [98, 100, 103, 105, 108, 111, 113, 116, 119, 122, 124]


def test_fused_plan_is_active_and_its_tensors_are_bit_exact(runner, graph, oracle_model, oracle_spec):
    """The shipped graph must match the fused pattern; every tensor the fused plan materialises
    (head output #96, stem #97, every DS-block output) equals the oracle's.  With BN_OPT_FUSION = 0 the
    depthwise outputs are materialised too and must also match."""
    from birdnet_stm32 import _lib as L

    assert runner.query().fast_path == 1
    B = oracle_spec.shape[0]
    # fusion 7 = the variant with the depthwise convs of the stride-1 blocks on the tensor core as well (bn_ds_tc.cu);
    # bit 3 (8) = whole-stage kernel for the 8 x 16 stage (bn_stage.cu): its inner block outputs #112 / #115 / #118 stay in
    # shared memory unless bit 4 (16) asks for them (debug taps); bit 5 (32) = quantising frontend, irrelevant for the
    # spectrogram entry used here; bit 6 (64) = warp-specialised pipeline form of the per-block kernel (bn_ds_ws.cu).
    # 139 = 11 + 128 is the default.
    inner = {112, 115, 118}
    for fusion, taps in ((11, [t for t in BLOCK_OUT_TAPS if t not in inner]), (27, BLOCK_OUT_TAPS), (3, BLOCK_OUT_TAPS), (7, BLOCK_OUT_TAPS),
                         (67, BLOCK_OUT_TAPS), (75, [t for t in BLOCK_OUT_TAPS if t not in inner]),
                         (139, [t for t in BLOCK_OUT_TAPS if t not in inner]),     # bit 7 (128): stem as an im2col GEMM on tcgen05 (bn_stem_tc.cu)
                         (395, [t for t in BLOCK_OUT_TAPS if t not in inner and t != 97]),   # bit 8 (256): stem computed inside the first block's kernel (#97 stays on the SM)
                         (0, BLOCK_OUT_TAPS + DW_OUT_TAPS)):
        runner.set_option(L.BN_OPT_FUSION, fusion)
        try:
            got = runner.predict(oracle_spec)
            np.testing.assert_array_equal(got, oracle_model.predict(oracle_spec))
            for tid in taps:
                t = graph.tensor(tid)
                nb = int(np.prod(t.shape[1:]))
                g = runner.dump_tensor(tid, nb * B)
                _, o = oracle_model.run(oracle_spec, tap_id=tid)
                assert np.array_equal(g, o.reshape(-1)), f"fusion={fusion} tensor {tid} differs in {(g != o.reshape(-1)).sum()} of {g.size}"
        finally:
            runner.set_option(L.BN_OPT_FUSION, 139)


def test_fused_kernels_equal_layer_kernels_on_ragged_batches(blob, cfg, synth):
    """Fused DS-block / frontend kernels vs one kernel per layer on batches that do not fill the last tile
    (the 4x8 layers pack 4 chunks per MMA tile) and that span several waves: identical scores."""
    from birdnet_stm32 import _lib as L
    from birdnet_stm32.evaluation.gpu_runner import GpuRunner

    pcm = synth.synth_pcm16(37, 66150, 22050, seed=77, edge_cases=True)
    peak = synth.file_peaks(pcm)
    r = GpuRunner(blob, cfg, wave=16)
    try:
        fused = r.predict_pcm16(pcm, peak)
        r.set_option(L.BN_OPT_FUSION, 7)            # tensor-core depthwise variant, ragged last tile
        np.testing.assert_array_equal(fused, r.predict_pcm16(pcm, peak))
        r.set_option(L.BN_OPT_FUSION, 395)          # stem inside the first DS block's kernel, ragged last tile
        np.testing.assert_array_equal(fused, r.predict_pcm16(pcm, peak))
        r.set_option(L.BN_OPT_FUSION, 139)          # tensor-core stem (bn_stem_tc.cu)
        np.testing.assert_array_equal(fused, r.predict_pcm16(pcm, peak))
        r.set_option(L.BN_OPT_FUSION, 67)           # warp-specialised DS blocks (bn_ds_ws.cu), ragged last tile
        np.testing.assert_array_equal(fused, r.predict_pcm16(pcm, peak))
        r.set_option(L.BN_OPT_FUSION, 43)           # quantising frontend K1q + K2q (bn_frontend_q.cu) instead of K1 + float32 scratch + K2
        np.testing.assert_array_equal(fused, r.predict_pcm16(pcm, peak))
        q = r.predict_pcm16(pcm[:5], peak[:5])       # a batch smaller than one CTA's pair of workers
        np.testing.assert_array_equal(fused[:5], q)
        r.set_option(L.BN_OPT_FUSION, 3)            # ... and one kernel per DS block instead of the whole-stage kernel
        np.testing.assert_array_equal(fused, r.predict_pcm16(pcm, peak))
        r.set_option(L.BN_OPT_FUSION, 0)
        layer = r.predict_pcm16(pcm, peak)
        np.testing.assert_array_equal(fused, layer)
        for n in (1, 2, 3, 5):
            r.set_option(L.BN_OPT_FUSION, 139)
            a = r.predict_pcm16(pcm[:n], peak[:n])
            np.testing.assert_array_equal(a, layer[:n])
    finally:
        r.close()


def test_full_path_same_codes_as_generic_plan(runner, pcm_batch):
    """PCM16 path: fused plan (frame-major STFT + fused quantise) vs generic plan (bin-major STFT,
    normalise, QUANTIZE op): identical scores -> the fused head quantises exactly like the op chain."""
    from birdnet_stm32 import _lib as L

    pcm, peak = pcm_batch
    fused = runner.predict_pcm16(pcm, peak)
    runner.set_option(L.BN_OPT_FORCE_GENERIC, 1)
    try:
        generic = runner.predict_pcm16(pcm, peak)
    finally:
        runner.set_option(L.BN_OPT_FORCE_GENERIC, 0)
    np.testing.assert_array_equal(fused, generic)


def test_tensor_core_pointwise_equals_cuda_core_pointwise(runner, pcm_batch):
    """tcgen05 kind::i8 GEMM + TMEM epilogue vs the dp4a GEMM: identical scores and block outputs."""
    from birdnet_stm32 import _lib as L

    pcm, peak = pcm_batch
    tc = runner.predict_pcm16(pcm, peak)
    taps_tc = [runner.dump_tensor(t, n * len(pcm)) for t, n in ((99, 32 * 64 * 32), (102, 32 * 64 * 32), (126, 4 * 8 * 256))]
    runner.set_option(L.BN_OPT_TENSOR_CORE, 0)
    try:
        cc = runner.predict_pcm16(pcm, peak)
        taps_cc = [runner.dump_tensor(t, n * len(pcm)) for t, n in ((99, 32 * 64 * 32), (102, 32 * 64 * 32), (126, 4 * 8 * 256))]
    finally:
        runner.set_option(L.BN_OPT_TENSOR_CORE, 1)
    for a, b in zip(taps_tc, taps_cc):
        assert np.array_equal(a, b), f"{(a != b).sum()} of {a.size} differ"
    np.testing.assert_array_equal(tc, cc)


def test_rounding_and_mean_variants_match_oracle(runner, blob, oracle_spec):
    from birdnet_stm32 import _lib as L
    from oracle import bn_oracle

    for rounding, variant in ((1, 0), (0, 1), (0, 3)):
        runner.set_option(L.BN_OPT_ROUNDING, rounding)
        runner.set_option(L.BN_OPT_MEAN_VARIANT, variant)
        try:
            got = runner.predict(oracle_spec)
        finally:
            runner.set_option(L.BN_OPT_ROUNDING, 0)
            runner.set_option(L.BN_OPT_MEAN_VARIANT, 0)
        ref = bn_oracle.OracleModel(blob, rounding=rounding, mean_variant=variant).predict(oracle_spec)
        np.testing.assert_array_equal(got, ref)


def test_full_path_pcm_to_scores(runner, oracle_model, pcm_batch, oracle_spec):
    """PCM16 -> scores: top-1 equal, dequantised outputs within 1 LSB (1/256) of the oracle path."""
    pcm, peak = pcm_batch
    got = runner.predict_pcm16(pcm, peak)
    ref = oracle_model.predict(oracle_spec)
    assert np.abs(got - ref).max() <= 1.0 / 256 + 1e-7
    np.testing.assert_array_equal(got.argmax(axis=1), ref.argmax(axis=1))


def test_wave_splitting_is_invisible(blob, cfg, pcm_batch):
    from birdnet_stm32.evaluation.gpu_runner import GpuRunner

    pcm, peak = pcm_batch
    a = GpuRunner(blob, cfg, wave=3)
    b = GpuRunner(blob, cfg, wave=64)
    try:
        np.testing.assert_array_equal(a.predict_pcm16(pcm, peak), b.predict_pcm16(pcm, peak))
        assert a.launches > 0
    finally:
        a.close()
        b.close()


@pytest.mark.parametrize("method", ["avg", "max", "lme"])
def test_pooled_path_matches_oracle_pooling(runner, pcm_batch, method):
    from oracle import bn_oracle

    pcm, peak = pcm_batch
    offs = np.array([0, 3, 3, 4, 8], dtype=np.int32)      # ragged, one empty file
    chunk = runner.predict_pcm16(pcm, peak)
    got = runner.predict_pooled(pcm, peak, offs, pooling=method, beta=10.0)
    assert got.shape == (4, 100)
    for f in range(4):
        ref = bn_oracle.pool_scores(chunk[offs[f]:offs[f + 1]], method, 10.0)
        np.testing.assert_allclose(got[f], ref, rtol=0, atol=2e-6)
    np.testing.assert_array_equal(got[1], 0.0)             # empty file -> zeros (pooling.py:40-41)
    np.testing.assert_allclose(runner.pool_scores(chunk, offs, method, 10.0), got, atol=0)


def test_errors_are_reported_not_swallowed(runner, blob):
    from birdnet_stm32 import _lib as L
    from birdnet_stm32.evaluation.gpu_runner import GpuRunner

    with pytest.raises(ValueError):
        runner.predict(np.zeros((2, 17), np.float32))
    with pytest.raises(ValueError, match="Unsupported"):
        runner.predict_pooled(np.zeros((1, 66150), np.int16), None, [0, 1], pooling="median")
    with pytest.raises(L.EngineError):
        GpuRunner(blob[:1000])
    assert runner.predict(np.zeros((0, 257, 256, 1), np.float32)).shape == (0, 100)
