"""The oracle derives every requantisation constant itself, straight from the `.tflite` file (`oracle/tflite_quant.py`: own
flatbuffer walker, exact rational arithmetic, glibc expf), and must agree integer for integer with what the product's
exporter wrote into the blob (VERDICT r1 "oracle independence").  The `oracle_model` fixture runs on the oracle's values."""

import numpy as np
import pytest

from conftest import TFLITE


def test_oracle_derivation_equals_the_blob_of_the_shipped_checkpoint(blob):
    from oracle import tflite_quant as tq

    derived = tq.derive(TFLITE)
    kinds = [d["kind"] for d in derived.values()]
    assert kinds.count("conv") == 30 and kinds.count("add") == 13 and kinds.count("mean") == 1 and kinds.count("logistic") == 1
    bad, n = tq.compare_with_blob(blob, derived)
    assert n > 6000 and not bad, bad[:5]
    assert tq.patch_blob(blob, derived) == blob


def test_a_wrong_constant_in_the_blob_is_caught(blob):
    """The comparison is not vacuous: flip one multiplier and one LUT byte."""
    import struct

    from oracle import tflite_quant as tq

    derived = tq.derive(TFLITE)
    ops = list(tq._blob_ops(blob))
    conv = next(o for o in ops if o["kind"] == 7)
    logi = next(o for o in ops if o["kind"] == 13)
    b = bytearray(blob)
    v = struct.unpack_from("<i", b, conv["off"][2])[0]
    struct.pack_into("<i", b, conv["off"][2], v + 1)          # mult[0] off by one unit in the last place
    b[logi["off"][0] + 200] ^= 1
    bad, _ = tq.compare_with_blob(bytes(b), derived)
    assert len(bad) == 2
    assert tq.patch_blob(bytes(b), derived) == blob            # and patching restores the oracle's values


def test_quantize_multiplier_edge_cases():
    from oracle import tflite_quant as tq

    assert tq.quantize_multiplier(0.0) == (0, 0)
    assert tq.quantize_multiplier(0.5) == (1 << 30, 0)
    assert tq.quantize_multiplier(1.0) == (1 << 30, 1)
    assert tq.quantize_multiplier(2.0 ** -32) == (1 << 30, -31)
    assert tq.quantize_multiplier(2.0 ** -33) == (0, 0)                      # shift < -31 flushes to zero
    assert tq.quantize_multiplier(2.0 ** -31) == (1 << 30, -30)
    assert tq.quantize_multiplier(float(np.nextafter(1.0, 0.0))) == (1 << 30, 1)   # rounds up to 2^31 -> renormalised
    # agrees with the frexp formulation on random doubles
    import math

    rng = np.random.default_rng(3)
    for m in np.exp(rng.uniform(np.log(1e-11), np.log(4.0), 4000)):
        q, e = math.frexp(float(m))
        qf = int(math.floor(q * 2.0 ** 31 + 0.5))
        if qf == 1 << 31:
            qf //= 2
            e += 1
        want = (0, 0) if e < -31 else (qf, e)
        assert tq.quantize_multiplier(float(m)) == want


@pytest.mark.parametrize("name", ["raw_48000", "wide_se_attn_per_channel", "wide_se_attn_per_tensor", "wide_ir_se_attn"])
def test_oracle_derivation_equals_the_blob_of_synthesised_graphs(name):
    """Configs 3 / 4 (SE gates -> MUL, keep_dims MEAN, per-tensor weight scales, requantising QUANTIZE)."""
    from oracle import tflite_quant as tq
    from test_ptq import _case

    _, raw, _, blob = _case(name)
    derived = tq.derive(raw)
    bad, n = tq.compare_with_blob(blob, derived)
    assert n > 1000 and not bad, bad[:5]
