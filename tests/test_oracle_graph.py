"""Oracle self-checks for the int8 graph (no GPU, no TensorFlow).

The reference holds no golden vectors for `interpreter.invoke()` (SURVEY.md
section 4), so the oracle is validated structurally: every arithmetic op's
integer output is compared with an exact float64 evaluation of the same op on
the *dequantised integer inputs* taken from the `.tflite` directly (not from
the blob).  A correct restatement of the TFLite kernels (layouts, SAME-padding
asymmetry, per-channel axis, bias scale, requantisation, the ADD
three-multiplier scheme, MEAN, FC, LOGISTIC LUT) deviates by at most ~0.5
output LSB (+ the double-rounding slack); anything structural shows up as
many LSBs.
"""

import numpy as np
import pytest

from birdnet_stm32.conversion.export_blob import (
    activation_range, logistic_lut, quantize_multiplier, same_padding,
)

LSB_TOL = 0.53  # 0.5 + double-rounding slack of SRDHM followed by RoundingDivideByPOT


@pytest.fixture(scope="module")
def spec(pcm_batch):
    from oracle import bn_oracle

    pcm, peak = pcm_batch
    return bn_oracle.frontend_hybrid(pcm[:3], peak[:3], 512, 66150 // 256, 256)


def _tap(model, spec, tid, graph):
    t = graph.tensor(tid)
    if t.is_const:
        return None
    dt = np.float32 if t.dtype == np.float32 else np.int8
    _, tap = model.run(spec, tap_id=tid, tap_dtype=dt)
    return tap.reshape((spec.shape[0],) + tuple(t.shape[1:]))


def _deq(q, t):
    return (q.astype(np.float64) - t.zp()) * np.float64(t.s())


def _float_conv(x, w, stride, pad_tl, out_hw, depthwise):
    """x [B,H,W,C] float64 (real values, already zero == padding), w OHWI or 1HWC."""
    B, H, W, C = x.shape
    sh, sw = stride
    pt, pl = pad_tl
    oh, ow = out_hw
    if depthwise:
        _, kh, kw, oc = w.shape
    else:
        oc, kh, kw, _ = w.shape
    xp = np.zeros((B, H + kh + sh, W + kw + sw, C))
    xp[:, pt:pt + H, pl:pl + W, :] = x
    y = np.zeros((B, oh, ow, oc))
    for fy in range(kh):
        for fx in range(kw):
            patch = xp[:, fy:fy + sh * oh:sh, fx:fx + sw * ow:sw, :]
            if depthwise:
                y += patch * w[0, fy, fx, :]
            else:
                y += patch @ w[:, fy, fx, :].T
    return y


def test_every_arithmetic_op_matches_float_shadow(graph, oracle_model, spec):
    g = graph
    checked = 0
    worst = 0.0
    for op in g.ops:
        if op.kind not in ("CONV_2D", "DEPTHWISE_CONV_2D", "FULLY_CONNECTED", "ADD", "MEAN", "LOGISTIC", "QUANTIZE", "DEQUANTIZE"):
            continue
        y_t = g.tensor(op.outputs[0])
        y_q = _tap(oracle_model, spec, op.outputs[0], g)
        x_t = g.tensor(op.inputs[0])
        x_q = _tap(oracle_model, spec, op.inputs[0], g)
        if op.kind == "QUANTIZE":
            ideal = x_q.astype(np.float64) / np.float64(y_t.s()) + y_t.zp()
            lo, hi = -128, 127
        elif op.kind == "DEQUANTIZE":
            np.testing.assert_array_equal(y_q, (np.float32(x_t.s()) * (x_q.astype(np.int32) - x_t.zp()).astype(np.float32)))
            checked += 1
            continue
        elif op.kind in ("CONV_2D", "DEPTHWISE_CONV_2D", "FULLY_CONNECTED"):
            w_t, b_t = g.tensor(op.inputs[1]), g.tensor(op.inputs[2])
            ws = w_t.scale.astype(np.float64)
            x = _deq(x_q, x_t)
            if op.kind == "FULLY_CONNECTED":
                w = w_t.data.astype(np.float64) * ws[:, None]
                real = x @ w.T
                act = op.options["act"]
            else:
                dw = op.kind == "DEPTHWISE_CONV_2D"
                if dw:
                    w = w_t.data.astype(np.float64) * ws[None, None, None, :]
                else:
                    w = w_t.data.astype(np.float64) * ws[:, None, None, None]
                kh, kw = w_t.shape[1], w_t.shape[2]
                sh, sw = op.options["stride_h"], op.options["stride_w"]
                oh, pt = same_padding(x.shape[1], kh, sh)
                ow, pl = same_padding(x.shape[2], kw, sw)
                assert (oh, ow) == tuple(y_t.shape[1:3])
                real = _float_conv(x, w, (sh, sw), (pt, pl), (oh, ow), dw)
                act = op.options["act"]
            bias_scale = np.float64(x_t.s()) * ws
            real = real + b_t.data.astype(np.float64) * bias_scale
            ideal = real / np.float64(y_t.s()) + y_t.zp()
            lo, hi = activation_range(act, y_t.s(), y_t.zp())
        elif op.kind == "ADD":
            b_t = g.tensor(op.inputs[1])
            b_q = b_t.data if b_t.is_const else _tap(oracle_model, spec, op.inputs[1], g)
            real = _deq(x_q, x_t) + _deq(np.asarray(b_q), b_t)
            ideal = real / np.float64(y_t.s()) + y_t.zp()
            lo, hi = activation_range(op.options["act"], y_t.s(), y_t.zp())
        elif op.kind == "MEAN":
            real = _deq(x_q, x_t).mean(axis=(1, 2))
            ideal = real / np.float64(y_t.s()) + y_t.zp()
            lo, hi = -128, 127
        elif op.kind == "LOGISTIC":
            real = 1.0 / (1.0 + np.exp(-_deq(x_q, x_t)))
            ideal = real / np.float64(y_t.s()) + y_t.zp()
            lo, hi = -128, 127
        ideal = np.clip(ideal, lo, hi)
        err = np.abs(ideal - y_q.astype(np.float64))
        # tiny-scale tensors (dead PWL branches, scale ~8e-9): ideal values explode; still must agree after clamping
        worst = max(worst, float(err.max()))
        assert err.max() <= LSB_TOL, f"op {op.index} {op.kind}: max deviation {err.max():.3f} LSB"
        assert y_q.min() >= lo and y_q.max() <= hi
        checked += 1
    assert checked >= 45
    print(f"float-shadow: {checked} ops checked, worst deviation {worst:.3f} LSB")


def test_dead_pwl_branches_are_constant_zero(graph, oracle_model, spec):
    """tensors #86/#89/#92 (relu(x - t_i), scale 7.8e-9, zp 0) are constant 0."""
    for tid in (86, 89, 92):
        tap = _tap(oracle_model, spec, tid, graph)
        assert np.all(tap == 0)


def test_layout_ops_copy_codes(graph, oracle_model, spec):
    q = _tap(oracle_model, spec, 11, graph)          # QUANTIZE out [B,257,256,1]
    t = _tap(oracle_model, spec, 76, graph)          # TRANSPOSE  [B,1,256,257]
    np.testing.assert_array_equal(t, q.transpose(0, 3, 2, 1))
    c = _tap(oracle_model, spec, 82, graph)          # CONCAT     [B,1,256,264]
    np.testing.assert_array_equal(c[..., :257], t)
    assert np.all(c[..., 257:] == -128)
    a = _tap(oracle_model, spec, 94, graph)
    b = _tap(oracle_model, spec, 96, graph)          # TRANSPOSE + SLICE -> [B,64,256,1]
    np.testing.assert_array_equal(b, a.transpose(0, 3, 2, 1))


def test_batch_independence_and_determinism(oracle_model, spec):
    a = oracle_model.predict(spec)
    b = oracle_model.predict(spec[::-1].copy())[::-1]
    np.testing.assert_array_equal(a, b)
    c = oracle_model.predict(spec[1:2])
    np.testing.assert_array_equal(a[1:2], c)
    assert a.dtype == np.float32 and a.shape == (spec.shape[0], 100)
    # LOGISTIC output is q/256 with q in [0,255]
    assert np.all(a >= 0) and np.all(a <= 255 / 256)
    np.testing.assert_array_equal(a * 256, np.round(a * 256))


def test_requant_primitives_known_answers():
    from oracle import bn_oracle as bo

    # QuantizeMultiplier known answers (TFLite quantization_util_test.cc style)
    assert quantize_multiplier(0.0) == (0, 0)
    assert quantize_multiplier(1.0) == (1 << 30, 1)
    assert quantize_multiplier(0.5) == (1 << 30, 0)
    assert quantize_multiplier(0.25) == (1 << 30, -1)
    assert quantize_multiplier(2.0 ** -40) == (0, 0)            # shift < -31 -> zero multiplier
    m, s = quantize_multiplier(0.99999999999)
    assert (m, s) == (1 << 30, 1)                                # rounds up to 2^31 -> halved, shift+1
    # SaturatingRoundingDoublingHighMul rounds ties toward +inf (-1.5 -> -1): gemmlowp's known quirk
    q, sh = quantize_multiplier(0.5)
    assert [bo.mbqm(v, q, sh) for v in (3, -3, 5, -5, 4, 0)] == [2, -1, 3, -2, 2, 0]
    # RoundingDivideByPOT rounds ties away from zero (x * 0.25: 6 -> 1.5 -> 2, -6 -> -1.5 -> -2)
    q, sh = quantize_multiplier(0.25)
    assert [bo.mbqm(v, q, sh) for v in (6, -6, 2, -2, 4)] == [2, -2, 1, -1, 1]
    # extreme parameters seen in the shipped checkpoint: shift -31, multiplier 0, |bias| 2^30
    assert bo.mbqm(1 << 30, 1 << 30, -31) == 0
    assert bo.mbqm(-(1 << 30), 0, 0) == 0
    assert bo.mbqm(-(2 ** 31), -(2 ** 31), 0) == 2 ** 31 - 1     # SRDHM saturation
    # exhaustive check of the fused formula (a*b + 2^30) >> 31 used on the GPU vs gemmlowp's
    rng = np.random.default_rng(0)
    xs = rng.integers(-(1 << 31), 1 << 31, 2000)
    ms = rng.integers(1 << 30, 1 << 31, 2000)
    for x, mm in zip(xs.tolist(), ms.tolist()):
        assert bo.mbqm(x, mm, 0) == ((x * mm + (1 << 30)) >> 31)
    # double vs single rounding do differ (that is why the mode is explicit)
    diff = sum(bo.mbqm(x, mm, -7, 0) != bo.mbqm(x, mm, -7, 1) for x, mm in zip((xs >> 12).tolist(), ms.tolist()))
    assert diff > 0


def test_logistic_lut_matches_c_library(graph):
    from oracle import bn_oracle as bo

    op = [o for o in graph.ops if o.kind == "LOGISTIC"][0]
    x, y = graph.tensor(op.inputs[0]), graph.tensor(op.outputs[0])
    a = logistic_lut(x.s(), x.zp(), y.s(), y.zp())
    b = bo.logistic_lut(x.s(), x.zp(), y.s(), y.zp())
    np.testing.assert_array_equal(a, b)
    assert a[0] <= a[255] and a.min() >= -128


def test_rounding_mode_and_mean_variant_are_switchable(blob, spec):
    from oracle import bn_oracle as bo

    base = bo.OracleModel(blob).run(spec, tap_id=127)[1]
    for variant in (1, 2, 3):
        v = bo.OracleModel(blob, mean_variant=variant).run(spec, tap_id=127)[1]
        assert np.abs(v.astype(int) - base.astype(int)).max() <= 1   # all are valid roundings
    single = bo.OracleModel(blob, rounding=1).run(spec, tap_id=83)[1]   # first requantised tensor (mel mixer)
    double = bo.OracleModel(blob, rounding=0).run(spec, tap_id=83)[1]
    d = np.abs(single.astype(int) - double.astype(int))
    assert d.max() <= 1 and 0 < (d > 0).mean() < 0.01


def test_gpu_closed_form_requant_equals_gemmlowp():
    """rq_fast (the closed form the CUDA kernels use for right shifts) == SRDHM + RoundingDivideByPOT."""
    from oracle import bn_oracle as bo

    rng = np.random.default_rng(1)
    edge = [0, 1, -1, 2, -2, 3, -3, (1 << 30), -(1 << 30), (1 << 31) - 1, -(1 << 31), 12345, -12345]
    mults = [0, 1 << 30, (1 << 31) - 1, 1518500250, 1073741825]
    for n in range(1, 32):
        xs = edge + rng.integers(-(1 << 31), 1 << 31, 200).tolist() + rng.integers(-(1 << n), 1 << n, 100).tolist()
        for m in mults + rng.integers(1 << 30, 1 << 31, 5).tolist():
            for x in xs:
                v = (x * m + (1 << 30)) >> 31
                if abs(v) + (1 << (n - 1)) >= (1 << 31):
                    continue    # outside the int32-safe domain; the engine proves |v| bounds per layer at plan build
                assert bo.rq_fast(x, m, n) == bo.mbqm(x, m, -n, 0), (x, m, n)


def test_saturating_requant_forms_equal_gemmlowp():
    """The folded forms of csrc/bn_ds.cu == the gemmlowp sequence, in exact integer arithmetic.

    (1) rq_hi: y = hi32(acc*m + C) >> (n-1) with C = bias*m + 2^30 + (2^(n-1) + zp*2^n)*2^31, zp = -128, then signed
        saturation to int8  ==  clamp(MBQM(acc + bias, m, -n) + zp, -128, 127);
    (2) residual term of ADD with zp1 = -128:  ((r+128)*m1 + 2^10 + 2^(n1+10)) >> (11+n1)
        ==  MBQM((r - zp1) << 20, m1, -n1);
    (3) ADD output: hi32(t*mo + 2^30 + (2^(no-1) + zp*2^no)*2^31) >> (no-1), saturated
        ==  clamp(MBQM(t, mo, -no) + zp, -128, 127).
    """
    from oracle import bn_oracle as bo

    rng = np.random.default_rng(5)
    sat = lambda v: max(-128, min(127, v))
    zp = -128
    for n in range(1, 25):
        for m in [1 << 30, (1 << 31) - 1, 1518500250] + rng.integers(1 << 30, 1 << 31, 6).tolist():
            bias = int(rng.integers(-(1 << 20), 1 << 20))
            C = bias * m + (1 << 30) + ((1 << (n - 1)) + zp * (1 << n)) * (1 << 31)
            span = min(400 << n, 1 << 30)         # covers the whole int8 output range around both clamps (int32-safe)
            accs = rng.integers(-span, span, 300).tolist() + [0, 1, -1, -bias, -bias + 1, -bias - 1]
            # exact ties of the second rounding, both signs
            accs += [(-bias) + k for k in range(-3, 4)]
            for acc in accs:
                want = sat(bo.mbqm(acc + bias, m, -n, 0) + zp)
                got = sat(((acc * m + C) >> 32) >> (n - 1))
                assert got == want, (acc, bias, m, n)
    for n1 in range(0, 22):
        for m1 in [1 << 30, (1 << 31) - 1] + rng.integers(1 << 30, 1 << 31, 6).tolist():
            c1 = (1 << 10) + ((1 << (n1 - 1 + 11)) if n1 > 0 else 0)
            for r in range(-128, 128):
                want = bo.mbqm((r + 128) << 20, m1, -n1, 0)
                got = ((r + 128) * m1 + c1) >> (11 + n1)
                assert got == want, (r, m1, n1)
    for no in range(1, 25):
        for mo in [1 << 30, (1 << 31) - 1] + rng.integers(1 << 30, 1 << 31, 6).tolist():
            co = (1 << 30) + ((1 << (no - 1)) + zp * (1 << no)) * (1 << 31)
            span = 400 << no
            ts = rng.integers(-min(span, 1 << 30), min(span, 1 << 30), 300).tolist() + list(range(-4, 5))
            for t in ts:
                want = sat(bo.mbqm(t, mo, -no, 0) + zp)
                got = sat(((t * mo + co) >> 32) >> (no - 1))
                assert got == want, (t, mo, no)
