"""Pooling: the reference's own known answers (tests/test_pooling.py:12-55 in the reference) and golden
outputs produced by importing the reference's evaluation/pooling.py (tests/golden/make_golden.py)."""

import os

import numpy as np
import pytest

from conftest import GOLDEN

IMPLS = ["package", "oracle"]


def _pool(impl):
    if impl == "package":
        from birdnet_stm32.evaluation.pooling import pool_scores

        return pool_scores
    from oracle.bn_oracle import pool_scores

    return pool_scores


@pytest.mark.parametrize("impl", IMPLS)
class TestKnownAnswers:
    def test_average(self, impl):
        s = np.array([[0.2, 0.8], [0.6, 0.4]], dtype=np.float32)
        np.testing.assert_allclose(_pool(impl)(s, method="avg"), [0.4, 0.6], rtol=1e-6)

    def test_max(self, impl):
        s = np.array([[0.1, 0.9], [0.7, 0.3]], dtype=np.float32)
        np.testing.assert_allclose(_pool(impl)(s, method="max"), [0.7, 0.9])

    def test_empty(self, impl):
        out = _pool(impl)(np.zeros((0, 5), dtype=np.float32), method="avg")
        np.testing.assert_array_equal(out, np.zeros(5))
        assert out.dtype == np.float32

    def test_invalid_method(self, impl):
        with pytest.raises(ValueError, match="Unsupported"):
            _pool(impl)(np.ones((3, 2), dtype=np.float32), method="invalid")

    def test_wrong_ndim(self, impl):
        with pytest.raises(ValueError, match="must be"):
            _pool(impl)(np.ones(5), method="avg")

    def test_lme_single_row(self, impl):
        s = np.array([[0.5, 0.3]], dtype=np.float32)
        np.testing.assert_allclose(_pool(impl)(s, method="lme", beta=10.0), [0.5, 0.3], atol=1e-5)

    def test_lme_high_beta_approaches_max(self, impl):
        s = np.array([[0.1, 0.9], [0.8, 0.2]], dtype=np.float32)
        np.testing.assert_allclose(_pool(impl)(s, method="lme", beta=100.0), [0.8, 0.9], atol=0.05)


@pytest.mark.parametrize("impl", IMPLS)
def test_matches_reference_outputs(impl):
    g = np.load(os.path.join(GOLDEN, "pooling_reference.npz"))
    pool = _pool(impl)
    n = 0
    for key in g.files:
        if not key.startswith("out_"):
            continue
        _, i, method, beta = key.split("_")
        got = pool(g[f"in_{i}"], method=method, beta=float(beta))
        np.testing.assert_allclose(got, g[key], rtol=0, atol=3e-6, err_msg=key)
        n += 1
    assert n == 30
    np.testing.assert_array_equal(pool(np.zeros((0, 5), np.float32), method="avg"), g["empty_avg"])


@pytest.mark.gpu
def test_device_pooling_matches_reference_outputs(blob, cfg):
    from birdnet_stm32.evaluation.gpu_runner import GpuRunner

    g = np.load(os.path.join(GOLDEN, "pooling_reference.npz"))
    r = GpuRunner(blob, cfg)
    try:
        for key in g.files:
            if not key.startswith("out_"):
                continue
            _, i, method, beta = key.split("_")
            s = g[f"in_{i}"]
            got = r.pool_scores(s, [0, s.shape[0]], method, float(beta))
            np.testing.assert_allclose(got[0], g[key], rtol=0, atol=3e-6, err_msg=key)
    finally:
        r.close()
