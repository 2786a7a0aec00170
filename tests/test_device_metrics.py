"""Device metric tail against scikit-learn, the library the reference calls
(`evaluation/metrics.py:152-190`): ROC-AUC micro, per-class AP, cmAP, micro AP, precision / recall / F1 at 0.5.
Tolerance: 1e-12 absolute on every float64 result (same integer counts; only the order of the float64 sums differs)."""

import numpy as np
import pytest

from birdnet_stm32.evaluation.metrics import _metrics_from_scores


def make_case(rng, F, C, quantised, empty_classes=0, full_class=False):
    y_true = np.zeros((F, C), dtype=np.float32)
    y_true[np.arange(F), rng.integers(0, C - empty_classes, size=F)] = 1.0
    if full_class:
        y_true[:, 0] = 1.0
    s = rng.random((F, C)).astype(np.float32) * 0.7 + 0.3 * y_true * rng.random((F, C)).astype(np.float32)
    if quantised:
        s = np.round(s * 256) / 256          # LOGISTIC-like codes: many ties
    return y_true, s.astype(np.float32)


@pytest.mark.gpu
def test_device_metrics_match_sklearn():
    import warnings

    from birdnet_stm32.evaluation.device_metrics import metrics_from_scores_device

    rng = np.random.default_rng(9)
    cases = [make_case(rng, 37, 5, False), make_case(rng, 500, 100, True, empty_classes=7), make_case(rng, 2000, 100, False),
             make_case(rng, 64, 3, True, full_class=True), make_case(rng, 1, 4, False), make_case(rng, 3000, 33, True)]
    worst = 0.0
    for y_true, y_scores in cases:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ref = _metrics_from_scores(y_true, y_scores)
        got = metrics_from_scores_device(y_true, y_scores)
        for key in ("roc-auc", "f1", "precision", "recall", "cmAP", "mAP"):
            a, b = ref[key], got[key]
            assert (np.isnan(a) and np.isnan(b)) or abs(a - b) <= 1e-12, (key, a, b, y_true.shape)
            if not np.isnan(a):
                worst = max(worst, abs(a - b))
        ra, ga = np.asarray(ref["ap_per_class"], dtype=np.float64), np.asarray(got["ap_per_class"])
        assert ra.shape == ga.shape and np.abs(ra - ga).max() <= 1e-12
    print(f"device metrics vs sklearn: worst |d| = {worst:.3g}")
    # single-class y_true: sklearn raises inside roc_auc_score, the reference stores NaN
    y_true = np.zeros((10, 3), dtype=np.float32)
    got = metrics_from_scores_device(y_true, rng.random((10, 3)).astype(np.float32))
    assert np.isnan(got["roc-auc"]) and got["cmAP"] == 0.0 and got["mAP"] == 0.0


@pytest.mark.gpu
def test_evaluate_with_device_metrics_backend(tmp_path, blob, cfg):
    import warnings

    from test_evaluate import make_dataset

    from birdnet_stm32.evaluation.gpu_runner import GpuRunner
    from birdnet_stm32.evaluation.metrics import evaluate

    classes = cfg["class_names"]
    files = make_dataset(str(tmp_path), classes[:4], n_per_class=2, sr=22050)
    runner = GpuRunner(blob, cfg)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            m_ref, _, yt, ys = evaluate(runner, files, classes, cfg, pooling="lme")
            m_dev, _, yt2, ys2 = evaluate(runner, files, classes, cfg, pooling="lme", metrics_backend="device")
        assert np.array_equal(ys, ys2)
        for key in ("roc-auc", "f1", "precision", "recall", "cmAP", "mAP"):
            assert abs(m_ref[key] - m_dev[key]) <= 1e-12, key
        with pytest.raises(ValueError, match="metrics backend"):
            evaluate(runner, files, classes, cfg, metrics_backend="nope")
    finally:
        runner.close()
