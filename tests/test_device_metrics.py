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


@pytest.mark.gpu
def test_bootstrap_ap_on_the_device():
    """Row f2 tail (reference `bootstrap_ap_ci`, evaluation/metrics.py:240-322).  (1) Given the SAME resamples (multiplicity vectors)
    the device APs equal the host formulation (`_weighted_ap_sorted`, itself equal to scikit-learn on the materialised
    resample) to 1e-12, ties and single-class resamples included; (2) with its own generator the intervals agree with the
    host version's within Monte-Carlo error and every file is drawn F times per resample."""
    from sklearn.metrics import average_precision_score

    from birdnet_stm32.evaluation.device_metrics import bootstrap_ap_ci_device, bootstrap_ap_samples_device
    from birdnet_stm32.evaluation.metrics import bootstrap_ap_ci

    rng = np.random.default_rng(11)
    F, Cn, R = 700, 9, 64
    y = (rng.random((F, Cn)) < 0.15).astype(np.float32)
    y[:, 7] = 0.0                                               # a class without positives
    y[:, 8] = 0.0
    y[3, 8] = 1.0                                               # a single positive: many single-class resamples
    s = np.round(rng.random((F, Cn)), 2).astype(np.float32)     # heavy ties
    s[:, :3] = rng.random((F, 3)).astype(np.float32)
    idx = rng.integers(0, F, size=(R, F))
    mult = np.stack([np.bincount(r, minlength=F) for r in idx]).astype(np.int32)
    got = bootstrap_ap_samples_device(y, s, R, multiplicities=mult)
    assert got.shape == (Cn, R)
    for c in range(Cn):
        for r in range(0, R, 7):
            yt, ys = y[idx[r], c], s[idx[r], c]
            if yt.sum() == 0 or yt.sum() == F:
                assert np.isnan(got[c, r])
            else:
                assert abs(got[c, r] - average_precision_score(yt, ys)) < 1e-12, (c, r)
    # own generator: multiplicities sum to F, intervals close to the host version's
    classes = [f"c{i}" for i in range(Cn)]
    dev = bootstrap_ap_ci_device(y, s, classes, n_bootstrap=600, seed=5)
    host = bootstrap_ap_ci(y, s, classes, n_bootstrap=600, seed=5)
    for d, h in zip(dev, host):
        assert d["class"] == h["class"] and d["n_positive"] == h["n_positive"] and d["n_total"] == h["n_total"]
        assert abs(d["ap"] - h["ap"]) < 1e-12
        # two independent sets of 600 resamples: percentile estimates differ by Monte-Carlo error (larger for the class with a
        # single positive, whose resampled AP takes few distinct values)
        width = max(h["ci_upper"] - h["ci_lower"], 1e-9)
        tol = 0.35 * width + 0.003
        assert abs(d["ci_lower"] - h["ci_lower"]) < tol and abs(d["ci_upper"] - h["ci_upper"]) < tol, (d, h)
        assert d["ci_lower"] <= d["ci_upper"]
    a = bootstrap_ap_samples_device(y, s, 16, seed=1)
    b = bootstrap_ap_samples_device(y, s, 16, seed=1)
    c = bootstrap_ap_samples_device(y, s, 16, seed=2)
    np.testing.assert_array_equal(a, b)
    assert not np.array_equal(a[:3], c[:3])
