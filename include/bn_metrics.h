/* bn_metrics.h -- C ABI of the device-side metric tail (libbn_b200.so), SURVEY section 8 row f2.
 *
 *   reference interface (birdnet_stm32/evaluation/metrics.py)          replaced by
 *   -------------------------------------------------------------    ---------------------------
 *   roc_auc_score(y_true, y_scores, average="micro")          :159     bn_metrics_compute -> roc_auc_micro
 *   precision / recall / F1 at threshold 0.5                  :165-174                      -> precision, recall, f1
 *   average_precision_score per class, cmAP = mean            :176-185                      -> ap_per_class, cmap
 *   average_precision_score(..., average="micro")             :188                          -> map_micro
 *
 * The arithmetic is scikit-learn's (third party; pinned 1.7.1 in the reference's requirements.txt:5, 1.9.0 in this image and
 * used as the oracle): scores are sorted in descending order, thresholds are the distinct score values, precision =
 * tp / (tp + fp) and recall = tp / P at every threshold, AP = sum (R_k - R_{k-1}) P_k, ROC-AUC = trapezoidal area under
 * (fp / N, tp / P); everything after the integer counts is float64.  A class without positives gets AP 0.0 (sklearn: "No
 * positive class found in y_true, recall is set to one for all thresholds") and is flagged in `classes_without_positives`;
 * ROC-AUC is NaN when y_true holds a single class (sklearn raises; the reference catches it and stores NaN).
 *
 * Conventions as in bn_engine.h.  y_true / y_score may be host or device pointers (both of one kind); the call is
 * synchronous (it is the serial tail of an evaluation, run on rank 0 after the score all-gather).
 */
#ifndef BN_METRICS_H
#define BN_METRICS_H

#include <stddef.h>
#include <stdint.h>

#include "bn_engine.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bn_metrics_result {
  double roc_auc_micro;
  double map_micro;
  double cmap;              /* mean of the per-class APs */
  double precision;         /* at threshold 0.5, over all (file, class) cells, + 1e-12 in the denominators like the reference */
  double recall;
  double f1;
  int64_t n_positive;       /* cells with y_true != 0 */
  int64_t n_cells;          /* F * C */
  int32_t classes_without_positives;
  int32_t n_launches;       /* kernels (own + CUB) launched by the call */
  int32_t reserved[4];
} bn_metrics_result;

/* y_true, y_score: float32 [F, C] row-major (y_true cells are 0 or 1).  ap_per_class: optional float64 [C] (host). */
BN_API int bn_metrics_compute(const float* y_true, const float* y_score, int F, int C, int device, bn_metrics_result* out,
                              double* ap_per_class);

/* Bootstrap resamples of the per-class average precision (reference: bootstrap_ap_ci, evaluation/metrics.py:240-322: n_bootstrap
 * resamples of the F files with replacement per class, AP of every resample, percentiles).  ap_samples: float64 [C, n_boot]
 * (host), NaN where the resample holds a single class of that column (the reference skips those).  A resample is a vector of
 * multiplicities over the files; every class is sorted ONCE and a resample's AP is two running sums in that order.
 * multiplicities: optional int32 [n_boot, F] (host or device) -- the caller's own resamples, shared by all classes (this is how
 * the tests compare with the host formulation exactly); NULL = drawn on the device from a counter-based generator
 * (Philox-4x32-10 keyed by `seed`): statistically equivalent to the reference's intervals, not its numpy PCG64 stream. */
BN_API int bn_metrics_bootstrap_ap(const float* y_true, const float* y_score, int F, int C, int n_boot, unsigned long long seed,
                                   const int32_t* multiplicities, double* ap_samples, int device);

#ifdef __cplusplus
}
#endif
#endif /* BN_METRICS_H */
