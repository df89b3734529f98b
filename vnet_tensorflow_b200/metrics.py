"""Step metrics of the reference's summary block (model.py:586-626) from the integer counts `vnb_read_metrics` returns.

The reference builds, per foreground class, `tf.metrics.true_positives / true_negatives / false_positives /
false_negatives / auc` on one-hot label and prediction volumes and derives sensitivity, specificity and Dice from the
update ops; `accuracy` is the mean of `pred == label`.  Its accumulators are reset before every step (model.py:730), so
each value describes one batch.  tf.metrics counts are float32 variables and the derived scalars are float32 divisions
(0/0 gives NaN, as `tf.divide` does); this module follows that arithmetic on the counts the device produced.
Counts are exact integers here and rounded to float32 once; TensorFlow sums 0/1 floats, which is the same number up
to 2^24 = 16.7 M voxels per count (a batch of eight 128^3 patches) and may differ in the last float32 digit beyond.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Optional, Sequence

import numpy as np

AUC_THRESHOLDS = 200      # tf.metrics.auc default num_thresholds
AUC_EPSILON = 1.0e-6      # metrics_impl.auc: epsilon of compute_auc


def auc_thresholds() -> np.ndarray:
    """metrics_impl.auc: [0 - 1e-7] + [(i + 1) / (n - 1) for i in range(n - 2)] + [1 + 1e-7], as a float32 constant."""
    n = AUC_THRESHOLDS
    return np.asarray([0.0 - 1e-7] + [(i + 1) * 1.0 / (n - 1) for i in range(n - 2)] + [1.0 + 1e-7], np.float32)


def _auc_from_counts(tp, fn, tn, fp) -> np.float32:
    """compute_auc(curve='ROC', summation_method='trapezoidal') on float32 count vectors over the thresholds."""
    f = np.float32
    eps = f(AUC_EPSILON)
    rec = (tp + eps) / (tp + fn + eps)
    fpr = fp / (fp + tn + eps)
    n = AUC_THRESHOLDS
    return f(np.sum((fpr[:n - 1] - fpr[1:]) * ((rec[:n - 1] + rec[1:]) / f(2.0)), dtype=np.float32))


def step_metrics(confusion: np.ndarray, auc_hist: Optional[np.ndarray] = None,
                 label_classes: Optional[Sequence[int]] = None) -> "OrderedDict[str, np.float32]":
    """confusion [K+1, K] (label row incl. the out-of-range row, argmax column), auc_hist [K, 2, 201] or None ->
    the scalars model.py:590,620-623 writes: accuracy, then sensitivity_/specificity_/dice_/auc_<label value> for every
    class but the first (model.py:601-603).  The four counts are returned as well (true_positives_<label value>, ...)."""
    f = np.float32
    cm = np.asarray(confusion).astype(np.int64)
    K = cm.shape[1]
    if cm.shape[0] != K + 1:
        raise ValueError("confusion must be [K+1, K]")
    names = [str(c) for c in (label_classes if label_classes is not None else range(K))]
    total = int(cm.sum())
    out: "OrderedDict[str, np.float32]" = OrderedDict()
    with np.errstate(divide="ignore", invalid="ignore"):
        out["accuracy"] = f(int(np.trace(cm[:K]))) / f(total)
        for i in range(1, K):
            tp = int(cm[i, i])
            fn = int(cm[i].sum()) - tp
            fp = int(cm[:, i].sum()) - tp
            tn = total - tp - fn - fp
            tp32, fn32, fp32, tn32 = f(tp), f(fn), f(fp), f(tn)
            out["true_positives_" + names[i]], out["true_negatives_" + names[i]] = tp32, tn32
            out["false_positives_" + names[i]], out["false_negatives_" + names[i]] = fp32, fn32
            out["sensitivity_" + names[i]] = tp32 / (tp32 + fn32)
            out["specificity_" + names[i]] = tn32 / (tn32 + fp32)
            out["dice_" + names[i]] = f(2.0) * tp32 / (f(2.0) * tp32 + fp32 + fn32)
            if auc_hist is not None:
                h = np.asarray(auc_hist).astype(np.int64)
                # bin b = number of thresholds below the probability: predicted positive at threshold j  <=>  b >= j + 1
                above_pos = np.cumsum(h[i, 1, ::-1])[::-1]      # above_pos[b] = #positives with bin >= b
                above_neg = np.cumsum(h[i, 0, ::-1])[::-1]
                tpj = above_pos[1:].astype(np.float32)
                fpj = above_neg[1:].astype(np.float32)
                fnj = (above_pos[0] - above_pos[1:]).astype(np.float32)
                tnj = (above_neg[0] - above_neg[1:]).astype(np.float32)
                out["auc_" + names[i]] = _auc_from_counts(tpj, fnj, tnj, fpj)
    return out
