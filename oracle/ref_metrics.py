"""CPU oracle of the reference's step metrics (TEST INFRASTRUCTURE ONLY, see oracle/ref_vnet.py).

Restates model.py:586-626 literally - one-hot label and prediction volumes, the tf.metrics.* counters as boolean
reductions, tf.metrics.auc as the confusion matrix at 200 thresholds (TensorFlow 1.15 `metrics_impl.auc`,
`_confusion_matrix_at_thresholds`: un-vendored dependency, restated from its source) - on NumPy arrays.
TensorFlow's metric variables are float32, so are the values here.  Unpinned against a real TensorFlow (none here).
"""
from collections import OrderedDict

import numpy as np


def step_metrics(logits, labels, softmax, label_classes=None):
    """logits / softmax [N,X,Y,Z,K] float32, labels [N,X,Y,Z] int -> OrderedDict like vnet_tensorflow_b200.metrics."""
    f = np.float32
    K = logits.shape[-1]
    names = [str(c) for c in (label_classes if label_classes is not None else range(K))]
    pred = np.argmax(logits, axis=-1)                                  # model.py:568 (first maximum)
    lab = np.asarray(labels).reshape(pred.shape)
    out = OrderedDict()
    out["accuracy"] = f(np.mean((pred == lab).astype(np.float32), dtype=np.float64))   # model.py:588-589
    eye = np.eye(K, dtype=np.float32)
    valid = (lab >= 0) & (lab < K)
    label_one_hot = np.where(valid[..., None], eye[np.clip(lab, 0, K - 1)], f(0))       # tf.one_hot: no class -> zero row
    pred_one_hot = eye[pred]
    # metrics_impl.auc
    n = 200
    kepsilon = 1e-7
    thresholds = [(i + 1) * 1.0 / (n - 1) for i in range(n - 2)]
    thresholds = np.asarray([0.0 - kepsilon] + thresholds + [1.0 + kepsilon], np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        for i in range(1, K):                                          # model.py:601-603
            l = label_one_hot[..., i].astype(bool).reshape(-1)
            p = pred_one_hot[..., i].astype(bool).reshape(-1)
            tp, tn = f(np.sum(l & p)), f(np.sum(~l & ~p))
            fp, fn = f(np.sum(~l & p)), f(np.sum(l & ~p))
            out["true_positives_" + names[i]], out["true_negatives_" + names[i]] = tp, tn
            out["false_positives_" + names[i]], out["false_negatives_" + names[i]] = fp, fn
            out["sensitivity_" + names[i]] = tp / (tp + fn)            # model.py:616
            out["specificity_" + names[i]] = tn / (tn + fp)            # model.py:617
            out["dice_" + names[i]] = f(2.0) * tp / (f(2.0) * tp + fp + fn)   # model.py:618
            s = np.asarray(softmax[..., i], np.float32).reshape(-1)
            tpj, fnj, tnj, fpj = (np.zeros(n, np.float32) for _ in range(4))
            for j in range(n):                                         # _confusion_matrix_at_thresholds
                pos = s > thresholds[j]
                tpj[j], fnj[j] = np.sum(l & pos), np.sum(l & ~pos)
                tnj[j], fpj[j] = np.sum(~l & ~pos), np.sum(~l & pos)
            eps = f(1.0e-6)
            rec = (tpj + eps) / (tpj + fnj + eps)
            fpr = fpj / (fpj + tnj + eps)
            out["auc_" + names[i]] = f(np.sum((fpr[:n - 1] - fpr[1:]) * ((rec[:n - 1] + rec[1:]) / f(2.0)), dtype=np.float32))
    return out
