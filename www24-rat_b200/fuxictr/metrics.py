"""Validation metrics (reference: fuxictr/metrics.py:21-41): sklearn AUC + logloss with predictions clipped to
[1e-7, 1-1e-7] (the reference passes eps=1e-7, which newer sklearn no longer accepts)."""
import logging

import numpy as np
from sklearn.metrics import log_loss, roc_auc_score


def evaluate_metrics(y_true, y_pred, metrics, **kwargs):
    result = dict()
    for metric in metrics:
        if metric in ["logloss", "binary_crossentropy"]:
            result[metric] = log_loss(y_true, np.clip(y_pred, 1e-7, 1 - 1e-7))
        elif metric == "AUC":
            result[metric] = roc_auc_score(y_true, y_pred)
        else:
            assert "group_index" in kwargs, "group_index is required for GAUC"   # stubs in the reference too
    logging.info("[Metrics] " + " - ".join("{}: {:.6f}".format(k, v) for k, v in result.items()))
    return result
