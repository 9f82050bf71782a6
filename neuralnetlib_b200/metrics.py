"""Host-side metrics used by fit()/EarlyStopping (mirror of neuralnetlib/metrics.py:16-89 plus the few scores the
reference's own unit tests import). Metrics are analytics over predictions that already left the device once per
epoch; argmax uses NumPy's first-index tie break, like the reference (metrics.py:89)."""
from __future__ import annotations

import numpy as np


def _reshape_inputs(y_pred, y_true):
    y_pred, y_true = np.asarray(y_pred), np.asarray(y_true)
    if y_pred.ndim == 1:
        y_pred = y_pred.reshape(-1, 1)
    if y_true.ndim == 1:
        y_true = y_true.reshape(-1, 1)
    return y_pred, y_true


def _classes(y_pred, y_true, threshold):
    y_pred, y_true = _reshape_inputs(y_pred, y_true)
    if y_pred.shape[1] == 1:
        return (y_pred >= threshold).astype(int).ravel(), y_true.astype(int).ravel()
    return np.argmax(y_pred, axis=1), np.argmax(y_true, axis=1)


def accuracy_score(y_pred, y_true, threshold: float = 0.5) -> float:
    y_pred, y_true = _reshape_inputs(y_pred, y_true)
    if y_pred.shape[1] == 1:
        return np.mean((y_pred >= threshold).astype(int) == y_true)
    return np.mean(np.argmax(y_pred, axis=1) == np.argmax(y_true, axis=1))


def confusion_matrix(y_pred, y_true, threshold: float = 0.5) -> np.ndarray:
    p, t = _classes(y_pred, y_true, threshold)
    y_pred2, _ = _reshape_inputs(y_pred, y_true)
    n = 2 if y_pred2.shape[1] == 1 else y_pred2.shape[1]
    cm = np.zeros((n, n), dtype=int)
    np.add.at(cm, (t, p), 1)
    return cm


def _per_class(y_pred, y_true, threshold):
    cm = confusion_matrix(y_pred, y_true, threshold)
    tp = np.diag(cm).astype(float)
    return tp, cm.sum(axis=0) - tp, cm.sum(axis=1) - tp   # tp, fp, fn


def precision_score(y_pred, y_true, threshold: float = 0.5) -> float:
    tp, fp, _ = _per_class(y_pred, y_true, threshold)
    with np.errstate(divide="ignore", invalid="ignore"):
        v = np.where(tp + fp > 0, tp / (tp + fp), 0.0)
    return float(np.mean(v))


def recall_score(y_pred, y_true, threshold: float = 0.5) -> float:
    tp, _, fn = _per_class(y_pred, y_true, threshold)
    with np.errstate(divide="ignore", invalid="ignore"):
        v = np.where(tp + fn > 0, tp / (tp + fn), 0.0)
    return float(np.mean(v))


def f1_score(y_pred, y_true, threshold: float = 0.5) -> float:
    p, r = precision_score(y_pred, y_true, threshold), recall_score(y_pred, y_true, threshold)
    return 2 * p * r / (p + r) if (p + r) > 0 else 0.0


def mean_squared_error(y_pred, y_true, threshold: float = 0.5) -> float:
    y_pred, y_true = _reshape_inputs(y_pred, y_true)
    return float(np.mean((y_pred - y_true) ** 2))


def mean_absolute_error(y_pred, y_true, threshold: float = 0.5) -> float:
    y_pred, y_true = _reshape_inputs(y_pred, y_true)
    return float(np.mean(np.abs(y_pred - y_true)))


_BY_NAME = {}
for _names, _fn in ((("accuracy", "accuracy_score", "accuracy-score", "acc"), accuracy_score),
                    (("f1", "f1_score", "f1-score"), f1_score),
                    (("recall", "recall_score", "recall-score", "sensitivity", "rec"), recall_score),
                    (("precision", "precision_score", "precision-score", "positive-predictive-value"), precision_score),
                    (("mean-squared-error", "mse"), mean_squared_error),
                    (("mean-absolute-error", "mae"), mean_absolute_error)):
    for _n in _names:
        _BY_NAME[_n] = _fn


class Metric:
    def __init__(self, name):
        if isinstance(name, str):
            if name not in _BY_NAME:
                raise ValueError(f"Metric {name} is not supported.")
            self.function = _BY_NAME[name]
        elif callable(name):
            self.function = name
        else:
            raise ValueError(f"Metric {name} is not supported.")
        self.name = self.function.__name__.split("_score")[0]

    def __call__(self, y_pred, y_true, threshold: float = 0.5) -> float:
        y_pred, y_true = _reshape_inputs(y_pred, y_true)
        return self.function(y_pred, y_true, threshold)
