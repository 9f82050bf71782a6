"""Mirror of neuralnetlib/losses.py:1-180 for the hot path: MeanSquaredError, BinaryCrossentropy, CategoricalCrossentropy,
SparseCategoricalCrossentropy, MeanAbsoluteError, Huber (value + derivative on the device through b200nn_loss_fwd_bwd_ex)."""
from __future__ import annotations

import numpy as np
import torch

from . import _backend as B
from . import _ops as ops


class LossFunction:
    EPSILON = 1e-15
    _kind = None

    def __call__(self, y_true: np.ndarray, y_pred: np.ndarray) -> float:
        raise NotImplementedError

    def derivative(self, y_true: np.ndarray, y_pred: np.ndarray) -> np.ndarray:
        raise NotImplementedError

    def get_config(self) -> dict:
        return {"name": self.__class__.__name__}

    @staticmethod
    def from_config(config: dict) -> "LossFunction":
        for cls in LossFunction.__subclasses__():
            if cls.__name__ == config["name"]:
                return cls(**{k: v for k, v in config.items() if k != "name"})

    @staticmethod
    def from_name(name: str) -> "LossFunction":
        aliases = {"mse": "meansquarederror", "bce": "binarycrossentropy", "cce": "categoricalcrossentropy",
                   "scce": "sparsecategoricalcrossentropy", "mae": "meanabsoluteerror", "huber": "huber"}
        key = name.lower().replace("_", "")
        key = aliases.get(key, key)
        for cls in LossFunction.__subclasses__():
            if cls.__name__.lower() == key:
                return cls()
        raise ValueError(f"No loss function found for the name: {name}")

    # host-array contract of the reference (losses.py:79-120), computed on the device
    def _run(self, y_true, y_pred, want_err):
        p = np.asarray(y_pred, dtype=np.float64)
        y = np.asarray(y_true, dtype=np.float64)
        shape = p.shape
        p2 = p.reshape(shape[0], -1) if p.ndim > 1 else p.reshape(-1, 1)
        y2 = np.broadcast_to(y.reshape(p2.shape) if y.size == p2.size else y, p2.shape)
        pd, yd = B.to_device(p2), B.to_device(np.ascontiguousarray(y2))
        err = B.empty(pd.shape)
        acc = B.zeros((4,), dtype=torch.float64)
        ops.loss_fwd_bwd(self._kind, 0, pd, yd, err, acc, None, None, 1, float(getattr(self, "_param", 1.0)))
        if want_err:
            return B.to_host(err).reshape(shape)
        return float(acc.cpu()[0].item()) / self._denominator(p2.shape)

    def _denominator(self, shape):
        return shape[0] * shape[1]


class MeanSquaredError(LossFunction):        # losses.py:79-87
    _kind = ops.LOSS_MSE

    def __call__(self, y_true, y_pred):
        return self._run(y_true, y_pred, False)

    def derivative(self, y_true, y_pred):
        return self._run(y_true, y_pred, True)

    def __str__(self):
        return "MeanSquaredError"


class BinaryCrossentropy(LossFunction):      # losses.py:90-102
    _kind = ops.LOSS_BCE

    def __call__(self, y_true, y_pred):
        return self._run(y_true, y_pred, False)

    def derivative(self, y_true, y_pred):
        return self._run(y_true, y_pred, True)

    def __str__(self):
        return "BinaryCrossentropy"


class CategoricalCrossentropy(LossFunction):  # losses.py:105-120 (sum / batch)
    _kind = ops.LOSS_CCE

    def __call__(self, y_true, y_pred):
        return self._run(y_true, y_pred, False)

    def derivative(self, y_true, y_pred):
        return self._run(y_true, y_pred, True)

    def _denominator(self, shape):
        return shape[0]

    def __str__(self):
        return "CategoricalCrossentropy"


class SparseCategoricalCrossentropy(LossFunction):   # losses.py:123-143: integer labels; CCE on one-hot rows
    _kind = ops.LOSS_CCE
    _sparse = True

    @staticmethod
    def one_hot(y_true, n_classes):
        y = np.asarray(y_true).astype(np.int64).reshape(-1)
        out = np.zeros((y.shape[0], int(n_classes)), dtype=np.float32)
        out[np.arange(y.shape[0]), y] = 1.0
        return out

    def __call__(self, y_true, y_pred):
        return self._run(self.one_hot(y_true, np.asarray(y_pred).shape[1]), y_pred, False)

    def derivative(self, y_true, y_pred):
        return self._run(self.one_hot(y_true, np.asarray(y_pred).shape[1]), y_pred, True)

    def _denominator(self, shape):
        return shape[0]

    def __str__(self):
        return "SparseCategoricalCrossentropy"


class MeanAbsoluteError(LossFunction):       # losses.py:146-155
    _kind = ops.LOSS_MAE

    def __call__(self, y_true, y_pred):
        return self._run(y_true, y_pred, False)

    def derivative(self, y_true, y_pred):
        return self._run(y_true, y_pred, True)

    def __str__(self):
        return "MeanAbsoluteError"


class Huber(LossFunction):                   # losses.py:158-180
    _kind = ops.LOSS_HUBER

    def __init__(self, delta: float = 1.0):
        super().__init__()
        self.delta = delta

    @property
    def _param(self):
        return float(self.delta)

    def __call__(self, y_true, y_pred):
        return self._run(y_true, y_pred, False)

    def derivative(self, y_true, y_pred):
        return self._run(y_true, y_pred, True)

    def __str__(self):
        return f"HuberLoss(delta={self.delta})"

    def get_config(self) -> dict:
        return {"name": self.__class__.__name__, "delta": self.delta}
