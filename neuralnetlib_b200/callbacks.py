"""Callbacks that must keep working against device-resident weights (callbacks.py:7-374):
they only touch `layer.weights` / `layer.bias` (read = device->host copy, assignment = in-place upload that keeps
captured graphs valid) and `optimizer.learning_rate` (a device scalar behind a property). No kernels here."""
from __future__ import annotations

import numpy as np

from .metrics import Metric


class ModelWeightManager:
    @staticmethod
    def get_model_weights(model):
        out = []
        for layer in getattr(model, "layers", []):
            if hasattr(layer, "weights") and layer.weights is not None:
                out.append((layer.weights.copy(), layer.bias.copy() if getattr(layer, "bias", None) is not None else None))
        return out

    @staticmethod
    def set_model_weights(model, params) -> None:
        it = iter(params)
        for layer in getattr(model, "layers", []):
            if hasattr(layer, "weights") and layer.weights is not None:
                try:
                    w, b = next(it)
                except StopIteration:
                    return
                layer.weights = w.copy()
                if b is not None:
                    layer.bias = b.copy()


class Callback:
    def on_train_begin(self, logs=None): pass
    def on_train_end(self, logs=None): pass
    def on_epoch_begin(self, epoch, logs=None): pass
    def on_epoch_end(self, epoch, logs=None): pass
    def on_batch_begin(self, batch, logs=None): pass
    def on_batch_end(self, batch, logs=None): pass


class EarlyStopping(Callback):
    """Same knobs and decisions as callbacks.py:156-264."""

    def __init__(self, patience: int = 5, min_delta: float = 0.001, restore_best_weights: bool = True,
                 start_from_epoch: int = 0, monitor: str = "loss", mode: str = "auto", baseline: float | None = None):
        self.patience, self.min_delta, self.restore_best_weights = patience, min_delta, restore_best_weights
        self.start_from_epoch, self.mode, self.baseline = start_from_epoch, mode, baseline
        if monitor in ("loss", "val_loss"):
            self.monitor, self.is_metric = monitor, False
        else:
            try:
                self.monitor, self.is_metric = Metric(monitor), True
            except ValueError as e:
                if "val_" not in monitor:
                    raise ValueError(f"Invalid monitor metric: {monitor}") from e
                try:
                    Metric(monitor.replace("val_", ""))
                except ValueError:
                    raise ValueError(f"Invalid monitor metric: {monitor}") from e
                self.monitor, self.is_metric = monitor, False
        self.best_weights, self.best_metric = None, None
        self.patience_counter, self.stop_training = 0, False
        self.weight_manager = ModelWeightManager()

    def on_train_begin(self, logs=None):
        self.patience_counter, self.best_metric, self.stop_training = 0, None, False

    def _name(self):
        return self.monitor if isinstance(self.monitor, str) else self.monitor.name

    def _value(self, logs):
        v = logs.get(self._name())
        if v is None:
            if "loss" not in logs:
                raise ValueError(f"Monitored metric '{self._name()}' is not available. Available metrics are: "
                                 f"{', '.join(logs.keys())}")
            v = logs["loss"]
        return float(v)

    def on_epoch_end(self, epoch, logs=None) -> bool:
        logs = logs or {}
        model = logs.get("model")
        if epoch < self.start_from_epoch or model is None:
            return False
        cur = self._value(logs)
        if self.baseline is not None and self.best_metric is None:
            if (self.mode == "min" and cur > self.baseline) or (self.mode == "max" and cur < self.baseline):
                print(f"\nEarly stopping: baseline {self.baseline} was not met.")
                return True
        if self.best_metric is None:
            self.best_metric = cur
            if self.mode == "auto":
                self.mode = "min" if "loss" in self._name().lower() else "max"
        improved = cur < self.best_metric - self.min_delta if self.mode == "min" else cur > self.best_metric + self.min_delta
        if improved:
            self.best_metric, self.patience_counter = cur, 0
            if self.restore_best_weights:
                self.best_weights = self.weight_manager.get_model_weights(model)
        else:
            self.patience_counter += 1
        if self.patience_counter >= self.patience:
            self.stop_training = True
            if self.restore_best_weights and self.best_weights is not None:
                self.weight_manager.set_model_weights(model, self.best_weights)
            print(f"\nEarly stopping triggered after epoch {epoch + 1}")
            return True
        return False


class LearningRateScheduler(Callback):
    """Sets `optimizer.learning_rate` at the start of every epoch from `schedule(epoch, lr)` or a named rule
    ('step', 'exponential'); the optimizer's property pushes the value to the device scalar the kernels read."""

    def __init__(self, schedule="step", initial_learning_rate: float | None = None, decay_rate: float = 0.5,
                 decay_steps: int = 10, verbose: bool = False, **kwargs):
        self.schedule, self.initial_learning_rate = schedule, initial_learning_rate
        self.decay_rate, self.decay_steps, self.verbose = decay_rate, decay_steps, verbose

    def _lr(self, epoch, lr0):
        if callable(self.schedule):
            return self.schedule(epoch, lr0)
        if self.schedule == "step":
            return lr0 * self.decay_rate ** (epoch // self.decay_steps)
        if self.schedule == "exponential":
            return lr0 * np.exp(-self.decay_rate * epoch / max(self.decay_steps, 1))
        raise ValueError(f"unknown schedule {self.schedule}")

    def on_epoch_begin(self, epoch, logs=None):
        model = (logs or {}).get("model")
        opt = getattr(model, "optimizer", None)
        if opt is None:
            return
        if self.initial_learning_rate is None:
            self.initial_learning_rate = opt.learning_rate
        opt.learning_rate = float(self._lr(epoch, self.initial_learning_rate))
        if self.verbose:
            print(f"\nEpoch {epoch + 1}: learning rate {opt.learning_rate:.6g}")
