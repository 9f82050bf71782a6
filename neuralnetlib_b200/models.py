"""Mirror of neuralnetlib/models.py on the device: Sequential (:81-721) and GAN.train_on_batch/fit core (:2366-2936).

The Python surface is the reference's (add/compile/fit/train_on_batch/forward_pass/backward_pass/evaluate/predict/
save/load; same arguments, defaults, exceptions); the per-batch work runs as one captured CUDA graph over
HBM-resident buffers (engine.Plan). NumPy appears only at the entry points and when an attribute is read.
"""
from __future__ import annotations

import json
import time

import numpy as np
import torch

from . import _backend as B
from . import _ops as ops
from .activations import ActivationFunction
from .engine import Plan
from .layers import *  # noqa: F401,F403  (the reference re-exports the layers from models)
from .layers import Activation, Dense, Input, Layer, incompatibility_dict
from .losses import BinaryCrossentropy, CategoricalCrossentropy, LossFunction  # noqa: F401
from .metrics import Metric
from .optimizers import Optimizer
from .utils import History, format_number, progress_bar, shuffle_indices, train_test_split

_RESIDENT_DATASET_BYTES = 8 << 30   # datasets up to 8 GiB are kept in HBM for the whole fit()


class BaseModel:
    def __init__(self, gradient_clip_threshold: float = 5.0, enable_padding: bool = False, padding_size: int = 32,
                 random_state: int | None = None):
        self.gradient_clip_threshold = gradient_clip_threshold   # unused by Sequential in the reference too
        self.enable_padding = enable_padding
        self.padding_size = padding_size
        self.random_state = random_state if random_state is not None else time.time_ns()


class Sequential(BaseModel):
    def __init__(self, gradient_clip_threshold: float = 5.0, enable_padding: bool = False, padding_size: int = 32,
                 random_state: int | None = None, n_classes: int | None = None):
        super().__init__(gradient_clip_threshold, enable_padding, padding_size, random_state)
        self.layers = []
        self.loss_function = None
        self.optimizer = None
        self.y_true = None
        self._initialized = False
        self.n_classes = n_classes
        self.metrics = None
        self._plans = {}
        self._last_plan = None
        self._pred_host = None
        self._dp = None   # parallel.DataParallel once make_data_parallel() has been called

    # ------------------------------------------------------------------------------------------------------------
    def __str__(self) -> str:
        s = (f"Sequential(gradient_clip_threshold={self.gradient_clip_threshold}, enable_padding={self.enable_padding}, "
             f"padding_size={self.padding_size}, random_state={self.random_state})\n")
        s += "-------------------------------------------------\n"
        for i, layer in enumerate(self.layers):
            s += f"Layer {i + 1}: {str(layer)}\n"
        s += "-------------------------------------------------\n"
        s += f"Loss function: {str(self.loss_function)}\nOptimizer: {str(self.optimizer)}\n"
        s += "-------------------------------------------------\n"
        return s

    def summary(self):
        print(str(self))

    def add(self, layer: Layer):
        """models.py:111-148."""
        if not self.layers:
            if not isinstance(layer, Input):
                raise ValueError("The first layer must be an Input layer.")
        else:
            if self.random_state and hasattr(layer, "random_state"):
                layer.random_state = self.random_state
            if isinstance(layer, Dense) and isinstance(self.layers[0], Input):
                layer.input_dim = self.layers[0].input_dim
            prev_t, cur_t = type(self.layers[-1]), type(layer)
            if cur_t in incompatibility_dict.get(prev_t, []):
                raise ValueError(f"{cur_t.__name__} layer cannot follow {prev_t.__name__} layer.")
        self.layers.append(layer)
        act = getattr(layer, "activation", getattr(layer, "activation_function", None))
        if act and not isinstance(layer, Activation):
            if isinstance(act, str):
                self.layers.append(Activation.from_name(act))
            elif isinstance(act, ActivationFunction):
                self.layers.append(Activation(act))
            elif isinstance(act, Activation):
                self.layers.append(act)
            else:
                raise ValueError(f"Invalid activation function: {act}")
        self._drop_plans()

    def compile(self, loss_function, optimizer, verbose: bool = False, metrics: list | None = None):
        self.loss_function = loss_function if isinstance(loss_function, LossFunction) else LossFunction.from_name(loss_function)
        self.optimizer = optimizer if isinstance(optimizer, Optimizer) else Optimizer.from_name(optimizer)
        self.metrics = metrics
        self._drop_plans()
        if verbose:
            print(str(self))

    # ------------------------------------------------------------------------------------------------------------
    def _drop_plans(self):
        for p in self._plans.values():
            p.destroy()
        self._plans = {}
        self._last_plan = None

    def __del__(self):
        # a model that goes out of scope must not leave captured graphs behind: under data parallelism they reference the NCCL
        # communicator, whose destruction waits for them
        try:
            self._drop_plans()
        except Exception:   # noqa: BLE001  (interpreter shutdown: the library may be gone already)
            pass

    def _plan(self, batch_shape, training: bool) -> Plan:
        from . import math_version
        if getattr(self, "_plans_math", None) != math_version():   # set_math() after the first step: the mode is frozen into a plan
            self._drop_plans()
            self._plans_math = math_version()
        key = (tuple(int(s) for s in batch_shape), bool(training))
        p = self._plans.get(key)
        if p is None:
            p = Plan(self, key[0], key[1])
            self._plans[key] = p
        return p

    @property
    def predictions(self):
        """Output of the last forward pass as a float64 array (device->host copy on first read)."""
        if self._pred_host is None and self._last_plan is not None:
            self._pred_host = B.to_host(self._last_plan.pred)[:self._last_rows]
        return self._pred_host

    @predictions.setter
    def predictions(self, v):
        self._pred_host = v

    def _prepare_input(self, X, labels):
        if self.n_classes is not None and labels is not None:
            X = np.asarray(X)
            if X.ndim != 2:
                raise ValueError("Input shape must be (batch_size, features) for conditional models")
            labels = np.asarray(labels)
            if labels.ndim == 1:
                labels = np.eye(self.n_classes)[labels]
            elif labels.shape[1] != self.n_classes:
                raise ValueError(f"Labels must have {self.n_classes} classes")
            X = np.concatenate([X, labels], axis=1)
        rows = X.shape[0]
        if self.enable_padding and rows % self.padding_size:
            X = np.asarray(X) if not isinstance(X, torch.Tensor) else X.cpu().numpy()
            padded = np.zeros(((rows + self.padding_size - 1) // self.padding_size * self.padding_size,) + X.shape[1:], dtype=X.dtype)
            padded[:rows] = X
            X = padded
        return X, rows

    def _forward_dev(self, X, training: bool = True, labels=None) -> torch.Tensor:
        X, rows = self._prepare_input(X, labels)
        shape = tuple(X.shape) if len(X.shape) > 1 else (1, int(X.shape[0]))
        plan = self._plan(shape, training)
        plan.set_input(X)
        plan.program("forward").run()
        self._last_plan, self._last_rows, self._pred_host = plan, rows, None
        return plan.pred[:rows]

    def forward_pass(self, X, training: bool = True, labels=None):
        """models.py:159-193. NumPy in -> float64 NumPy out; device tensor in -> device tensor out."""
        out = self._forward_dev(X, training, labels)
        if isinstance(X, torch.Tensor) and X.is_cuda:
            return out
        return self.predictions

    def backward_pass(self, error, gan: bool = False, compute_only: bool = False):
        """models.py:195-249 on the plan of the last forward pass. Returns the error w.r.t. the model input."""
        plan = self._last_plan
        if plan is None:
            raise RuntimeError("backward_pass requires a preceding forward_pass")
        dev_in = isinstance(error, torch.Tensor) and error.is_cuda
        kind = "backward_gan" if gan else ("backward_compute_only" if compute_only else "backward")
        if gan and compute_only:
            raise NotImplementedError("gan=True together with compute_only=True is not used by the reference")
        prog = plan.program(kind)
        Plan._upload(plan.err_in, "_err_stage", plan, error if dev_in else np.asarray(error).reshape(tuple(plan.err_in.shape)))
        if self.y_true is not None and plan.y_in is not None:
            plan.set_target(self.y_true if isinstance(self.y_true, torch.Tensor) else np.asarray(self.y_true).reshape(tuple(plan.y_in.shape)))
        prog.run()
        return plan.in_err_out if dev_in else B.to_host(plan.in_err_out)

    # ------------------------------------------------------------------------------------------------------------
    def _loss_denominator(self, plan, after_train: bool = False):
        rows = plan.pred.shape[0]
        cols = plan.pred.numel() // rows
        if after_train:
            rows *= plan.world   # data parallel: a train step leaves the RANK SUM of the numerators in plan.acc
        by_rows = isinstance(self.loss_function, CategoricalCrossentropy) or getattr(self.loss_function, "_sparse", False)
        return rows if by_rows else rows * cols

    def _dense_targets(self, y, n_out):
        """SparseCategoricalCrossentropy takes integer labels (losses.py:123-143): one-hot rows for the device loss kernel."""
        if getattr(self.loss_function, "_sparse", False):
            y = np.asarray(y)
            if y.ndim == 1 or (y.ndim == 2 and y.shape[1] == 1 and n_out > 1):
                return self.loss_function.one_hot(y, n_out)
        return y

    def _train_step(self, x_batch, y_batch) -> Plan:
        """One fused step on the device; does not synchronise (the loss numerator stays in plan.acc[0])."""
        shape = tuple(x_batch.shape)
        plan = self._plan(shape, True)
        prog = plan.program("train")
        plan.set_input(x_batch)
        plan.set_target(self._dense_targets(y_batch, plan.pred.shape[-1]))   # same element count as the prediction buffer: uploaded flat
        prog.run()
        self._last_plan, self._last_rows, self._pred_host = plan, shape[0], None
        return plan

    def train_on_batch(self, x_batch, y_batch) -> float:
        """models.py:251-264: forward, loss, backward with clips, optimizer updates; returns the loss as a float
        (the only device->host transfer of the step)."""
        if self.loss_function is None or self.optimizer is None:
            raise ValueError("The model must be compiled before training")
        self.y_true = y_batch
        plan = self._train_step(x_batch, y_batch)
        return plan.read_acc()[0] / self._loss_denominator(plan, True)

    # ------------------------------------------------------------------------------------------------------------
    def fit(self, x_train, y_train, epochs: int, batch_size: int | None = None, verbose: bool = True,
            metrics: list | None = None, random_state: int | None = None, validation_data: tuple | None = None,
            validation_split: float | None = None, callbacks: list = [], plot_decision_boundary: bool = False) -> dict:
        """models.py:266-563 with the dataset resident in HBM: the epoch permutation (same generator as utils.shuffle)
        is uploaded once per epoch and every mini-batch is gathered on the device; losses and predictions are
        accumulated on the device and read back once per epoch unless a callback / verbose output needs them."""
        if plot_decision_boundary:
            raise NotImplementedError("plot_decision_boundary is outside the B200 hot path")
        if getattr(self, "metrics", None) is not None:
            metrics = self.metrics
        history = History({"loss": [], "val_loss": []})
        if validation_split is not None and validation_data is not None:
            raise ValueError("Cannot specify both validation_data and validation_split")
        seed = random_state if random_state is not None else self.random_state
        if validation_split is not None:
            x_train, x_val, y_train, y_val = train_test_split(x_train, y_train, test_size=validation_split, random_state=seed)
            validation_data = (x_val, y_val)
        x_train = np.asarray(x_train)
        y_train = np.asarray(y_train)
        for layer in self.layers:
            if hasattr(layer, "random_state"):
                layer.random_state = seed
        if validation_data is not None:
            x_val, y_val = np.asarray(validation_data[0]), np.asarray(validation_data[1])
        mobjs = None
        if metrics is not None:
            mobjs = []
            for m in metrics:
                if m == "val_loss":
                    continue
                try:
                    mo = Metric(m)
                except ValueError as e:
                    raise ValueError(f"Invalid metric: {m}") from e
                mobjs.append(mo)
                history[mo.name] = []
                if validation_data is not None:
                    history[f"val_{mo.name}"] = []
        callbacks = callbacks if callbacks is not None else []
        logs = {"model": self, "params": {"epochs": epochs, "batch_size": batch_size, "verbose": verbose,
                                          "metrics": [m.name for m in (mobjs or [])], "validation": validation_data is not None}}
        for cb in callbacks:
            cb.on_train_begin(logs)

        N = x_train.shape[0]
        bs = int(batch_size) if batch_size is not None else N
        dp = self._dp
        if dp is not None:
            from .parallel import shard_rows   # every rank holds the full dataset and trains on its rows of each batch
        y_metric = y_train.reshape(N, -1) if y_train.ndim > 1 else y_train.reshape(N, 1)   # what the metrics compare against
        if getattr(self.loss_function, "_sparse", False):   # integer labels -> one-hot rows for the device loss kernel
            n_out = self._plan((min(bs, N),) + x_train.shape[1:], True).pred.shape[-1]
            y_train = np.asarray(self._dense_targets(y_train, n_out))
        y2 = y_train.reshape(N, -1) if y_train.ndim > 1 else y_train.reshape(N, 1)
        resident = (x_train.size + y2.size) * 4 <= _RESIDENT_DATASET_BYTES
        if resident:
            x_dev, y_dev = B.to_device(x_train), B.to_device(y2)
        num_batches = (N + bs - 1) // bs
        sync_each_batch = bool(callbacks) or verbose
        loss_hist = B.zeros((num_batches,), dtype=torch.float64)
        keep_pred = mobjs is not None
        try:
            for epoch in range(epochs):
                epoch_logs = {"model": self}
                for cb in callbacks:
                    cb.on_epoch_begin(epoch, epoch_logs)
                start = time.time()
                perm = shuffle_indices(N, seed)                      # models.py:386-389: same seed => same permutation
                y_shuf = y_metric[perm] if keep_pred else None
                perm_dev = B.to_device(perm, dtype=torch.int64) if resident else None
                pred_all = None
                running = 0.0
                denoms = []
                seen = []   # rows (of the permuted dataset) whose predictions this rank holds
                for bi, j in enumerate(range(0, N, bs)):
                    rows = min(bs, N - j)
                    batch_logs = {"batch": bi, "size": rows, "model": self}
                    for cb in callbacks:
                        cb.on_batch_begin(bi, batch_logs)
                    j0 = j
                    if dp is not None:
                        lo, hi = shard_rows(rows, dp.world, dp.rank)
                        j0, rows = j + lo, hi - lo
                    plan = self._plan((rows,) + x_train.shape[1:], True)
                    prog = plan.program("train")
                    if resident:
                        ops.gather_rows(x_dev, perm_dev[j0:j0 + rows], plan.x_in)
                        ops.gather_rows(y_dev, perm_dev[j0:j0 + rows], plan.y_in)
                    else:
                        plan.set_input(x_train[perm[j0:j0 + rows]])
                        plan.set_target(y2[perm[j0:j0 + rows]].reshape(tuple(plan.y_in.shape)))
                    prog.run()
                    self._last_plan, self._last_rows, self._pred_host = plan, rows, None
                    j = j0   # predictions / metrics below are those of this rank's rows
                    loss_hist[bi:bi + 1].copy_(plan.acc[0:1], non_blocking=True)
                    denoms.append(self._loss_denominator(plan, True))
                    if keep_pred:
                        if pred_all is None:
                            pred_all = B.empty((N, plan.pred.numel() // rows))
                        pred_all[j:j + rows].copy_(plan.pred.reshape(rows, -1), non_blocking=True)
                        seen.append((j, j + rows))
                    if sync_each_batch:
                        batch_loss = float(plan.acc[0].item()) / denoms[-1]
                        running += batch_loss
                        batch_logs["loss"] = batch_loss
                        if mobjs is not None:
                            pb = B.to_host(plan.pred.reshape(rows, -1))
                            for mo in mobjs:
                                batch_logs[mo.name] = mo(pb, y_shuf[j:j + rows])
                        for cb in callbacks:
                            cb.on_batch_end(bi, batch_logs)
                        if verbose:
                            msg = f"Epoch {epoch + 1}/{epochs} - {time.time() - start:.2f}s - loss: {format_number(running / (bi + 1))}"
                            progress_bar(bi + 1, num_batches, message=msg)
                losses = loss_hist.cpu().numpy() / np.asarray(denoms, dtype=np.float64)   # one read-back per epoch
                error = float(np.sum(losses) / num_batches)
                history["loss"].append(error)
                epoch_logs.update({"loss": error, "time": time.time() - start})
                if mobjs is not None:
                    preds, y_seen = B.to_host(pred_all), y_shuf
                    if dp is not None:   # metrics over this rank's rows only (the loss above is global)
                        sel = np.concatenate([np.arange(a, b) for a, b in seen])
                        preds, y_seen = preds[sel], y_shuf[sel]
                    for mo in mobjs:
                        v = mo(preds, y_seen)
                        history[mo.name].append(v)
                        epoch_logs[mo.name] = v
                if validation_data is not None:
                    val_loss, val_pred = self.evaluate(x_val, y_val, bs)
                    history["val_loss"].append(val_loss)
                    epoch_logs["val_loss"] = val_loss
                    if mobjs is not None:
                        for mo in mobjs:
                            v = mo(val_pred, y_val)
                            history[f"val_{mo.name}"].append(v)
                            epoch_logs[f"val_{mo.name}"] = v
                if verbose:
                    extra = "".join(f" - {k}: {format_number(v)}" for k, v in epoch_logs.items() if k not in ("model", "time"))
                    progress_bar(num_batches, num_batches, message=f"Epoch {epoch + 1}/{epochs} - {epoch_logs['time']:.2f}s{extra}")
                    print()
                stop = False
                for cb in callbacks:
                    if cb.on_epoch_end(epoch, epoch_logs):
                        stop = True
                        break
                if stop:
                    break
        finally:
            for cb in callbacks:
                cb.on_train_end({"model": self, "history": history})
            if verbose:
                print()
        return history

    # ------------------------------------------------------------------------------------------------------------
    def evaluate(self, x_test, y_test, batch_size: int = 32) -> tuple:
        """models.py:565-601: inference-mode forward + loss per batch; (mean of batch losses, all predictions)."""
        x_test, y_test = np.asarray(x_test), np.asarray(y_test)
        n = len(x_test)
        batch_size = int(batch_size) if batch_size else n
        total, preds, nb = 0.0, [], 0
        for i in range(0, n, batch_size):
            xb, yb = x_test[i:i + batch_size], y_test[i:i + batch_size]
            out = self._forward_dev(xb, training=False)
            plan = self._last_plan
            if plan.pred.shape[0] != len(xb):
                # enable_padding: the plan ran on zero-padded rows; the reference slices the predictions back before the loss
                # (models.py:190-192), so the loss is taken over the real rows only (host contract of the loss, device kernel inside)
                ph = B.to_host(out)[:len(xb)]
                total += float(self.loss_function(np.asarray(self._dense_targets(yb, plan.pred.shape[-1])).reshape(ph.shape), ph))
                preds.append(ph)
                nb += 1
                continue
            plan.set_target(np.asarray(self._dense_targets(yb, plan.pred.shape[-1])).reshape(tuple(plan.pred.shape)))
            plan.program("loss").run()
            total += float(plan.acc[0].item()) / self._loss_denominator(plan)
            preds.append(B.to_host(out))
            nb += 1
        return total / max(nb, 1), np.vstack(preds)

    def predict(self, X, temperature: float = 1.0) -> np.ndarray:
        """models.py:603-615."""
        X = np.array(X)
        self._forward_dev(X, training=False)
        predictions = self.predictions
        if not np.isclose(temperature, 1.0, rtol=1e-09, atol=1e-09):
            predictions = np.exp(np.log(np.clip(predictions, 1e-7, 1.0)) / temperature)
            predictions /= np.sum(predictions, axis=-1, keepdims=True)
        return predictions

    # ------------------------------------------------------------------------------------------------------------
    def save(self, filename: str):
        """models.py:674-696 — same JSON layout (weights as nested lists, optimizer state incl. Adam t/m/v)."""
        state = {"type": "Sequential", "layers": [layer.get_config() for layer in self.layers],
                 "gradient_clip_threshold": self.gradient_clip_threshold, "enable_padding": self.enable_padding,
                 "padding_size": self.padding_size, "random_state": self.random_state, "n_classes": self.n_classes}
        if self.loss_function:
            state["loss_function"] = self.loss_function.get_config()
        if self.optimizer:
            state["optimizer"] = self.optimizer.get_config()
        with open(filename, "w") as f:
            json.dump(state, f, indent=4, default=_json_default)
        return state

    @classmethod
    def load(cls, filename: str) -> "Sequential":
        with open(filename, "r") as f:
            state = json.load(f)
        model = cls()
        for k in ("gradient_clip_threshold", "enable_padding", "padding_size", "random_state", "n_classes"):
            if k in state:
                setattr(model, k, state[k])
        model.layers = [Layer.from_config(c) for c in state.get("layers", [])]
        if "loss_function" in state:
            model.loss_function = LossFunction.from_config(state["loss_function"])
        if "optimizer" in state:
            model.optimizer = Optimizer.from_config(state["optimizer"])
        return model


def _json_default(o):
    if isinstance(o, (np.integer,)):
        return int(o)
    if isinstance(o, (np.floating,)):
        return float(o)
    if isinstance(o, np.ndarray):
        return o.tolist()
    raise TypeError(f"not JSON serialisable: {type(o)}")


# ================================================================================================================
class GAN(BaseModel):
    """models.py:2366-2936, hot path only: compile, train_on_batch (D step on real+fake, G step through the frozen D
    with compute_only / gan=True) and the fit loop around it. Plotting, metrics on generated samples and
    save/load are out of scope. Both networks stay on the device for the whole step."""

    def __init__(self, latent_dim: int = 100, n_classes: int | None = None, gradient_clip_threshold: float = 0.1,
                 enable_padding: bool = False, padding_size: int = 32, random_state: int | None = None,
                 use_spectral_norm: bool = True, use_gradient_penalty: bool = True, gp_weight: float = 10.0,
                 image_height: int | None = None, image_width: int | None = None, label_smoothing: float = 0.9):
        super().__init__(gradient_clip_threshold, enable_padding, padding_size, random_state)
        if n_classes is not None:
            raise NotImplementedError("conditional GANs are outside the B200 hot path")
        self.latent_dim, self.n_classes = latent_dim, n_classes
        self.generator = self.discriminator = None
        self.generator_optimizer = self.discriminator_optimizer = None
        self.generator_loss = self.discriminator_loss = None
        # spectral norm / gradient penalty are never applied by the reference's train_on_batch
        self.use_spectral_norm, self.use_gradient_penalty, self.gp_weight = use_spectral_norm, use_gradient_penalty, gp_weight
        self._image_height, self._image_width = image_height, image_width
        self.label_smoothing = label_smoothing
        self._train_step = 0
        self._dp = None   # parallel.make_gan_data_parallel

    def compile(self, generator: Sequential, discriminator: Sequential, generator_optimizer, discriminator_optimizer,
                loss_function="bce", verbose: bool = False, metrics: list | None = None):
        self.generator, self.discriminator = generator, discriminator
        if not any(isinstance(layer, Dense) for layer in generator.layers):
            raise ValueError("The generator must contain at least one Dense layer.")
        as_opt = lambda o: o if isinstance(o, Optimizer) else Optimizer.from_name(o)  # noqa: E731
        as_loss = lambda l: l if isinstance(l, LossFunction) else LossFunction.from_name(l)  # noqa: E731
        self.generator_optimizer, self.discriminator_optimizer = as_opt(generator_optimizer), as_opt(discriminator_optimizer)
        self.generator_loss, self.discriminator_loss = as_loss(loss_function), as_loss(loss_function)
        generator.loss_function, generator.optimizer = self.generator_loss, self.generator_optimizer
        discriminator.loss_function, discriminator.optimizer = self.discriminator_loss, self.discriminator_optimizer
        generator._drop_plans()
        discriminator._drop_plans()
        self.metrics = metrics
        self.last_activation = generator.layers[-1].activation_function.get_config()["name"].lower()
        if verbose:
            print(f"GAN(latent_dim={self.latent_dim})")

    def forward_pass(self, latent_vectors, training: bool = True):
        if self.generator is None:
            raise ValueError("Model must be compiled before forward pass")
        return self.generator.forward_pass(latent_vectors, training)

    def _generate_latent_points(self, n_samples: int, labels=None):
        """models.py:2522-2565: re-seeded identically on every call."""
        rng = np.random.default_rng(self.random_state)
        return rng.normal(0, 1, (n_samples, self.latent_dim)), None

    def _loss_value(self, net: Sequential, plan, after_train: bool = False) -> float:
        return float(plan.acc[0].item()) / net._loss_denominator(plan, after_train)

    def train_on_batch(self, real_samples, labels=None, batch_size: int = 32, n_critic: int = 1) -> tuple[float, float]:
        """models.py:2609-2667."""
        G, D = self.generator, self.discriminator
        real_samples = np.asarray(real_samples)
        rng = np.random.default_rng(self.random_state)
        d_loss_total = 0.0
        # data parallel (parallel.make_gan_data_parallel): every rank is given the same GLOBAL arguments; the globally seeded
        # index choice and latent noise are generated in full and this rank keeps its rows [lo, hi) of each
        gdp = getattr(self, "_dp", None)
        lo, hi = (0, batch_size)
        if gdp is not None:
            from .parallel import shard_rows
            lo, hi = shard_rows(batch_size, gdp.world, gdp.rank)
        global_batch, batch_size = batch_size, hi - lo
        comb_labels = np.zeros((2 * batch_size, 1), dtype=np.float32)
        comb_labels[:batch_size] = self.label_smoothing
        comb_labels[batch_size:] = 1 - self.label_smoothing
        for _ in range(n_critic):
            idx = rng.choice(len(real_samples), global_batch, replace=False)[lo:hi]
            z = self._generate_latent_points(global_batch)[0][lo:hi]
            fake = G._forward_dev(z, training=False)
            d_in_shape = (2 * batch_size,) + tuple(real_samples.shape[1:])
            d_plan = D._plan(d_in_shape, True)
            d_prog = d_plan.program("train")
            d_plan.x_in[:batch_size].copy_(B.to_device(real_samples[idx]).reshape(d_plan.x_in[:batch_size].shape))
            d_plan.x_in[batch_size:].copy_(fake.reshape(d_plan.x_in[batch_size:].shape))
            d_plan.set_target(comb_labels)
            D.y_true = comb_labels
            d_prog.run()
            D._last_plan, D._last_rows, D._pred_host = d_plan, 2 * batch_size, None
            d_loss_total += self._loss_value(D, d_plan, True)
        d_loss_avg = d_loss_total / n_critic

        z = self._generate_latent_points(global_batch)[0][lo:hi]
        fake = G._forward_dev(z, training=True)
        D._forward_dev(fake.reshape((batch_size,) + tuple(real_samples.shape[1:])), training=False)
        ones = np.ones((batch_size, 1), dtype=np.float32)
        D.y_true = ones
        dp = D._last_plan
        dp.set_target(ones)
        dp.program("loss").run()
        if self._dp is not None:   # the generator loss is the mean over the GLOBAL batch: sum the numerators over the ranks
            self._dp.comm.allreduce(dp.acc[0:1])
        g_loss = self._loss_value(D, dp) / (self._dp.world if self._dp is not None else 1)
        # models.py:2661-2664: g_grad = loss.derivative(ones, predictions); with a Sigmoid+BCE discriminator the
        # shortcut in Sequential.backward_pass replaces it by predictions - ones
        g_grad = B.empty(tuple(dp.pred.shape))
        ops.loss_fwd_bwd(self.generator_loss._kind, 0, dp.pred, dp.y_in, g_grad, None, None, None)
        d_grad = D.backward_pass(g_grad, compute_only=True)
        G.backward_pass(d_grad, gan=True)
        self._train_step += 1
        return d_loss_avg, g_loss

    def fit(self, x_train, y_train=None, epochs: int = 100, batch_size: int | None = None, n_critic: int = 5,
            verbose: bool = True, random_state: int | None = None, callbacks: list = [], **unused) -> dict:
        """Core of models.py:2669-2936: per epoch, shuffle, then one train_on_batch per mini-batch."""
        history = History({"discriminator_loss": [], "generator_loss": []})
        x_train = np.asarray(x_train)
        bs = batch_size or len(x_train)
        seed = random_state if random_state is not None else self.random_state
        for epoch in range(epochs):
            start = time.time()
            xs = x_train[shuffle_indices(len(x_train), seed)]
            dl, gl, nb = 0.0, 0.0, 0
            for j in range(0, len(xs) - bs + 1, bs):
                d, g = self.train_on_batch(xs[j:j + bs], None, min(bs, len(xs[j:j + bs])), n_critic)
                dl, gl, nb = dl + d, gl + g, nb + 1
                if verbose:
                    progress_bar(nb, max(len(xs) // bs, 1), message=f"Epoch {epoch + 1}/{epochs} - {time.time() - start:.2f}s - "
                                 f"d_loss: {format_number(dl / nb)} - g_loss: {format_number(gl / nb)}")
            history["discriminator_loss"].append(dl / max(nb, 1))
            history["generator_loss"].append(gl / max(nb, 1))
            if verbose:
                print()
        return history


class Autoencoder(BaseModel):
    """Mirror of neuralnetlib/models.py:778-1565 (the training loop) on the device kernels: encoder / decoder
    layer lists, separate losses and optimizers, and backward_pass's OWN clip rule (models.py:1033-1048: norm clip to
    gradient_clip_threshold, clamp to +-10, divide by np.std + 1e-8) applied to every error tensor and every gradient
    (b200nn_ae_clip). Activations, errors, gradients and optimizer state stay in HBM; the per-layer weight update happens right
    after each layer's backward, as in the reference. Not captured into a graph (this loop is a 'next' row, not the bench path).
    Reference behaviours kept: the decoder is fed the raw encoder output also in variational mode (the reparameterised latent only
    enters the loss VALUE: KL, latent penalties, models.py:1001-1010, :1124-1137); skip connections add encoder Dense outputs to
    decoder Dense outputs of the same width in the forward only (:1014-1019), the backward ignores them."""

    def __init__(self, encoder_layers: list = None, decoder_layers: list = None, gradient_clip_threshold: float = 5.0,
                 enable_padding: bool = False, padding_size: int = 32, random_state: int | None = None,
                 skip_connections: bool = False, l1_reg: float = 0.0, l2_reg: float = 0.0, variational: bool = False):
        super().__init__(gradient_clip_threshold, enable_padding, padding_size, random_state)
        self.encoder_layers = encoder_layers if encoder_layers is not None else []
        self.decoder_layers = decoder_layers if decoder_layers is not None else []
        self.encoder_optimizer = self.decoder_optimizer = None
        self.encoder_loss = self.decoder_loss = None
        self.y_true = self.predictions = self.latent_space = None
        self.skip_connections, self.variational = skip_connections, variational
        self.l1_reg, self.l2_reg = l1_reg, l2_reg
        self.epsilon = 1e-7
        self.latent_dim = None
        self.latent_mean = self.latent_log_var = None
        self.metrics = None
        self._scratch = None

    # ---- construction (models.py:881-974) ---------------------------------------------------------------------------
    @staticmethod
    def _append(lst, layer, first_must_be_input):
        if not lst:
            if first_must_be_input and not isinstance(layer, Input):
                raise ValueError("The first encoder layer must be an Input layer.")
        else:
            prev_t, cur_t = type(lst[-1]), type(layer)
            if cur_t in incompatibility_dict.get(prev_t, []):
                raise ValueError(f"{cur_t.__name__} layer cannot follow {prev_t.__name__} layer.")
        lst.append(layer)
        act = getattr(layer, "activation", getattr(layer, "activation_function", None))
        if act and not isinstance(layer, Activation):
            if isinstance(act, str):
                lst.append(Activation.from_name(act))
            elif isinstance(act, ActivationFunction):
                lst.append(Activation(act))
            elif isinstance(act, Activation):
                lst.append(act)
            else:
                raise ValueError(f"Invalid activation function: {act}")

    def add_encoder_layer(self, layer: Layer):
        self._append(self.encoder_layers, layer, True)

    def add_decoder_layer(self, layer: Layer):
        self._append(self.decoder_layers, layer, False)

    def compile(self, encoder_loss=None, decoder_loss=None, encoder_optimizer=None, decoder_optimizer=None, verbose: bool = False,
                metrics: list | None = None):
        encoder_loss = decoder_loss if encoder_loss is None else encoder_loss
        decoder_loss = encoder_loss if decoder_loss is None else decoder_loss
        encoder_optimizer = decoder_optimizer if encoder_optimizer is None else encoder_optimizer
        decoder_optimizer = encoder_optimizer if decoder_optimizer is None else decoder_optimizer
        if encoder_loss is None or encoder_optimizer is None:
            raise ValueError("At least one loss and optimizer must be specified")
        as_loss = lambda v: v if isinstance(v, LossFunction) else LossFunction.from_name(v)        # noqa: E731
        as_opt = lambda v: v if isinstance(v, Optimizer) else Optimizer.from_name(v)               # noqa: E731
        self.encoder_loss, self.decoder_loss = as_loss(encoder_loss), as_loss(decoder_loss)
        self.encoder_optimizer, self.decoder_optimizer = as_opt(encoder_optimizer), as_opt(decoder_optimizer)
        self.metrics = metrics
        if verbose:
            print(str(self))
        last = self.encoder_layers[-1]
        # models.py:974 reads .units (variational) or .output_shape[-1]; the latter only exists on Embedding-like layers upstream
        self.latent_dim = last.units if self.variational else (last.output_shape[-1] if hasattr(last, "output_shape")
                                                               else getattr(last, "units", None))

    def __str__(self) -> str:
        s = f"Autoencoder(gradient_clip_threshold={self.gradient_clip_threshold}, random_state={self.random_state})\n"
        s += "Encoder:\n" + "".join(f"  Layer {i + 1}: {l}\n" for i, l in enumerate(self.encoder_layers))
        s += "Decoder:\n" + "".join(f"  Layer {i + 1}: {l}\n" for i, l in enumerate(self.decoder_layers))
        return s + f"Losses: {self.encoder_loss} / {self.decoder_loss}\nOptimizers: {self.encoder_optimizer} / {self.decoder_optimizer}\n"

    def summary(self):
        print(str(self))

    # ---- forward (models.py:976-1030) ---------------------------------------------------------------------------------
    def _forward_dev(self, X, training: bool = True) -> torch.Tensor:
        t = X if isinstance(X, torch.Tensor) else B.to_device(np.asarray(X))
        rows = t.shape[0]
        if self.enable_padding and rows % self.padding_size:
            pad = (rows + self.padding_size - 1) // self.padding_size * self.padding_size
            tp = B.zeros((pad,) + tuple(t.shape[1:]))
            tp[:rows].copy_(t)
            t = tp
        skip = {}
        for layer in self.encoder_layers:
            t = layer.forward_pass(t, training) if isinstance(layer, Dropout) else layer.forward_pass(t)   # noqa: F405
            if self.skip_connections and isinstance(layer, Dense):
                skip[layer.units] = t
        self._latent_dev = t
        for layer in self.decoder_layers:
            t = layer.forward_pass(t, training) if isinstance(layer, Dropout) else layer.forward_pass(t)   # noqa: F405
            if self.skip_connections and isinstance(layer, Dense) and skip.get(layer.units) is not None:
                s = skip[layer.units]
                out = B.empty(tuple(t.shape))
                ops.axpy(t.contiguous(), s.contiguous(), 1.0 / np.sqrt(layer.units), out)
                t = out
        return t[:rows]

    def _latent_terms(self, encoded: np.ndarray):
        """(latent_space, kl) of models.py:817-831, 1001-1007: the reparameterisation draws from default_rng(random_state) anew each call."""
        if not self.variational:
            self.latent_mean = self.latent_log_var = None
            return encoded, 0.0
        half = encoded.shape[-1] // 2
        self.latent_mean, self.latent_log_var = encoded[:, :half], encoded[:, half:]
        eps = np.random.default_rng(self.random_state).normal(size=self.latent_mean.shape)
        latent = self.latent_mean + np.exp(0.5 * self.latent_log_var) * eps
        kl = -0.5 * np.mean(1 + self.latent_log_var - np.square(self.latent_mean) - np.exp(self.latent_log_var))
        return latent, float(kl)

    def forward_pass(self, X, training: bool = True):
        out = self._forward_dev(X, training)
        self.latent_space, _ = self._latent_terms(B.to_host(self._latent_dev))
        return out if isinstance(X, torch.Tensor) else B.to_host(out)

    # ---- backward with the Autoencoder's clip rule (models.py:1032-1114) ------------------------------------------------
    def _clip(self, t: torch.Tensor) -> torch.Tensor:
        if t is None:
            return None
        if self._scratch is None:
            self._scratch = B.zeros((3,), dtype=torch.float64)
        out = B.empty(tuple(t.shape))
        ops.ae_clip(t.contiguous(), out, float(self.gradient_clip_threshold), self._scratch)
        return out

    def _update_layer(self, layer, idx, optimizer):
        if not hasattr(layer, "weights") or getattr(layer, "_weights", None) is None:
            return
        dw = self._clip(layer._d_weights)
        if getattr(layer, "_d_bias", None) is not None:
            optimizer.update(idx, layer._weights, dw, layer._bias, self._clip(layer._d_bias))
        else:
            optimizer.update(idx, layer._weights, dw)

    def _backward_dev(self, error: torch.Tensor, pred: torch.Tensor = None, y: torch.Tensor = None):
        for i, layer in enumerate(reversed(self.decoder_layers)):
            if i == 0 and isinstance(layer, Activation):
                name = type(layer.activation_function).__name__
                if pred is not None and name == "Softmax" and isinstance(self.decoder_loss, CategoricalCrossentropy):
                    error = B.to_device(B.to_host(pred) - B.to_host(y))
                elif pred is not None and name == "Sigmoid" and isinstance(self.decoder_loss, BinaryCrossentropy):
                    p, yy = B.to_host(pred), B.to_host(y)
                    error = B.to_device((p - yy) / (p * (1 - p) + 1e-15))
                # any other pair: the loss derivative passes THROUGH the last activation unchanged (models.py:1054-1062)
            else:
                error = layer.backward_pass(self._clip(error))
            self._update_layer(layer, len(self.decoder_layers) - 1 - i, self.decoder_optimizer)
        for i, layer in enumerate(reversed(self.encoder_layers)):
            error = layer.backward_pass(self._clip(error))
            self._update_layer(layer, len(self.encoder_layers) - 1 - i, self.encoder_optimizer)
        return error

    def backward_pass(self, error):
        e = error if isinstance(error, torch.Tensor) else B.to_device(np.asarray(error))
        pred = None if self.predictions is None else B.to_device(np.asarray(self.predictions))
        y = None if self.y_true is None else B.to_device(np.asarray(self.y_true))
        self._backward_dev(e, pred, y)

    def _calculate_regularization(self) -> float:
        reg = 0.0
        if self.l1_reg > 0 or self.l2_reg > 0:
            for layer in list(self.encoder_layers) + list(self.decoder_layers):
                if hasattr(layer, "weights") and layer.weights is not None:
                    w = np.asarray(layer.weights)
                    reg += self.l1_reg * np.sum(np.abs(w)) if self.l1_reg > 0 else 0.0
                    reg += self.l2_reg * np.sum(np.square(w)) if self.l2_reg > 0 else 0.0
        return float(reg)

    def _loss_and_error(self, lf, y_dev, pred_dev, want_err=True):
        p2 = pred_dev.reshape(pred_dev.shape[0], -1)
        y2 = y_dev.reshape(p2.shape)
        err = B.empty(tuple(p2.shape)) if want_err else None
        acc = B.zeros((4,), dtype=torch.float64)
        ops.loss_fwd_bwd(lf._kind, 0, p2.contiguous(), y2.contiguous(), err, acc, None, None, 1, float(getattr(lf, "_param", 1.0)))
        loss = float(acc.cpu()[0].item()) / lf._denominator(tuple(p2.shape))
        return loss, (err.reshape(pred_dev.shape) if want_err else None)

    def train_on_batch(self, x_batch, y_batch=None) -> float:
        """models.py:1116-1146."""
        x = x_batch if isinstance(x_batch, torch.Tensor) else B.to_device(np.asarray(x_batch))
        y = x if y_batch is None else (y_batch if isinstance(y_batch, torch.Tensor) else B.to_device(np.asarray(y_batch)))
        pred = self._forward_dev(x, True)
        self.y_true = B.to_host(y)
        self.predictions = B.to_host(pred)
        self.latent_space, kl = self._latent_terms(B.to_host(self._latent_dev))
        loss, err = self._loss_and_error(self.decoder_loss, y, pred)
        latent_l2 = 0.0001 * float(np.mean(np.square(self.latent_space)))
        penalty = 0.0001 * float(np.mean(np.abs(np.std(self.latent_space, axis=0) - 1.0)))
        if self.skip_connections:
            latent_l2 *= 0.1
            penalty *= 0.1
        total = loss + self._calculate_regularization() + latent_l2 + penalty + 0.01 * kl
        if err.dim() == 1:
            err = err.reshape(-1, 1)
        self._backward_dev(err, pred, y)
        return total

    # ---- inference (models.py:1148-1195) --------------------------------------------------------------------------------
    def predict(self, X, output_latent: bool = False, temperature: float = 1.0) -> np.ndarray:
        out = self._forward_dev(np.asarray(X), False)
        if output_latent:
            return B.to_host(self._latent_dev)
        dec = B.to_host(out)
        if not np.isclose(temperature, 1.0, rtol=1e-09, atol=1e-09):
            dec = np.exp(np.log(np.clip(dec, 1e-7, 1.0)) / temperature)
            dec /= np.sum(dec, axis=-1, keepdims=True)
        return dec

    def evaluate(self, x_test, y_test=None, batch_size: int = 32) -> tuple:
        x_test = np.asarray(x_test)
        y_test = x_test if y_test is None else np.asarray(y_test)
        total, preds = 0.0, []
        n_batches = int(np.ceil(len(x_test) / batch_size))
        for i in range(0, len(x_test), batch_size):
            bx, by = x_test[i:i + batch_size], y_test[i:i + batch_size]
            p = self._forward_dev(bx, False)
            yd = B.to_device(by)
            dl, _ = self._loss_and_error(self.decoder_loss, yd, p, False)
            el, _ = self._loss_and_error(self.encoder_loss, yd, p, False)
            total += (dl + el) / 2
            preds.append(B.to_host(p))
        return total / n_batches, np.vstack(preds)

    def fit(self, x_train, epochs: int, batch_size: int | None = None, verbose: bool = True, metrics: list | None = None,
            random_state: int | None = None, validation_data: tuple | None = None, validation_split: float | None = None,
            callbacks: list = []) -> dict:
        """models.py:1286-1539 (loss / val_loss history, per-epoch reshuffle with the same seed, callbacks)."""
        if self.metrics is not None:
            metrics = self.metrics
        history = History({"loss": [], "val_loss": []})
        if validation_data is not None and validation_split is not None:
            raise ValueError("Cannot specify both validation_data and validation_split")
        x_train = np.asarray(x_train)
        if validation_data is None and validation_split is not None:
            x_train, x_val = train_test_split(x_train, test_size=validation_split, random_state=random_state)
            validation_data = (x_val, x_val)
        if metrics is not None:
            metrics = [Metric(m) for m in metrics]
            for m in metrics:
                history[m.name] = []
                history[f"val_{m.name}"] = []
        callbacks = list(callbacks or [])
        for cb in callbacks:
            cb.on_train_begin({"history": history, "model": self})
        seed = random_state if random_state is not None else self.random_state
        xd = B.to_device(x_train)
        for epoch in range(epochs):
            for cb in callbacks:
                cb.on_epoch_begin(epoch, {"model": self})
            idx = torch.from_numpy(shuffle_indices(x_train.shape[0], seed)).to(B.device())
            t0, error, preds, inputs = time.time(), 0.0, [], []
            step = batch_size if batch_size is not None else x_train.shape[0]
            n_batches = int(np.ceil(x_train.shape[0] / step))
            for bi, j in enumerate(range(0, x_train.shape[0], step)):
                rows = idx[j:j + step]
                xb = B.empty((rows.numel(),) + tuple(xd.shape[1:]))
                ops.gather_rows(xd, rows, xb)
                for cb in callbacks:
                    cb.on_batch_begin(bi, {"batch": bi, "size": rows.numel(), "model": self})
                be = self.train_on_batch(xb)
                error += be
                logs = {"batch": bi, "size": rows.numel(), "model": self, "loss": be}
                if metrics is not None:
                    preds.append(self.predictions); inputs.append(self.y_true)
                    for m in metrics:
                        logs[m.name] = m(preds[-1], inputs[-1])
                for cb in callbacks:
                    cb.on_batch_end(bi, logs)
                if verbose:
                    progress_bar(bi + 1, n_batches, message=f"Epoch {epoch + 1}/{epochs} - loss: {format_number(error / (bi + 1))} - "
                                                            f"{time.time() - t0:.2f}s")
            error /= n_batches
            history["loss"].append(error)
            epoch_logs = {"loss": error, "model": self, "history": history}
            if metrics is not None:
                for m in metrics:
                    history[m.name].append(m(np.vstack(preds), np.vstack(inputs)))
                    epoch_logs[m.name] = history[m.name][-1]
            if validation_data is not None:
                x_val = validation_data[0]
                y_val = validation_data[1] if len(validation_data) > 1 else x_val
                val_loss, val_pred = self.evaluate(x_val, y_val, batch_size or 32)
                history["val_loss"].append(val_loss)
                epoch_logs["val_loss"] = val_loss
                if metrics is not None:
                    for m in metrics:
                        history[f"val_{m.name}"].append(m(val_pred, np.asarray(y_val)))
            if verbose:
                print()
            stop = False
            for cb in callbacks:
                if cb.on_epoch_end(epoch, epoch_logs):
                    stop = True
            if stop:
                break
        for cb in callbacks:
            cb.on_train_end({"history": history, "model": self})
        return history
