"""Mirror of neuralnetlib/optimizers.py (Optimizer, SGD :40-60, Momentum :63-111, RMSprop :114-164, Adam :167-283,
AdaBelief :286-402, RAdam :405-523) with device-resident state.

`update(layer_index, weights, weights_grad, bias, bias_grad)` keeps the reference contract: it mutates
`weights` / `bias` IN PLACE (NumPy arrays are uploaded, updated by the fused multi-tensor kernel and written
back; device tensors are updated where they live). Inside a compiled Sequential plan the same kernel updates
every parameter tensor of the model in one launch, with Adam's per-layer `t` (optimizers.py:236) and the
per-tensor gradient clip (models.py:240-242) computed on the device.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _backend as B
from . import _ops as ops


class Optimizer:
    def __init__(self, learning_rate: float = 0.01):
        self._hyper = None
        self.learning_rate = learning_rate

    def update(self, layer_index: int, weights, weights_grad, bias=None, bias_grad=None):
        raise NotImplementedError

    def get_config(self) -> dict:
        return {"name": self.__class__.__name__}

    @staticmethod
    def from_config(config: dict):
        for cls in _public_optimizers():
            if cls.__name__ == config["name"]:
                return cls.from_config(config) if cls.from_config is not Optimizer.from_config else cls(
                    **{k: v for k, v in config.items() if k != "name"})
        raise ValueError(f"No optimizer found for the name: {config['name']}")

    @staticmethod
    def from_name(name: str) -> "Optimizer":
        key = name.lower().replace("_", "")
        for cls in _public_optimizers():
            if cls.__name__.lower() == key:
                return cls()
        raise ValueError(f"No optimizer found for the name: {name}")

    # -- learning_rate is read by the kernels from device memory so that callbacks.LearningRateScheduler
    #    (callbacks.py:346-354) can change it without re-capturing the step graph
    @property
    def learning_rate(self):
        return self._lr

    @learning_rate.setter
    def learning_rate(self, v):
        self._lr = v
        if getattr(self, "_hyper", None) is not None:
            self._upload_hyper()

    def _hyper_values(self):
        return [float(self._lr), 0.0, 0.0, 0.0, 0.0, 0.0]

    def _upload_hyper(self):
        vals = torch.tensor(self._hyper_values(), dtype=torch.float64)
        if self._hyper is None:
            self._hyper = vals.to(B.device())
        else:
            self._hyper.copy_(vals)

    def _ensure_device(self):
        if self._hyper is None:
            B.stream_obj()
            self._upload_hyper()

    # plan interface -------------------------------------------------------------------------------
    def _t_counter(self):
        return None

    def _make_plan(self, updates):
        """updates: list of (layer_index, w, dw, b, db) device tensors in update order (last layer first)."""
        raise NotImplementedError

    def _run_plan(self, plan, stream=None):
        raise NotImplementedError

    @staticmethod
    def _plan_io(plan):
        """Tensors one optimizer launch reads/writes once each: p, g (+ m, v); the clip norm re-reads g."""
        io = []
        for entry in plan.keep:
            io.extend(entry[:4])
        return tuple(io)

    @staticmethod
    def _as_pairs(weights, weights_grad, bias, bias_grad):
        pairs = [(weights, weights_grad)]
        if bias is not None and bias_grad is not None:
            pairs.append((bias, bias_grad))
        return pairs


def _public_optimizers():
    out, todo = [], list(Optimizer.__subclasses__())
    while todo:
        c = todo.pop()
        todo.extend(c.__subclasses__())
        if not c.__name__.startswith("_"):
            out.append(c)
    return out


def _stage(param, grad):
    """-> (device param, device grad, writeback or None)."""
    if isinstance(param, torch.Tensor):
        g = grad if isinstance(grad, torch.Tensor) else B.to_device(grad)
        return param, g.reshape(-1) if g.dim() == 0 else g, None
    if not isinstance(param, np.ndarray):
        raise TypeError("parameters must be numpy arrays or device tensors")
    pd = B.to_device(param)
    gd = B.to_device(np.broadcast_to(np.asarray(grad, dtype=np.float64), param.shape))
    return pd, gd, param


class SGD(Optimizer):
    def __init__(self, learning_rate: float = 0.01, **kwargs):
        super().__init__(learning_rate)
        for k, v in kwargs.items():
            setattr(self, k, v)

    def update(self, layer_index, weights, weights_grad, bias=None, bias_grad=None):
        self._ensure_device()
        staged = [_stage(p, g) for p, g in self._as_pairs(weights, weights_grad, bias, bias_grad)]
        plan = ops.OptPlan([(p, g.contiguous(), None, None, 1, 0) for p, g, _ in staged], 1)
        ops.sgd_multi(plan, self._hyper)
        for p, _, host in staged:
            if host is not None:
                host[...] = B.to_host(p).reshape(host.shape)

    def _make_plan(self, updates, ss=None, j0=1, n_layers=None):
        self._ensure_device()
        entries = []
        for j, upd in enumerate(updates, start=j0):
            idx, w, dw, b, db = upd[:5]
            chain = upd[5] if len(upd) > 5 else ()
            entries.append((w, dw, None, None, j, 1, chain))
            if b is not None:
                entries.append((b, db, None, None, j, 1, chain))
        return ops.OptPlan(entries, n_layers or max(len(updates), 1), ss)

    def _run_plan(self, plan, stream=None):
        ops.sgd_multi(plan, self._hyper, stream)

    def get_config(self) -> dict:
        return {"name": self.__class__.__name__, "learning_rate": self.learning_rate}

    def __str__(self):
        return f"{self.__class__.__name__}(learning_rate={self.learning_rate})"

    @staticmethod
    def from_config(config: dict):
        return SGD(config["learning_rate"])


class _StateOptimizer(Optimizer):
    """Update rules with per-layer state tensors (and, for the Adam family, the per-layer step count `t`): state lives on the device
    as (first, second) tensor pairs per (kind, layer); subclasses name the reference attributes they appear under."""
    _rule = ops.OPT_ADAM        # b200nn_opt_multi kind
    _uses_t = True
    _fast_adam = False          # plain Adam: the tuned b200nn_adam_multi kernels
    _slots = (True, True)       # which of (first, second) state tensors the rule keeps

    def _init_state(self):
        self._t_host = 0
        self._t_dev = None
        self._state = {}    # (kind, layer_index) -> (first, second) device tensors, kind in {"w", "b"}
        self._shapes = {}   # (kind, layer_index) -> reference shape of the parameter
        self._min_denom = 1e-16

    def _hyper_values(self):
        """{lr, beta1 | momentum | rho, beta2, eps, clip_norm, clip_value} (include/b200nn.h)."""
        raise NotImplementedError

    def _ensure_device(self):
        super()._ensure_device()
        if self._t_dev is None:
            self._t_dev = torch.tensor([self._t_host], dtype=torch.int64).to(B.device())

    # `t` counts update() calls, i.e. LAYERS, not steps (optimizers.py:236). It lives on the device.
    @property
    def t(self):
        if self._t_dev is None:
            return self._t_host
        return int(self._t_dev.cpu()[0].item())

    @t.setter
    def t(self, v):
        self._t_host = int(v)
        if self._t_dev is not None:
            self._t_dev.fill_(int(v))

    def _moments(self, kind, layer_index, like: torch.Tensor, shape=None):
        key = (kind, layer_index)
        if key not in self._state:
            self._state[key] = tuple(B.zeros(like.shape) if keep else None for keep in self._slots)
            self._shapes[key] = tuple(shape if shape is not None else like.shape)
        return self._state[key]

    def update(self, layer_index, weights, weights_grad, bias=None, bias_grad=None) -> None:
        self._ensure_device()
        staged = []
        for kind, (p, g) in zip(("w", "b"), self._as_pairs(weights, weights_grad, bias, bias_grad)):
            pd, gd, host = _stage(p, g)
            m, v = self._moments(kind, layer_index, pd, None if host is None else host.shape)
            staged.append((pd, gd.contiguous(), m, v, host))
        plan = ops.OptPlan([(pd, gd, m, v, 1, 0) for pd, gd, m, v, _ in staged], 1)
        if self._uses_t:
            ops.step_begin(None, None, self._t_dev, 1)   # self.t += 1
        self._run_plan(plan)
        for pd, _, _, _, host in staged:
            if host is not None:
                host[...] = B.to_host(pd).reshape(host.shape)

    def _make_plan(self, updates, ss=None, j0=1, n_layers=None):
        """updates: (layer_index, w, dw, b, db[, pending chain]) in update order (last layer first). j0 / n_layers let one
        step's updates be split over several launches (the per-layer `t` is counter - n_layers + j, optimizers.py:236)."""
        self._ensure_device()
        entries = []
        for j, upd in enumerate(updates, start=j0):
            idx, w, dw, b, db = upd[:5]
            chain = upd[5] if len(upd) > 5 else ()
            m, v = self._moments("w", idx, w)
            entries.append((w, dw, m, v, j, 1, chain))
            if b is not None:
                m, v = self._moments("b", idx, b)
                entries.append((b, db, m, v, j, 1, chain))
        return ops.OptPlan(entries, n_layers or max(len(updates), 1), ss)

    def _t_counter(self):
        self._ensure_device()
        return self._t_dev if self._uses_t else None

    def _run_plan(self, plan, stream=None):
        if self._fast_adam and self.clip_norm is None and self.clip_value is None:
            ops.adam_multi(plan, self._hyper, self._t_dev, stream)
        else:
            ops.opt_multi(self._rule, plan, self._hyper, self._t_dev if self._uses_t else None, stream)

    # reference-shaped views of the state (optimizers.py:247-261) --------------------------------------
    def _host_state(self, kind, which):
        return {idx: B.to_host(mv[which]).reshape(self._shapes[(k, idx)])
                for (k, idx), mv in self._state.items() if k == kind}

    def _set_state(self, kind, which, d):
        self._ensure_device()
        for idx, arr in d.items():
            arr = np.asarray(arr, dtype=np.float64)
            key = (kind, int(idx))
            if key not in self._state:
                self._state[key] = tuple(B.zeros(arr.shape) if keep else None for keep in self._slots)
                self._shapes[key] = arr.shape
            self._state[key][which].copy_(B.to_device(arr).reshape(self._state[key][which].shape))


def _state_property(kind, which):
    return property(lambda self: self._host_state(kind, which), lambda self, d: self._set_state(kind, which, d))


def _tolist(d):
    return {k: v.tolist() for k, v in d.items()}


def _restore(opt, config, names):
    for name in names:
        if config.get(name):
            setattr(opt, name, {int(k): np.array(v) for k, v in config[name].items()})


class Momentum(_StateOptimizer):
    """optimizers.py:63-111: velocity = momentum * velocity - lr * grad; param += velocity."""
    _rule, _uses_t, _slots = ops.OPT_MOMENTUM, False, (True, False)
    velocity_w, velocity_b = _state_property("w", 0), _state_property("b", 0)

    def __init__(self, learning_rate: float = 0.01, momentum: float = 0.9, **kwargs):
        self.momentum = momentum
        self.clip_norm = self.clip_value = None
        self._init_state()
        super().__init__(learning_rate)
        for k, v in kwargs.items():
            setattr(self, k, v)

    def _hyper_values(self):
        return [float(self._lr), float(self.momentum), 0.0, 0.0, 0.0, 0.0]

    def get_config(self) -> dict:
        return {"name": self.__class__.__name__, "learning_rate": self.learning_rate, "momentum": self.momentum,
                "velocity_w": _tolist(self.velocity_w), "velocity_b": _tolist(self.velocity_b)}

    @staticmethod
    def from_config(config: dict):
        opt = Momentum(config["learning_rate"], config["momentum"])
        _restore(opt, config, ("velocity_w", "velocity_b"))
        return opt

    def __str__(self):
        return f"{self.__class__.__name__}(learning_rate={self.learning_rate}, momentum={self.momentum})"


class RMSprop(_StateOptimizer):
    """optimizers.py:114-164: sq = rho * sq + (1 - rho) * grad^2; param -= lr * grad / (sqrt(sq) + eps)."""
    _rule, _uses_t, _slots = ops.OPT_RMSPROP, False, (False, True)
    sq_grads_w, sq_grads_b = _state_property("w", 1), _state_property("b", 1)

    def __init__(self, learning_rate: float = 0.01, rho: float = 0.9, epsilon: float = 1e-8, **kwargs):
        self.rho, self.epsilon = rho, epsilon
        self.clip_norm = self.clip_value = None
        self._init_state()
        super().__init__(learning_rate)
        for k, v in kwargs.items():
            setattr(self, k, v)

    def _hyper_values(self):
        return [float(self._lr), float(self.rho), 0.0, float(self.epsilon), 0.0, 0.0]

    def get_config(self) -> dict:
        return {"name": self.__class__.__name__, "learning_rate": self.learning_rate, "rho": self.rho, "epsilon": self.epsilon,
                "sq_grads_w": _tolist(self.sq_grads_w), "sq_grads_b": _tolist(self.sq_grads_b)}

    @staticmethod
    def from_config(config: dict):
        opt = RMSprop(config["learning_rate"], config["rho"], config["epsilon"])
        _restore(opt, config, ("sq_grads_w", "sq_grads_b"))
        return opt

    def __str__(self):
        return f"{self.__class__.__name__}(learning_rate={self.learning_rate}, rho={self.rho}, epsilon={self.epsilon})"


class _AdamFamily(_StateOptimizer):
    _second = "v"

    def __init__(self, learning_rate, beta_1, beta_2, epsilon, clip_norm, clip_value, **kwargs):
        self.beta_1, self.beta_2, self.epsilon = beta_1, beta_2, epsilon
        self.clip_norm, self.clip_value = clip_norm, clip_value
        self._init_state()
        Optimizer.__init__(self, learning_rate)
        for k, v in kwargs.items():
            setattr(self, k, v)

    def _hyper_values(self):
        return [float(self._lr), float(self.beta_1), float(self.beta_2), float(self.epsilon),
                float(self.clip_norm) if self.clip_norm is not None else 0.0,
                float(self.clip_value) if self.clip_value is not None else 0.0]

    def get_config(self) -> dict:
        sec = self._second
        return {"name": self.__class__.__name__, "learning_rate": self.learning_rate, "beta_1": self.beta_1,
                "beta_2": self.beta_2, "epsilon": self.epsilon, "clip_norm": self.clip_norm, "clip_value": self.clip_value,
                "t": self.t, "m_w": _tolist(self.m_w), f"{sec}_w": _tolist(getattr(self, f"{sec}_w")), "m_b": _tolist(self.m_b),
                f"{sec}_b": _tolist(getattr(self, f"{sec}_b"))}

    @classmethod
    def from_config(cls, config: dict):
        opt = cls(learning_rate=config["learning_rate"], beta_1=config["beta_1"], beta_2=config["beta_2"],
                  epsilon=config["epsilon"], clip_norm=config.get("clip_norm"), clip_value=config.get("clip_value"))
        opt.t = config["t"]
        _restore(opt, config, ("m_w", f"{cls._second}_w", "m_b", f"{cls._second}_b"))
        return opt

    def __str__(self):
        return (f"{self.__class__.__name__}(learning_rate={self.learning_rate}, beta_1={self.beta_1}, "
                f"beta_2={self.beta_2}, epsilon={self.epsilon}, clip_norm={self.clip_norm}, clip_value={self.clip_value})")


class Adam(_AdamFamily):
    """optimizers.py:167-283. clip_norm / clip_value (:190-202) are applied on the device after Sequential's own gradient clip."""
    _rule, _fast_adam = ops.OPT_ADAM, True
    m_w, v_w, m_b, v_b = _state_property("w", 0), _state_property("w", 1), _state_property("b", 0), _state_property("b", 1)

    def __init__(self, learning_rate: float = 0.001, beta_1: float = 0.9, beta_2: float = 0.999, epsilon: float = 1e-8,
                 clip_norm: float = None, clip_value: float = None, **kwargs) -> None:
        super().__init__(learning_rate, beta_1, beta_2, epsilon, clip_norm, clip_value, **kwargs)


class AdaBelief(_AdamFamily):
    """optimizers.py:286-402: the second moment tracks (grad - m)^2; epsilon inside the square root."""
    _rule, _second = ops.OPT_ADABELIEF, "s"
    m_w, s_w, m_b, s_b = _state_property("w", 0), _state_property("w", 1), _state_property("b", 0), _state_property("b", 1)

    def __init__(self, learning_rate: float = 0.001, beta_1: float = 0.9, beta_2: float = 0.999, epsilon: float = 1e-16,
                 clip_norm: float = None, clip_value: float = None, **kwargs) -> None:
        super().__init__(learning_rate, beta_1, beta_2, epsilon, clip_norm, clip_value, **kwargs)


class RAdam(_AdamFamily):
    """optimizers.py:405-523: variance rectification r_t once rho_t > 4, plain momentum before."""
    _rule = ops.OPT_RADAM
    m_w, v_w, m_b, v_b = _state_property("w", 0), _state_property("w", 1), _state_property("b", 0), _state_property("b", 1)

    def __init__(self, learning_rate: float = 0.001, beta_1: float = 0.9, beta_2: float = 0.999, epsilon: float = 1e-8,
                 clip_norm: float = None, clip_value: float = None, **kwargs) -> None:
        super().__init__(learning_rate, beta_1, beta_2, epsilon, clip_norm, clip_value, **kwargs)
        self.rho_inf = 2 / (1 - beta_2) - 1
