"""Mirror of neuralnetlib/activations.py:1-166 (Sigmoid, ReLU, Tanh, Softmax, Linear, LeakyReLU, ELU, SELU, GELU) on the device.

`fn(x)` / `fn.derivative(x)` keep the reference's NumPy-in / NumPy-out contract (float64 results, Q10);
inside a compiled Sequential plan the same kernels are fused into the producing GEMM's epilogue instead.
"""
from __future__ import annotations

import numpy as np

from . import _backend as B
from . import _ops as ops


class ActivationFunction:
    _kind = None      # C-ABI activation code; None = not elementwise-fusable (Softmax)
    _alpha = 0.0

    def __init__(self):
        pass

    def __call__(self, x: np.ndarray) -> np.ndarray:
        raise NotImplementedError

    def derivative(self, x: np.ndarray) -> np.ndarray:
        raise NotImplementedError

    def get_config(self) -> dict:
        return {"name": self.__class__.__name__}

    @staticmethod
    def from_config(config: dict):
        name = config.get("name")
        if not name:
            raise ValueError('Config must contain "name" field')
        params = {k: v for k, v in config.items() if k not in ("name", "config")}
        nested = config.get("config") or {}
        params.update({k: v for k, v in nested.items() if k != "name"})   # e.g. LeakyReLU's alpha
        for cls in _concrete_subclasses(ActivationFunction):
            if cls.__name__ == name:
                return cls(**params)
        raise ValueError(f"Unknown activation function: {name}")

    @staticmethod
    def from_name(name: str) -> "ActivationFunction":
        key = name.lower().replace("_", "")
        for cls in _concrete_subclasses(ActivationFunction):
            if cls.__name__.lower() == key:
                return cls()
        raise ValueError(f"No activation function found for the name: {name}")

    # -- device helpers ---------------------------------------------------------------------------
    def _apply_host(self, x) -> np.ndarray:
        x = np.asarray(x)
        xd = B.to_device(x)
        yd = B.empty(xd.shape)
        ops.act_fwd(self._kind, self._alpha, xd, yd)
        return B.to_host(yd).reshape(x.shape)

    _from_input = False   # GELU: the derivative is a function of the INPUT (the C ABI takes x in place of y)

    def _derivative_host(self, x) -> np.ndarray:
        x = np.asarray(x)
        xd = B.to_device(x)
        yd = B.empty(xd.shape)
        if self._from_input:
            yd = xd
        else:
            ops.act_fwd(self._kind, self._alpha, xd, yd)
        ones = B.to_device(np.ones(x.shape, dtype=np.float32))
        dd = B.empty(xd.shape)
        ops.act_bwd(self._kind, self._alpha, yd, ones, None, (), dd, None)
        return B.to_host(dd).reshape(x.shape)


def _concrete_subclasses(cls):
    out = []
    for sub in cls.__subclasses__():
        if not sub.__name__.startswith("_"):
            out.append(sub)
        out.extend(_concrete_subclasses(sub))
    return out


class _Elementwise(ActivationFunction):
    def __call__(self, x):
        return self._apply_host(x)

    def derivative(self, x):
        return self._derivative_host(x)


class Sigmoid(_Elementwise):      # activations.py:33-43
    _kind = ops.ACT_SIGMOID


class ReLU(_Elementwise):         # activations.py:46-54
    _kind = ops.ACT_RELU


class Tanh(_Elementwise):         # activations.py:57-65
    _kind = ops.ACT_TANH


class Softmax(ActivationFunction):  # activations.py:68-78 (axis 1, no derivative)
    _kind = None

    def __call__(self, x):
        x = np.asarray(x)
        if x.ndim != 2:
            raise ValueError("Softmax expects a 2-D (batch, classes) input")
        xd = B.to_device(x)
        pd = B.empty(xd.shape)
        ops.softmax_fwd(xd, pd)
        return B.to_host(pd)

    def derivative(self, x):
        raise NotImplementedError("Derivative of Softmax is not implemented. It is not needed for backpropagation.")


class Linear(_Elementwise):       # activations.py:81-89
    _kind = ops.ACT_NONE


class LeakyReLU(_Elementwise):    # activations.py:92-109
    _kind = ops.ACT_LEAKYRELU

    def __init__(self, alpha: float = 0.01):
        self.alpha = alpha

    @property
    def _alpha(self):
        return float(self.alpha)

    def get_config(self) -> dict:
        return {"name": self.__class__.__name__, "alpha": self.alpha}

    def __str__(self):
        return f"{self.__class__.__name__}(alpha={self.alpha})"


class ELU(_Elementwise):          # activations.py:112-123
    _kind = ops.ACT_ELU


class SELU(_Elementwise):         # activations.py:126-147
    _kind = ops.ACT_SELU
    DEFAULT_ALPHA = 1.6732632423543772848170429916717
    DEFAULT_SCALE = 1.0507009873554804934193349852946

    def __init__(self, alpha: float = None, scale: float = None):
        self.alpha = alpha if alpha is not None else SELU.DEFAULT_ALPHA
        self.scale = scale if scale is not None else SELU.DEFAULT_SCALE
        if abs(self.scale - SELU.DEFAULT_SCALE) > 1e-12:
            raise NotImplementedError("SELU with a non-default scale is not on the B200 path (the kernels carry one parameter)")

    @property
    def _alpha(self):
        return float(self.alpha)

    def get_config(self) -> dict:
        return {"name": self.__class__.__name__, "alpha": self.alpha, "scale": self.scale}

    def __str__(self):
        return f"{self.__class__.__name__}(alpha={self.alpha}, scale={self.scale})"


class GELU(_Elementwise):         # activations.py:150-166 (tanh approximation); its derivative needs the pre-activation
    _kind = ops.ACT_GELU
    _from_input = True
