"""Mirror of the hot-path layers of neuralnetlib/layers.py on the device:
Layer :14-36, Input :39-64, Dense :67-218, Activation :221-261, Conv2D :345-530, MaxPooling2D :533-643,
Flatten :646-664, BatchNormalization :1205-1304, Reshape :1563-1588, Conv2DTranspose :2536-2692.

Same constructor signatures, attributes (`weights`, `bias`, `d_weights`, `d_bias`, `gamma`, ... read as float64
NumPy arrays, assignable), `get_config`/`from_config` and error behaviour. Parameters, gradients and caches live
in HBM as float32; weights are initialised on the host with the reference's own NumPy generator calls so that
seeds reproduce. `forward_pass` / `backward_pass` accept NumPy arrays (returning float64 arrays,
Q10) or device tensors. Inside `Sequential`, layers are compiled into a fused, graph-captured plan
(`engine.py`) through the `_plan_forward` / `_plan_backward` hooks.
"""
from __future__ import annotations

import time

import os

import numpy as np
import torch

from . import _backend as B
from . import _ops as ops
from . import get_math
from .activations import ActivationFunction, Softmax


def _math_code():
    return {"tf32": ops.MATH_TF32, "tf32x3": ops.MATH_TF32X3}.get(get_math(), ops.MATH_FP32)


def _is_dev(x):
    return isinstance(x, torch.Tensor)


def _in(x):
    """-> (device float32 tensor, was_host)."""
    if _is_dev(x):
        return x.contiguous(), False
    return B.to_device(np.asarray(x)), True


def _out(t, was_host):
    return B.to_host(t) if was_host else t


class _Err:
    """An error tensor together with the clip slots still pending on it (see include/b200nn.h)."""
    __slots__ = ("t", "chain")

    def __init__(self, t, chain=()):
        self.t, self.chain = t, tuple(chain)


def _pending_factor(pend):
    """Multiplier of a clip chain that is still pending on a gradient attribute: pend = (ss tensor, first slot, count);
    the same recurrence as chain_scale in csrc/common.cuh, evaluated on the host when the attribute is read."""
    if pend is None or pend[2] <= 0:
        return 1.0
    ss, lo, cnt = pend
    vals = ss[lo:lo + cnt].cpu().numpy()
    acc = 1.0
    for v in vals:
        n2 = acc * acc * float(v)
        if n2 > 1.0:
            acc /= np.sqrt(n2)
    return acc


class _EagerCtx:
    """Allocation context of the layer-by-layer (non-plan) API: fresh buffers per call, no clip slots."""
    ss = None
    dp = None
    world = 1
    defer_clips = False

    def pending_chain(self):
        return None

    def buf(self, shape):
        return B.empty(tuple(int(s) for s in shape))

    def slot(self):
        return None

    def need_ws(self, nbytes):
        self._ws = B.empty((max(int(nbytes) // 4, 1),))

    @property
    def ws(self):
        return getattr(self, "_ws", None)

    def emit(self, fn, name=None, io=()):
        fn()

    def sync_slots_now(self):
        pass

    def peer_region(self, nbytes):
        return None


def _act_of(layer):
    """(kind, alpha) of a fusable Activation layer, else (ACT_NONE, 0)."""
    if layer is None:
        return ops.ACT_NONE, 0.0
    f = layer.activation_function
    return f._kind, float(f._alpha)


class Layer:
    def __init__(self):
        self.input = None
        self.output = None

    def forward_pass(self, input_data):
        raise NotImplementedError

    def backward_pass(self, output_error):
        raise NotImplementedError

    def get_config(self) -> dict:
        raise NotImplementedError

    @staticmethod
    def from_config(config: dict) -> "Layer":
        name = config["name"]
        cls = globals().get(name)
        if cls is None or not isinstance(cls, type) or not issubclass(cls, Layer):
            raise ValueError(f"Invalid layer name: {name}. Make sure the class {name} is defined.")
        return cls.from_config(config)

    # ---- plan protocol ----------------------------------------------------------------------------------
    _is_view = False         # Input / Flatten / Reshape: no kernel, no clip slot
    _fuses_act = False       # forward can apply a following Activation in its epilogue
    _supports_mask = False   # backward can fold the preceding Activation's backward into its dx epilogue

    def _out_shape(self, in_shape):
        return in_shape


class _ParamMixin:
    """Device-resident parameters with the reference's NumPy attribute surface."""

    def _param_get(self, name):
        t = getattr(self, "_" + name, None)
        if t is None:
            return None
        a = B.to_host(t).reshape(self._shapes[name])
        if name in ("d_weights", "d_bias") and getattr(self, "_grad_pending", None) is not None:
            a = a * _pending_factor(self._grad_pending)   # deferred clips: the device holds the gradient of the UNCLIPPED error
        return a

    def _param_set(self, name, value):
        if value is None:
            setattr(self, "_" + name, None)
            return
        arr = np.asarray(value, dtype=np.float64)
        cur = getattr(self, "_" + name, None)
        if not hasattr(self, "_shapes"):
            self._shapes = {}
        if cur is not None and cur.numel() == arr.size:
            cur.copy_(B.to_device(arr).reshape(cur.shape))   # keep the pointer: captured graphs stay valid
        else:
            setattr(self, "_" + name, B.to_device(arr))
            if cur is not None:      # the pointer changed under plans / graphs / optimizer tables that captured the old one
                from . import invalidate_plans
                invalidate_plans()
        self._shapes[name] = arr.shape


def _param_property(name):
    return property(lambda self: self._param_get(name), lambda self, v: self._param_set(name, v))


# =======================================================================================================
class Input(Layer):
    _is_view = True

    def __init__(self, input_shape):
        if isinstance(input_shape, int):
            input_shape = (input_shape,)
        self.input_shape = input_shape
        self.input_dim = np.prod(input_shape)

    def __str__(self) -> str:
        return f"Input(input_shape={self.input_shape})"

    def forward_pass(self, input_data):
        return input_data

    def backward_pass(self, output_error):
        return output_error

    def get_config(self) -> dict:
        return {"name": self.__class__.__name__, "input_shape": self.input_shape, "input_dim": self.input_dim}

    @staticmethod
    def from_config(config: dict):
        return Input(config["input_shape"])


# =======================================================================================================
class Dense(Layer, _ParamMixin):
    _fuses_act = True
    _supports_mask = True
    weights = _param_property("weights")
    bias = _param_property("bias")
    d_weights = _param_property("d_weights")
    d_bias = _param_property("d_bias")

    def __init__(self, units: int, weights_init: str = "glorot_uniform", bias_init: str = "zeros",
                 random_state: int = None, init_scale: float = 1.0, input_dim: int = None, **kwargs):
        super().__init__()
        self.units = units
        self._weights = self._bias = self._d_weights = self._d_bias = None
        self._shapes = {}
        self.weights_init = weights_init
        self.bias_init = bias_init
        self.random_state = random_state
        self.init_scale = init_scale
        self.input_dim = input_dim
        for key, value in kwargs.items():
            setattr(self, key, value)

    def __str__(self) -> str:
        return f"Dense(units={self.units})"

    def initialize_weights(self, input_size: int):
        """Host-side, same generator calls as layers.py:89-150 so that seeds reproduce."""
        self.rng = np.random.default_rng(self.random_state if self.random_state is not None else int(time.time_ns()))
        fan_in, fan_out = input_size, self.units
        shape = (input_size, self.units)
        uniform = {"glorot_uniform": 6 / (fan_in + fan_out), "xavier_uniform": 6 / (fan_in + fan_out),
                   "he_uniform": 6 / fan_in, "lecun_uniform": 3 / fan_in}
        normal = {"glorot_normal": 2 / (fan_in + fan_out), "xavier_normal": 2 / (fan_in + fan_out),
                  "he_normal": 2 / fan_in, "lecun_normal": 1 / fan_in}
        if self.weights_init == "scaled_normal":
            w = self.rng.normal(0, self.init_scale / np.sqrt(fan_in), shape)
        elif self.weights_init in uniform:
            lim = np.sqrt(uniform[self.weights_init])
            w = self.rng.uniform(-lim, lim, shape)
        elif self.weights_init in normal:
            w = self.rng.normal(0, np.sqrt(normal[self.weights_init]), shape)
        elif self.weights_init == "orthogonal":
            q, r = np.linalg.qr(self.rng.normal(0, 1, shape))
            w = q * np.sign(np.diag(r))
        else:
            raise ValueError(
                "Invalid weights_init value. Possible values are 'scaled_normal', 'glorot_uniform', 'glorot_normal', "
                "'he_uniform', 'he_normal', 'lecun_uniform', 'lecun_normal', 'orthogonal'")
        if self.bias_init == "zeros":
            b = np.zeros((1, self.units))
        elif self.bias_init == "ones":
            b = np.ones((1, self.units))
        elif self.bias_init == "normal":
            b = self.rng.normal(0, 0.05, (1, self.units))
        elif self.bias_init == "uniform":
            b = self.rng.uniform(-0.05, 0.05, (1, self.units))
        else:
            raise ValueError("Invalid bias_init value. Possible values are 'zeros', 'ones', 'normal', 'uniform'")
        self.weights, self.bias = w, b
        self.d_weights, self.d_bias = np.zeros_like(w), np.zeros_like(b)

    def _ensure_grads(self):
        if self._d_weights is None:
            self.d_weights = np.zeros(self._shapes["weights"])
            self.d_bias = np.zeros(self._shapes["bias"])

    # -- plan hooks
    def _out_shape(self, in_shape):
        return in_shape[:-1] + (self.units,)

    def _plan_forward(self, ctx, st, x, act_layer, training):
        feat = x.shape[-1]
        if self._weights is None:
            self.initialize_weights(feat)
        self._ensure_grads()
        if self._weights.shape[0] != feat:
            raise ValueError(f"shapes {tuple(x.shape)} and {tuple(self._weights.shape)} not aligned")
        M = x.numel() // feat
        ctx.need_ws(ops.dense_ws_bytes(M, self.units, feat, _math_code()))
        y = ctx.buf(tuple(x.shape[:-1]) + (self.units,))
        kind, alpha = _act_of(act_layer)
        math = _math_code()
        st.update(x=x, math=math)
        w, b = self._weights, self._bias
        ctx.emit(lambda: ops.dense_fwd(math, x, w, b, y, kind, alpha, ctx.ws), "dense_fwd", (x, w, b, y))
        return y

    def _plan_backward(self, ctx, st, err, mask_layer, need_dx):
        x = st["x"]
        dx = ctx.buf(x.shape) if need_dx else None
        kind, alpha = _act_of(mask_layer)
        s0 = ctx.slot() if need_dx else None
        s1 = ctx.slot() if (need_dx and mask_layer is not None) else None
        w, dw, db, ss, chain, math = self._weights, self._d_weights, self._d_bias, ctx.ss, err.chain, st["math"]
        et = err.t
        dwp = st.get("dw_partial")   # set by the engine when the fused head kernel of this train step already formed x^T . err partials
        if dwp is not None:
            ctx.emit(lambda: ops.dense_bwd_head(x, w, et, ss, chain, dx, dw, db, kind, alpha, s0, s1, dwp), "dense_bwd_head",
                     (et, dw, db) + ((w, dx) if need_dx else ()))
        else:
            ctx.emit(lambda: ops.dense_bwd(math, x, w, et, ss, chain, dx, dw, db, kind, alpha, s0, s1, ctx.ws), "dense_bwd",
                     (x, et, dw, db) + ((w, dx) if need_dx else ()))
        if not need_dx:
            return None
        return _Err(dx, tuple(s for s in (s0, s1) if s is not None))

    # -- reference API (layers.py:152-198)
    def forward_pass(self, input_data):
        x, host = _in(input_data)
        self.input_shape = tuple(x.shape)
        if x.dim() == 1:
            x = x.reshape(1, -1)
        self.input = input_data
        ctx = _EagerCtx()
        self._st = {}
        return _out(self._plan_forward(ctx, self._st, x, None, True), host)

    def backward_pass(self, output_error):
        e, host = _in(output_error)
        dx = self._plan_backward(_EagerCtx(), self._st, _Err(e.reshape(-1, self.units)), None, True)
        return _out(dx.t.reshape(self._st["x"].shape), host)

    def get_config(self) -> dict:
        w, b = self.weights, self.bias
        return {"name": self.__class__.__name__, "weights": w.tolist() if w is not None else None,
                "bias": b.tolist() if b is not None else None, "units": self.units, "weights_init": self.weights_init,
                "bias_init": self.bias_init, "random_state": self.random_state}

    @staticmethod
    def from_config(config: dict):
        layer = Dense(config["units"], config["weights_init"], config["bias_init"], config["random_state"])
        if config["weights"] is not None:
            layer.weights = np.array(config["weights"])
            layer.bias = np.array(config["bias"])
        return layer


# =======================================================================================================
class Activation(Layer):
    def __init__(self, activation_function):
        super().__init__()
        self.activation_function = activation_function if isinstance(activation_function, ActivationFunction) \
            else ActivationFunction.from_name(activation_function)

    def __str__(self) -> str:
        return f"Activation({type(self.activation_function).__name__})"

    @property
    def _fusable(self):
        f = self.activation_function
        return (f._kind is not None and not (f._kind == ops.ACT_LEAKYRELU and f._alpha <= 0)
                and not getattr(f, "_from_input", False))   # GELU's derivative needs the pre-activation: never fused

    def _plan_forward(self, ctx, st, x, act_layer, training):
        y = ctx.buf(x.shape)
        f = self.activation_function
        if isinstance(f, Softmax):
            if x.dim() != 2:
                raise ValueError("Softmax expects a 2-D (batch, classes) input")
            ctx.emit(lambda: ops.softmax_fwd(x, y), "softmax_fwd", (x, y))
        else:
            kind, alpha = f._kind, float(f._alpha)
            ctx.emit(lambda: ops.act_fwd(kind, alpha, x, y), "act_fwd", (x, y))
        st.update(y=y, x=x)
        return y

    def _plan_backward(self, ctx, st, err, mask_layer, need_dx, y=None):
        f = self.activation_function
        if f._kind is None:
            raise NotImplementedError("Derivative of Softmax is not implemented. It is not needed for backpropagation.")
        y = st["y"] if y is None else y
        if getattr(f, "_from_input", False):
            y = st["x"]          # b200nn_act_bwd takes the INPUT for GELU (activations.py:157-161)
        dx = ctx.buf(y.shape)
        s = ctx.slot()
        kind, alpha, ss, chain, et = f._kind, float(f._alpha), ctx.ss, err.chain, err.t
        ctx.emit(lambda: ops.act_bwd(kind, alpha, y, et, ss, chain, dx, s), "act_bwd", (y, et, dx))
        return _Err(dx, () if s is None else (s,))

    def forward_pass(self, input_data):
        x, host = _in(input_data)
        self.input = input_data
        self._st = {}
        y = self._plan_forward(_EagerCtx(), self._st, x, None, True)
        self.output = _out(y, host)
        return self.output

    def backward_pass(self, output_error):
        e, host = _in(output_error)
        return _out(self._plan_backward(_EagerCtx(), self._st, _Err(e.reshape(self._st["y"].shape)), None, True).t, host)

    def get_config(self) -> dict:
        f = self.activation_function
        return {"name": self.__class__.__name__,
                "activation_function": {"name": f.__class__.__name__, "config": f.get_config() if hasattr(f, "get_config") else {}}}

    @staticmethod
    def from_config(config: dict):
        return Activation(ActivationFunction.from_config(config["activation_function"]))

    @staticmethod
    def from_name(name: str) -> "Activation":
        key = name.lower().replace("_", "")
        for sub in _all_subclasses(ActivationFunction):
            if sub.__name__.lower() == key:
                return Activation(sub())
        raise ValueError(f"No activation function found for the name: {name}")


def _all_subclasses(cls):
    out = []
    for s in cls.__subclasses__():
        if not s.__name__.startswith("_"):
            out.append(s)
        out.extend(_all_subclasses(s))
    return out


# =======================================================================================================
def _pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


class Conv2D(Layer, _ParamMixin):
    _fuses_act = True
    _supports_mask = True
    weights = _param_property("weights")
    bias = _param_property("bias")
    d_weights = _param_property("d_weights")
    d_bias = _param_property("d_bias")

    def __init__(self, filters: int, kernel_size, strides=1, padding: str = "valid", weights_init: str = "default",
                 bias_init: str = "default", random_state: int = None, **kwargs):
        self.filters = filters
        self.kernel_size = _pair(kernel_size)
        self.strides = _pair(strides)
        self.padding = padding
        self._weights = self._bias = self._d_weights = self._d_bias = None
        self._shapes = {}
        self.weights_init = weights_init
        self.bias_init = bias_init
        self.random_state = random_state
        for key, value in kwargs.items():
            setattr(self, key, value)

    def initialize_weights(self, input_shape: tuple):
        """layers.py:367-398 (same generator calls)."""
        in_channels = input_shape[3]
        fan_in = np.prod(self.kernel_size) * in_channels
        fan_out = np.prod(self.kernel_size) * self.filters
        self.rng = np.random.default_rng(self.random_state if self.random_state is not None else int(time.time_ns()))
        shape = (*self.kernel_size, in_channels, self.filters)
        if self.weights_init in ("glorot_uniform", "xavier"):
            lim = np.sqrt(6 / (fan_in + fan_out))
            w = self.rng.uniform(-lim, lim, shape)
        elif self.weights_init == "he":
            w = self.rng.normal(0, np.sqrt(2 / fan_in), shape)
        elif self.weights_init == "default":
            w = self.rng.normal(0, 0.01, shape)
        else:
            raise ValueError("Invalid weights_init value. Possible values are 'xavier', 'he', and 'default'.")
        if self.bias_init == "default":
            b = np.zeros((1, self.filters))
        elif self.bias_init == "normal":
            b = self.rng.normal(0, 0.01, (1, self.filters))
        elif self.bias_init == "uniform":
            b = self.rng.uniform(-0.1, 0.1, (1, self.filters))
        elif self.bias_init == "small":
            b = np.full((1, self.filters), 0.01)
        else:
            raise ValueError("Invalid bias_init value. Possible values are 'normal', 'uniform', 'small' and 'default'.")
        self.weights, self.bias = w, b
        self.d_weights, self.d_bias = np.zeros_like(w), np.zeros((self.filters,))  # d_bias is (F,) after backward (Q9)

    def __str__(self) -> str:
        return (f"Conv2D(num_filters={self.filters}, kernel_size={self.kernel_size}, strides={self.strides}, "
                f"padding={self.padding})")

    def _geometry(self, shape):
        """layers.py:450-469."""
        N, H, W, Cin = (int(s) for s in shape)
        kh, kw = self.kernel_size
        sh, sw = self.strides
        if self.padding == "same":
            oh, ow = -(-H // sh), -(-W // sw)
            ph = max(0, (oh - 1) * sh + kh - H) // 2
            pw = max(0, (ow - 1) * sw + kw - W) // 2
        else:
            ph = pw = 0
        oh = (H + 2 * ph - kh) // sh + 1
        ow = (W + 2 * pw - kw) // sw + 1
        return B.ConvGeom(N, H, W, Cin, self.filters, kh, kw, sh, sw, ph, pw, oh, ow)

    def _out_shape(self, in_shape):
        g = self._geometry(in_shape)
        return (in_shape[0], g.OH, g.OW, self.filters)

    def _plan_forward(self, ctx, st, x, act_layer, training):
        assert x.dim() == 4, f"Conv2D input must be 4D (batch_size, height, width, channels), got {tuple(x.shape)}"
        if self._weights is None:
            self.initialize_weights(tuple(x.shape))
        if self._d_weights is None:
            self.d_weights = np.zeros(self._shapes["weights"]); self.d_bias = np.zeros((self.filters,))
        assert x.shape[3] == self._weights.shape[2], "Number of input channels must match"
        g = self._geometry(x.shape)
        if g.OH <= 0 or g.OW <= 0:
            raise ValueError(f"Size mismatch: Conv2D output would be {g.OH}x{g.OW}")
        math = _math_code()   # tensor-core modes apply to the geometries b200nn_conv2d_ws_bytes_math names; FFMA otherwise
        ctx.need_ws(ops.conv2d_ws_bytes(g, math))
        y = ctx.buf((x.shape[0], g.OH, g.OW, self.filters))
        kind, alpha = _act_of(act_layer)
        st.update(x=x, g=g, math=math)
        w, b = self._weights, self._bias
        ctx.emit(lambda: ops.conv2d_fwd(math, g, x, w, b, y, kind, alpha, ctx.ws), "conv2d_fwd", (x, w, b, y))
        return y

    def _plan_backward(self, ctx, st, err, mask_layer, need_dx):
        x, g = st["x"], st["g"]
        dx = ctx.buf(x.shape) if need_dx else None
        kind, alpha = _act_of(mask_layer)
        s0 = ctx.slot() if need_dx else None
        s1 = ctx.slot() if (need_dx and mask_layer is not None) else None
        w, dw, db, ss, chain, math, et = self._weights, self._d_weights, self._d_bias, ctx.ss, err.chain, st["math"], err.t
        ctx.emit(lambda: ops.conv2d_bwd(math, g, x, w, et, ss, chain, dx, dw, db, kind, alpha, s0, s1, ctx.ws), "conv2d_bwd",
                 (x, et, dw, db) + ((w, dx) if need_dx else ()))
        if not need_dx:
            return None
        return _Err(dx, tuple(s for s in (s0, s1) if s is not None))

    # -- fused first block: Conv2D(3x3 'same', C <= 3) [+ Activation] -> BatchNormalization -> MaxPooling2D(2) (csrc/conv_block.cu)
    def _can_fuse_block(self, x, act_layer, bn, pool, training):
        if not (training and type(self) is Conv2D and isinstance(bn, BatchNormalization) and isinstance(pool, MaxPooling2D)):
            return False
        if x.dim() != 4 or (act_layer is not None and not act_layer._fusable):
            return False
        if not (pool.pool_size == (2, 2) and pool.strides == (2, 2) and pool.padding == "valid"):
            return False
        kind, alpha = _act_of(act_layer)
        return ops.convblock_supported(self._geometry(x.shape), kind, alpha)

    def _plan_forward_block(self, ctx, st, st_bn, x, act_layer, bn):
        if self._weights is None:
            self.initialize_weights(tuple(x.shape))
        if self._d_weights is None:
            self.d_weights = np.zeros(self._shapes["weights"]); self.d_bias = np.zeros((self.filters,))
        assert x.shape[3] == self._weights.shape[2], "Number of input channels must match"
        g = self._geometry(x.shape)
        bn._ensure_params((g.OH, g.OW, self.filters))
        P = bn._gamma.numel()
        kind, alpha = _act_of(act_layer)
        ctx.need_ws(max(ops.convblock_ws_bytes(g), ops.conv2d_ws_bytes(g)))
        yp = ctx.buf((x.shape[0], g.OH // 2, g.OW // 2, self.filters))
        sm, si = ctx.buf((P,)), ctx.buf((P,))
        st.update(x=x, g=g, math=ops.MATH_FP32, block=True, act=(kind, alpha))
        st_bn.update(save_mean=sm, save_invstd=si, trained=True, fused_pool=True)
        w, b, gm, bt, rm, rv = self._weights, self._bias, bn._gamma, bn._beta, bn._running_mean, bn._running_var
        mom, eps = float(bn.momentum), float(bn.epsilon)
        peer = ctx.peer_region(ops.convblock_region_bytes(g, ctx.world)) if ctx.dp is not None else None
        if peer is not None:   # one kernel: the batch moments are merged across the GPUs over NVLink peer memory inside it
            b_total = x.shape[0] * ctx.world
            ctx.emit(lambda: ops.convblock_fwd(g, x, w, b, kind, alpha, gm, bt, rm, rv, sm, si, None, 3, b_total, yp, mom, eps, peer),
                     "convblock_fwd_peer", (x, w, yp))
        elif ctx.dp is not None:
            stats, b_total = ctx.buf64((2 * P,)), x.shape[0] * ctx.world
            ctx.emit(lambda: ops.convblock_fwd(g, x, w, b, kind, alpha, gm, bt, rm, rv, sm, si, stats, 1, b_total, None, mom, eps),
                     "convblock_stats", (x,))
            ctx.allreduce(stats, "bn_stats_exchange")
            ctx.emit(lambda: ops.convblock_fwd(g, x, w, b, kind, alpha, gm, bt, rm, rv, sm, si, stats, 2, b_total, yp, mom, eps),
                     "convblock_fwd_apply", (x, yp))
        else:
            ctx.emit(lambda: ops.convblock_fwd(g, x, w, b, kind, alpha, gm, bt, rm, rv, sm, si, None, 0, x.shape[0], yp, mom, eps),
                     "convblock_fwd", (x, w, yp))
        return yp

    def _plan_backward_block(self, ctx, st, st_bn, bn, err, has_act, need_dx):
        """err: error w.r.t. the pooled output. Emits pool-bwd + BN-bwd + act-bwd + conv weight gradient as one kernel (two
        under data parallelism); returns the error w.r.t. the conv INPUT when `need_dx`, else None."""
        x, g = st["x"], st["g"]
        kind, alpha = st["act"]
        s0, s1 = ctx.slot(), ctx.slot()
        s2 = ctx.slot() if has_act else None
        chain_out = ctx.pending_chain() or tuple(s for s in (s0, s1, s2) if s is not None)
        P = bn._gamma.numel()
        r_out = ctx.buf((x.shape[0], g.OH, g.OW, self.filters)) if need_dx else None
        w, b, dw, db, gm, bt = self._weights, self._bias, self._d_weights, self._d_bias, bn._gamma, bn._beta
        sm, si, ss, chain, et = st_bn["save_mean"], st_bn["save_invstd"], ctx.ss, err.chain, err.t
        bn._dgrad_pending = bn._pending_upto(ctx, s0)
        peer = ctx.peer_region(ops.convblock_region_bytes(g, ctx.world)) if ctx.dp is not None else None
        if peer is not None:
            b_total = x.shape[0] * ctx.world
            dgp, dbp = bn._d_gamma.data_ptr(), bn._d_beta.data_ptr()
            ctx.emit(lambda: ops.convblock_bwd(g, x, w, b, kind, alpha, gm, bt, sm, si, et, ss, chain, dgp, dbp, None, None, r_out, 3,
                                               b_total, s0, s1, s2, (), ctx.ws, peer), "convblock_bwd_peer", (x, et, r_out))
            if getattr(ctx, "_merge_final", False) and ctx.defer_clips and not need_dx:
                # last backward op of a train step: the reduce emits the RAW weight gradient and the optimizer applies the whole
                # pending chain (as for every other deferred gradient), so that the clip norms and this gradient can cross the GPUs
                # in ONE exchange instead of norms -> scaled reduce -> gradient
                st["wgrad_chain_deferred"] = chain_out
                ctx.emit(lambda: ops.convblock_wgrad_reduce(g, ctx.ws, ss, (), dw, db), "convblock_reduce", (dw,))
            else:
                ctx.sync_slots_now()   # the weight gradient owes the three clips just produced: their norms are global
                ctx.emit(lambda: ops.convblock_wgrad_reduce(g, ctx.ws, ss, chain_out, dw, db), "convblock_reduce", (dw,))
        elif ctx.dp is not None:
            sums, b_total = bn._dp_sums(ctx), x.shape[0] * ctx.world
            dgp, dbp = sums.data_ptr() + 4 * P, sums.data_ptr()
            ctx.emit(lambda: ops.convblock_bwd(g, x, w, b, kind, alpha, gm, bt, sm, si, et, ss, chain, dgp, dbp, None, None, None, 1,
                                               b_total, s0, None, None, (), None), "convblock_bwd_sums", (x, et))
            ctx.allreduce(sums, "bn_sums_exchange")
            ctx.emit(lambda: ops.convblock_bwd(g, x, w, b, kind, alpha, gm, bt, sm, si, et, ss, chain, dgp, dbp, None, None, r_out, 2,
                                               b_total, None, s1, s2, (), ctx.ws), "convblock_bwd_apply", (x, et, r_out))
            ctx.sync_slots_now()   # the weight gradient owes the three clips just produced: their norms are global
            ctx.emit(lambda: ops.convblock_wgrad_reduce(g, ctx.ws, ss, chain_out, dw, db), "convblock_reduce", (dw,))
        else:
            dgp, dbp = bn._d_gamma.data_ptr(), bn._d_beta.data_ptr()
            ctx.emit(lambda: ops.convblock_bwd(g, x, w, b, kind, alpha, gm, bt, sm, si, et, ss, chain, dgp, dbp, dw, db, r_out, 0,
                                               x.shape[0], s0, s1, s2, chain_out, ctx.ws), "convblock_bwd", (x, et, dw, r_out))
        if not need_dx:
            return None
        dx = ctx.buf(x.shape)
        s3 = ctx.slot()
        math = st["math"]
        chain_dx = () if ctx.defer_clips else chain_out
        ctx.emit(lambda: ops.conv2d_bwd(math, g, x, w, r_out, ss, chain_dx, dx, None, None, ops.ACT_NONE, 0.0, s3, None, ctx.ws),
                 "conv2d_dx", (r_out, w, dx))
        return _Err(dx, (s3,))

    def forward_pass(self, input_data):
        x, host = _in(input_data)
        self.input = input_data
        self._st = {}
        return _out(self._plan_forward(_EagerCtx(), self._st, x, None, True), host)

    def backward_pass(self, output_error):
        e, host = _in(output_error)
        return _out(self._plan_backward(_EagerCtx(), self._st, _Err(e), None, True).t, host)

    def get_config(self) -> dict:
        w, b = self.weights, self.bias
        return {"name": self.__class__.__name__, "weights": w.tolist() if w is not None else None,
                "bias": b.tolist() if b is not None else None, "filters": self.filters, "kernel_size": self.kernel_size,
                "strides": self.strides, "padding": self.padding, "weights_init": self.weights_init,
                "bias_init": self.bias_init, "random_state": self.random_state}

    @classmethod
    def from_config(cls, config: dict):
        layer = cls(config["filters"], config["kernel_size"], config["strides"], config["padding"], config["weights_init"],
                    config["bias_init"], config["random_state"])
        if config["weights"] is not None:
            layer.weights = np.array(config["weights"])
            layer.bias = np.array(config["bias"])
        return layer


# =======================================================================================================
class Conv2DTranspose(Conv2D):
    """layers.py:2536-2692. weights (kh, kw, F, C_in); bias and d_bias (1, 1, 1, F)."""

    def initialize_weights(self, input_shape: tuple):
        in_channels = input_shape[3]
        self.rng = np.random.default_rng(self.random_state if self.random_state is not None else int(time.time_ns()))
        shape = (*self.kernel_size, self.filters, in_channels)
        if self.weights_init == "xavier":
            w = self.rng.normal(0, np.sqrt(2 / (np.prod(self.kernel_size) * self.filters)), shape)
        elif self.weights_init == "he":
            w = self.rng.normal(0, np.sqrt(2 / (in_channels * np.prod(self.kernel_size))), shape)
        elif self.weights_init == "default":
            w = self.rng.normal(0, 0.01, shape)
        else:
            raise ValueError("Invalid weights_init value. Possible values are 'xavier', 'he', and 'default'.")
        bshape = (1, 1, 1, self.filters)
        if self.bias_init == "default":
            b = np.zeros(bshape)
        elif self.bias_init == "normal":
            b = self.rng.normal(0, 0.01, bshape)
        elif self.bias_init == "uniform":
            b = self.rng.uniform(-0.1, 0.1, bshape)
        elif self.bias_init == "small":
            b = np.full(bshape, 0.01)
        else:
            raise ValueError("Invalid bias_init value.")
        self.weights, self.bias = w, b
        self.d_weights, self.d_bias = np.zeros_like(w), np.zeros_like(b)

    def __str__(self) -> str:
        return (f"Conv2DTranspose(filters={self.filters}, kernel_size={self.kernel_size}, strides={self.strides}, "
                f"padding={self.padding})")

    def _geometry(self, shape):
        """layers.py:2597-2609."""
        N, H, W, Cin = (int(s) for s in shape)
        kh, kw = self.kernel_size
        sh, sw = self.strides
        if self.padding == "same":
            oh, ow = H * sh, W * sw
            pt = max((H - 1) * sh + kh - oh, 0) // 2
            pl = max((W - 1) * sw + kw - ow, 0) // 2
        else:
            oh, ow = (H - 1) * sh + kh, (W - 1) * sw + kw
            pt = pl = 0
        return B.ConvTGeom(N, H, W, Cin, self.filters, kh, kw, sh, sw, pt, pl, oh, ow)

    def _plan_forward(self, ctx, st, x, act_layer, training):
        assert x.dim() == 4, "Conv2DTranspose input must be 4D (batch_size, height, width, channels)"
        if self._weights is None:
            self.initialize_weights(tuple(x.shape))
        if self._d_weights is None:
            self.d_weights = np.zeros(self._shapes["weights"]); self.d_bias = np.zeros((1, 1, 1, self.filters))
        assert x.shape[3] == self._weights.shape[3], "Number of input channels must match"
        g = self._geometry(x.shape)
        math = _math_code()   # tensor-core modes: column buffer + tcgen05 contractions where eligible, FFMA otherwise
        ctx.need_ws(ops.convt_ws_bytes(g, math))
        y = ctx.buf((x.shape[0], g.OH, g.OW, self.filters))
        kind, alpha = _act_of(act_layer)
        st.update(x=x, g=g, math=math)
        w, b = self._weights, self._bias
        ctx.emit(lambda: ops.convt_fwd(math, g, x, w, b, y, kind, alpha, ctx.ws), "convt_fwd", (x, w, b, y))
        return y

    def _plan_backward(self, ctx, st, err, mask_layer, need_dx):
        x, g = st["x"], st["g"]
        dx = ctx.buf(x.shape) if need_dx else None
        kind, alpha = _act_of(mask_layer)
        s0 = ctx.slot() if need_dx else None
        s1 = ctx.slot() if (need_dx and mask_layer is not None) else None
        w, dw, db, ss, chain, math, et = self._weights, self._d_weights, self._d_bias, ctx.ss, err.chain, st["math"], err.t
        ctx.emit(lambda: ops.convt_bwd(math, g, x, w, et, ss, chain, dx, dw, db, kind, alpha, s0, s1, ctx.ws), "convt_bwd",
                 (x, et, dw, db) + ((w, dx) if need_dx else ()))
        if not need_dx:
            return None
        return _Err(dx, tuple(s for s in (s0, s1) if s is not None))


# =======================================================================================================
class MaxPooling2D(Layer):
    def __init__(self, pool_size, strides=None, padding: str = "valid"):
        self.pool_size = _pair(pool_size)
        self.strides = _pair(strides) if strides is not None else self.pool_size
        self.padding = padding

    def __str__(self) -> str:
        return f"MaxPooling2D(pool_size={self.pool_size}, strides={self.strides}, padding={self.padding})"

    def _geometry(self, shape):
        """layers.py:574-589."""
        N, H, W, Cc = (int(s) for s in shape)
        (ph, pw), (sh, sw) = self.pool_size, self.strides
        if self.padding == "same":
            pad_h = ((H - 1) * sh + ph - H) // 2
            pad_w = ((W - 1) * sw + pw - W) // 2
            if pad_h <= 0 or pad_w <= 0:
                # the reference crops d_input[pad:-pad], which is empty for pad == 0 (layers.py:639-641)
                raise ValueError("MaxPooling2D(padding='same') needs a positive pad; the reference backward breaks otherwise")
        else:
            pad_h = pad_w = 0
        oh = (H + 2 * pad_h - ph) // sh + 1
        ow = (W + 2 * pad_w - pw) // sw + 1
        return B.PoolGeom(N, H, W, Cc, ph, pw, sh, sw, pad_h, pad_w, oh, ow)

    def _out_shape(self, in_shape):
        g = self._geometry(in_shape)
        return (in_shape[0], g.OH, g.OW, in_shape[3])

    def _plan_forward(self, ctx, st, x, act_layer, training):
        assert x.dim() == 4, f"MaxPooling2D input must be 4D (batch_size, channels, height, width), got {tuple(x.shape)}"
        g = self._geometry(x.shape)
        y = ctx.buf((x.shape[0], g.OH, g.OW, x.shape[3]))
        st.update(x=x, y=y, g=g)
        ctx.emit(lambda: ops.maxpool_fwd(g, x, y), "maxpool_fwd", (x, y))
        return y

    def _plan_backward(self, ctx, st, err, mask_layer, need_dx):
        x, y, g = st["x"], st["y"], st["g"]
        dx = ctx.buf(x.shape)
        s = ctx.slot()
        ss, chain, et = ctx.ss, err.chain, err.t
        ctx.emit(lambda: ops.maxpool_bwd(g, x, y, et, ss, chain, dx, s), "maxpool_bwd", (x, y, et, dx))
        return _Err(dx, () if s is None else (s,))

    def forward_pass(self, input_data):
        x, host = _in(input_data)
        self.input = input_data
        self._st = {}
        return _out(self._plan_forward(_EagerCtx(), self._st, x, None, True), host)

    def backward_pass(self, output_error):
        e, host = _in(output_error)
        return _out(self._plan_backward(_EagerCtx(), self._st, _Err(e), None, True).t, host)

    def get_config(self) -> dict:
        return {"name": self.__class__.__name__, "pool_size": self.pool_size, "strides": self.strides, "padding": self.padding}

    @staticmethod
    def from_config(config: dict):
        return MaxPooling2D(config["pool_size"], config["strides"], config["padding"])


# =======================================================================================================
class Flatten(Layer):
    _is_view = True

    def __init__(self):
        pass

    def __str__(self) -> str:
        return "Flatten"

    def _out_shape(self, in_shape):
        return (in_shape[0], int(np.prod(in_shape[1:])))

    def forward_pass(self, input_data):
        assert len(input_data.shape) >= 2, f"Flatten input must be at least 2D, got {input_data.shape}"
        self.input_shape = tuple(input_data.shape)
        return input_data.reshape(input_data.shape[0], -1)

    def backward_pass(self, output_error):
        return output_error.reshape(self.input_shape)

    def get_config(self) -> dict:
        return {"name": self.__class__.__name__}

    @staticmethod
    def from_config(config: dict):
        return Flatten()


class Reshape(Layer):
    _is_view = True

    def __init__(self, target_shape: tuple, input_shape: tuple | None = None):
        super().__init__()
        self.target_shape = tuple(target_shape)
        self.input_shape = None

    def __str__(self) -> str:
        return f"Reshape(target_shape={self.target_shape}, input_shape={self.input_shape})"

    def _out_shape(self, in_shape):
        return (in_shape[0],) + self.target_shape

    def forward_pass(self, input_data):
        self.input_shape = tuple(input_data.shape)
        return input_data.reshape((input_data.shape[0],) + self.target_shape)

    def backward_pass(self, output_error):
        return output_error.reshape(self.input_shape)

    def get_config(self) -> dict:
        return {"name": self.__class__.__name__, "target_shape": self.target_shape, "input_shape": self.input_shape}

    @staticmethod
    def from_config(config: dict):
        return Reshape(config["target_shape"], config["input_shape"])


# =======================================================================================================
class BatchNormalization(Layer, _ParamMixin):
    """layers.py:1205-1304. Statistics are per POSITION (gamma/beta/running_* have the shape of one sample);
    has no `weights` attribute, so `Sequential` never updates gamma/beta."""
    _supports_mask = True
    gamma = _param_property("gamma")
    beta = _param_property("beta")

    def _dgrad_get(self, name):
        """d_gamma / d_beta as the reference leaves them. The fused pool+BN backward kernel accumulates them BEFORE the clip
        the reference applies between MaxPooling2D.backward_pass and BatchNormalization.backward_pass (models.py:214; the
        kernel carries that clip lazily as a slot), so the pending multiplier is applied when the attribute is read."""
        v = self._param_get(name)
        pend = getattr(self, "_dgrad_pending", None)
        if v is not None and pend is not None:
            v = v * _pending_factor(pend)
        return v

    d_gamma = property(lambda self: self._dgrad_get("d_gamma"), lambda self, v: self._param_set("d_gamma", v))
    d_beta = property(lambda self: self._dgrad_get("d_beta"), lambda self, v: self._param_set("d_beta", v))
    running_mean = _param_property("running_mean")
    running_var = _param_property("running_var")

    def __init__(self, momentum: float = 0.9, epsilon: float = 1e-5, **kwargs):
        self._gamma = self._beta = self._d_gamma = self._d_beta = self._running_mean = self._running_var = None
        self._shapes = {}
        self.momentum = momentum
        self.epsilon = epsilon
        for key, value in kwargs.items():
            setattr(self, key, value)

    def initialize_weights(self, input_shape: tuple):
        self.gamma, self.beta = np.ones(input_shape), np.zeros(input_shape)
        self.d_gamma, self.d_beta = np.zeros(input_shape), np.zeros(input_shape)
        self.running_mean, self.running_var = np.zeros(input_shape), np.ones(input_shape)

    def __str__(self) -> str:
        return f"BatchNormalization(momentum={self.momentum}, epsilon={self.epsilon})"

    def _can_fuse_pool(self, x, pool, training):
        """BatchNorm(train) -> MaxPool 2x2/s2 'valid' on an even, 32-channel-aligned NHWC tensor runs as one kernel."""
        # batches beyond the single-pass cluster kernels (fused_block_max_batch) take the streaming kernels, which have no batch limit
        return (training and isinstance(pool, MaxPooling2D) and x.dim() == 4 and pool.pool_size == (2, 2)
                and pool.strides == (2, 2) and pool.padding == "valid" and x.shape[1] % 2 == 0 and x.shape[2] % 2 == 0
                and x.shape[3] % 32 == 0 and (x.shape[0] <= ops.fused_block_max_batch() or x.shape[0] >= self._stream_rows()))

    @staticmethod
    def _stream_rows():
        """Batch rows from which the batch-coupled kernels run in their streaming form (csrc/bn_pool.cu): the single-read cluster
        kernels stage a window's whole batch through 128-byte lines 256 KB apart, which stops paying beyond ~256 rows."""
        return int(os.environ.get("B200NN_BN_SPLIT_ROWS", "512"))

    def _ensure_params(self, shp):
        if self._gamma is None or self._running_mean is None or self._running_var is None:
            self.initialize_weights(shp)
        if self._d_gamma is None:
            self.d_gamma, self.d_beta = np.zeros(shp), np.zeros(shp)
        if int(np.prod(shp)) != self._gamma.numel():
            raise ValueError(f"BatchNormalization was initialised for {tuple(self._shapes['gamma'])}, got {shp}")

    def _plan_forward(self, ctx, st, x, act_layer, training, pool=None):
        shp = tuple(int(s) for s in x.shape[1:])
        self._ensure_params(shp)
        P = self._gamma.numel()
        gm, bt, rm, rv = self._gamma, self._beta, self._running_mean, self._running_var
        mom, eps = float(self.momentum), float(self.epsilon)
        # Large batches on one GPU take the same two streaming kernels: the single-read cluster kernel stages a window's whole
        # batch in shared memory through 128-byte lines 256 KB apart, which stops paying beyond ~256 rows.
        big = self._stream_rows()
        if training and (ctx.dp is not None or x.shape[0] >= big):
            # data parallel (parallel.py): local moments -> sum over ranks -> apply, so that mean / var are those of the
            # GLOBAL batch (layers.py:1237-1238)
            b_total = x.shape[0] * ctx.world
            stats = ctx.buf64((2 * P,))
            sm, si = ctx.buf((P,)), ctx.buf((P,))
            st.update(x=x, save_mean=sm, save_invstd=si, trained=True, fused_pool=pool is not None)
            if x.shape[0] >= 128 and P % 4 == 0:   # one sweep over x (shifted sums per batch chunk, merged in fp64)
                ctx.need_ws(ops.bn_stats_stream_ws_bytes(x.shape[0], P))
                ctx.emit(lambda: ops.bn_stats_stream(x, stats, ctx.ws), "bn_stats_stream", (x,))
            else:
                ctx.emit(lambda: ops.bn_stats_partial(x, stats), "bn_stats_partial", (x,))
            ctx.allreduce(stats, "bn_stats_exchange")
            if pool is not None:
                yp = ctx.buf((x.shape[0], x.shape[1] // 2, x.shape[2] // 2, x.shape[3]))
                ctx.emit(lambda: ops.bn_pool_fwd_apply(x, gm, bt, rm, rv, sm, si, stats, yp, b_total, mom, eps),
                         "bn_pool_fwd_apply", (x, yp))
                return yp
            y = ctx.buf(x.shape)
            ctx.emit(lambda: ops.bn_fwd_apply(x, gm, bt, rm, rv, sm, si, stats, y, b_total, mom, eps), "bn_fwd_apply", (x, y))
            return y
        if pool is not None:
            yp = ctx.buf((x.shape[0], x.shape[1] // 2, x.shape[2] // 2, x.shape[3]))
            sm, si = ctx.buf((P,)), ctx.buf((P,))
            st.update(x=x, save_mean=sm, save_invstd=si, trained=True, fused_pool=True)
            ctx.emit(lambda: ops.bn_pool_fwd_train(x, gm, bt, rm, rv, sm, si, yp, mom, eps), "bn_pool_fwd_train", (x, yp))
            return yp
        y = ctx.buf(x.shape)
        if training:
            sm, si = ctx.buf((P,)), ctx.buf((P,))
            st.update(x=x, save_mean=sm, save_invstd=si, trained=True)
            ctx.emit(lambda: ops.bn_fwd_train(x, gm, bt, rm, rv, sm, si, y, mom, eps), "bn_fwd_train", (x, y))
        else:
            st.update(x=x, trained=False)
            ctx.emit(lambda: ops.bn_fwd_infer(x, gm, bt, rm, rv, y, eps), "bn_fwd_infer", (x, y))
        return y

    def _plan_backward(self, ctx, st, err, mask_layer, need_dx):
        if not st.get("trained"):
            raise NotImplementedError("BatchNormalization.backward_pass after an inference-mode forward relies on stale "
                                      "training caches in the reference (layers.py:1258-1276); not supported")
        x = st["x"]
        dx = ctx.buf(x.shape)
        kind, alpha = _act_of(mask_layer)
        pc = ctx.pending_chain()   # deferred clips: d_gamma / d_beta owe every slot recorded BEFORE this op
        self._dgrad_pending = (ctx.ss, pc[1], pc[2]) if pc is not None else None
        s0 = ctx.slot()
        s1 = ctx.slot() if mask_layer is not None else None
        gm, sm, si, ss, chain, et = self._gamma, st["save_mean"], st["save_invstd"], ctx.ss, err.chain, err.t
        if ctx.dp is not None:
            sums, b_total = self._dp_sums(ctx), x.shape[0] * ctx.world
            ctx.emit(lambda: ops.bn_bwd_partial(x, et, ss, chain, sm, si, sums), "bn_bwd_partial", (x, et))
            ctx.allreduce(sums, "bn_sums_exchange")
            ctx.emit(lambda: ops.bn_bwd_apply(x, et, ss, chain, gm, sm, si, dx, sums, b_total, kind, alpha, s0, s1), "bn_bwd_apply",
                     (x, et, dx))
            return _Err(dx, tuple(s for s in (s0, s1) if s is not None))
        dg, dbt = self._d_gamma, self._d_beta
        ctx.emit(lambda: ops.bn_bwd(x, et, ss, chain, gm, sm, si, dx, dg, dbt, kind, alpha, s0, s1), "bn_bwd", (x, et, dx))
        return _Err(dx, tuple(s for s in (s0, s1) if s is not None))

    @staticmethod
    def _pending_upto(ctx, s0):
        """d_gamma / d_beta of the fused pool+BN kernels are accumulated before the clip recorded in slot s0 (and, with
        deferred clips, before every earlier one)."""
        if ctx.ss is None or s0 is None:
            return None
        pc = ctx.pending_chain()
        return (ctx.ss, pc[1], s0 + 1 - pc[1]) if pc is not None else (ctx.ss, s0, 1)

    def _dp_sums(self, ctx):
        """[d_beta | d_gamma] as ONE float32 buffer (a single exchange sums both over the ranks); d_beta / d_gamma become
        views of it."""
        P = self._gamma.numel()
        sums = getattr(self, "_dp_sums_buf", None)
        if sums is None or sums.numel() != 2 * P:
            sums = self._dp_sums_buf = B.zeros((2 * P,))
            self._d_beta, self._d_gamma = sums[:P], sums[P:]
        return sums

    def _plan_backward_fused_pool(self, ctx, st, err, mask_layer):
        """err is the error w.r.t. the POOLED output; emits MaxPool-bwd + BatchNorm-bwd (+ Activation-bwd) as one
        kernel and returns the error w.r.t. the BatchNorm input with the 2 (or 3) clip slots the reference applies
        between those layers (models.py:214)."""
        x = st["x"]
        r0 = ctx.buf(x.shape)
        kind, alpha = _act_of(mask_layer)
        s0, s1 = ctx.slot(), ctx.slot()
        s2 = ctx.slot() if mask_layer is not None else None
        gm, bt, sm, si, ss, chain, et = self._gamma, self._beta, st["save_mean"], st["save_invstd"], ctx.ss, err.chain, err.t
        self._dgrad_pending = self._pending_upto(ctx, s0)
        if x.shape[0] >= self._stream_rows():
            # streaming form (any batch size): partial sums per batch chunk -> fixed-order reduce [-> sum over ranks] -> apply
            sums, b_total = self._dp_sums(ctx), x.shape[0] * ctx.world
            ctx.need_ws(ops.pool_bn_bwd_stream_ws_bytes(*x.shape))
            ctx.emit(lambda: ops.pool_bn_bwd_stream_partial(x, et, ss, chain, gm, bt, sm, si, sums, ctx.ws, s0),
                     "pool_bn_bwd_stream_partial", (x, et))
            ctx.allreduce(sums, "bn_sums_exchange")
            ctx.emit(lambda: ops.pool_bn_act_bwd_stream_apply(x, et, ss, chain, gm, bt, sm, si, r0, sums, b_total, kind, alpha, s1, s2),
                     "pool_bn_act_bwd_stream_apply", (x, et, r0))
            return _Err(r0, tuple(s for s in (s0, s1, s2) if s is not None))
        if ctx.dp is not None:
            sums, b_total = self._dp_sums(ctx), x.shape[0] * ctx.world
            ctx.emit(lambda: ops.pool_bn_bwd_partial(x, et, ss, chain, gm, bt, sm, si, sums, s0), "pool_bn_bwd_partial", (x, et))
            ctx.allreduce(sums, "bn_sums_exchange")
            ctx.emit(lambda: ops.pool_bn_act_bwd_apply(x, et, ss, chain, gm, bt, sm, si, r0, sums, b_total, kind, alpha, s1, s2),
                     "pool_bn_act_bwd_apply", (x, et, r0))
            return _Err(r0, tuple(s for s in (s0, s1, s2) if s is not None))
        dg, dbt = self._d_gamma, self._d_beta
        ctx.emit(lambda: ops.pool_bn_act_bwd(x, et, ss, chain, gm, bt, sm, si, r0, dg, dbt, kind, alpha, s0, s1, s2),
                 "pool_bn_act_bwd", (x, et, r0))
        return _Err(r0, tuple(s for s in (s0, s1, s2) if s is not None))

    def forward_pass(self, input_data, training: bool = True):
        x, host = _in(input_data)
        self.input = input_data
        self._st = {}
        return _out(self._plan_forward(_EagerCtx(), self._st, x, None, training), host)

    def backward_pass(self, output_error):
        e, host = _in(output_error)
        return _out(self._plan_backward(_EagerCtx(), self._st, _Err(e), None, True).t, host)

    def get_config(self) -> dict:
        g = self.gamma
        tl = lambda a: a.tolist() if a is not None else None  # noqa: E731
        return {"name": self.__class__.__name__, "gamma": tl(g), "beta": tl(self.beta), "momentum": self.momentum,
                "epsilon": self.epsilon, "running_mean": tl(self.running_mean), "running_var": tl(self.running_var),
                "input_shape": g.shape if g is not None else None}

    @staticmethod
    def from_config(config: dict):
        layer = BatchNormalization(config["momentum"], config["epsilon"])
        if config["gamma"] is not None:
            layer.gamma, layer.beta = np.array(config["gamma"]), np.array(config["beta"])
            layer.running_mean, layer.running_var = np.array(config["running_mean"]), np.array(config["running_var"])
            layer.d_gamma, layer.d_beta = np.zeros_like(layer.gamma), np.zeros_like(layer.beta)
        return layer



# =======================================================================================================
# Sibling layers of the same kernel families (csrc/layers_extra.cu)
class Dropout(Layer):
    """layers.py:264-342. The mask is drawn on the HOST with the reference's own generator call (default_rng(random_state)
    .binomial(1, 1 - rate, shape) / (1 - rate), re-seeded on every call: with a fixed seed the SAME mask every step) and
    applied on the device; inference is the identity."""

    def __init__(self, rate: float = 0.5, adaptive: bool = False, min_rate: float = 0.1, max_rate: float = 0.9,
                 temperature: float = 1.0, random_state: int = None, **kwargs):
        super().__init__()
        if adaptive:
            raise NotImplementedError("adaptive Dropout (layers.py:280-289) is outside the B200 hot path")
        self.rate, self.adaptive, self.random_state = rate, adaptive, random_state
        self.mask = None
        for key, value in kwargs.items():
            setattr(self, key, value)

    def __str__(self) -> str:
        return f"Dropout(rate={self.rate})"

    def _host_mask(self, shape):
        rng = np.random.default_rng(self.random_state if self.random_state is not None else int(time.time_ns()))
        self.mask = rng.binomial(1, 1 - self.rate, size=tuple(shape)) / (1 - self.rate)
        return self.mask

    def _plan_forward(self, ctx, st, x, act_layer, training):
        if not training:
            st.update(mask=None)
            return x
        mask = B.to_device(self._host_mask(x.shape))
        y = ctx.buf(x.shape)
        st.update(mask=mask)
        layer, fixed = self, self.random_state is not None

        def run():
            if not fixed:   # time-seeded: a fresh mask per step, uploaded outside any captured graph replay is impossible -> eager only
                mask.copy_(B.to_device(layer._host_mask(x.shape)))
            ops.mul(x, mask, None, (), y, None)
        ctx.emit(run, "dropout_fwd", (x, mask, y))
        return y

    def _plan_backward(self, ctx, st, err, mask_layer, need_dx):
        mask = st.get("mask")
        if mask is None:
            return err
        dx = ctx.buf(err.t.shape)
        s = ctx.slot()
        ss, chain, et = ctx.ss, err.chain, err.t
        ctx.emit(lambda: ops.mul(et, mask, ss, chain, dx, s), "dropout_bwd", (et, mask, dx))
        return _Err(dx, () if s is None else (s,))

    def forward_pass(self, input_data, training: bool = True):
        x, host = _in(input_data)
        self.input = input_data
        self._st = {}
        return _out(self._plan_forward(_EagerCtx(), self._st, x, None, training), host)

    def backward_pass(self, output_error):
        e, host = _in(output_error)
        return _out(self._plan_backward(_EagerCtx(), self._st, _Err(e), None, True).t, host)

    def get_config(self) -> dict:
        return {"name": self.__class__.__name__, "rate": self.rate, "adaptive": self.adaptive, "min_rate": 0.1, "max_rate": 0.9,
                "temperature": 1.0, "random_state": self.random_state}

    @staticmethod
    def from_config(config: dict):
        return Dropout(rate=config["rate"], adaptive=config["adaptive"], random_state=config["random_state"])


class AveragePooling2D(MaxPooling2D):
    """layers.py:908-1011: the mean over every window of the zero-padded input (padding counts in the divisor)."""

    def __str__(self) -> str:
        return f"AveragePooling2D(pool_size={self.pool_size}, strides={self.strides}, padding={self.padding})"

    def _plan_forward(self, ctx, st, x, act_layer, training):
        assert x.dim() == 4, f"AveragePooling2D input must be 4D (batch_size, height, width, channels), got {tuple(x.shape)}"
        g = self._geometry(x.shape)
        y = ctx.buf((x.shape[0], g.OH, g.OW, x.shape[3]))
        st.update(x=x, g=g)
        ctx.emit(lambda: ops.avgpool2d_fwd(g, x, y), "avgpool2d_fwd", (x, y))
        return y

    def _plan_backward(self, ctx, st, err, mask_layer, need_dx):
        x, g = st["x"], st["g"]
        dx = ctx.buf(x.shape)
        s = ctx.slot()
        ss, chain, et = ctx.ss, err.chain, err.t
        ctx.emit(lambda: ops.avgpool2d_bwd(g, et, ss, chain, dx, s), "avgpool2d_bwd", (et, dx))
        return _Err(dx, () if s is None else (s,))

    @staticmethod
    def from_config(config: dict):
        return AveragePooling2D(config["pool_size"], config["strides"], config["padding"])


class GlobalAveragePooling2D(Layer):
    """layers.py:1419-1442."""

    def __init__(self):
        self.input_shape = None

    def __str__(self) -> str:
        return "GlobalAveragePooling2D"

    def _out_shape(self, in_shape):
        return (in_shape[0], in_shape[3])

    def _plan_forward(self, ctx, st, x, act_layer, training):
        assert x.dim() == 4, f"GlobalAveragePooling2D input must be 4D (batch_size, height, width, channels), got {tuple(x.shape)}"
        self.input_shape = tuple(x.shape)
        y = ctx.buf((x.shape[0], x.shape[3]))
        st.update(in_shape=tuple(x.shape))
        ctx.emit(lambda: ops.gap2d_fwd(x, y), "gap2d_fwd", (x, y))
        return y

    def _plan_backward(self, ctx, st, err, mask_layer, need_dx):
        dx = ctx.buf(st["in_shape"])
        s = ctx.slot()
        ss, chain, et = ctx.ss, err.chain, err.t
        ctx.emit(lambda: ops.gap2d_bwd(et, ss, chain, dx, s), "gap2d_bwd", (et, dx))
        return _Err(dx, () if s is None else (s,))

    def forward_pass(self, input_data):
        x, host = _in(input_data)
        self._st = {}
        return _out(self._plan_forward(_EagerCtx(), self._st, x, None, True), host)

    def backward_pass(self, output_error):
        e, host = _in(output_error)
        return _out(self._plan_backward(_EagerCtx(), self._st, _Err(e), None, True).t, host)

    def get_config(self) -> dict:
        return {"name": self.__class__.__name__}

    @staticmethod
    def from_config(config: dict):
        return GlobalAveragePooling2D()


class UpSampling2D(Layer):
    """layers.py:2695-2815, nearest-neighbour interpolation (np.repeat along H then W; block sums backward)."""

    def __init__(self, size=(2, 2), interpolation="nearest", **kwargs):
        super().__init__()
        self.size = tuple(size) if isinstance(size, (list, tuple)) else (size, size)
        self.interpolation = interpolation
        if len(self.size) != 2:
            raise ValueError("Size must have exactly 2 elements.")
        if not all(isinstance(v, (int, np.integer)) and v > 0 for v in self.size):
            raise ValueError("Size elements must be positive integers.")
        if interpolation not in ["nearest", "bilinear", "bicubic"]:
            raise ValueError('interpolation must be one of "nearest", "bilinear", or "bicubic"')
        if interpolation != "nearest":
            raise NotImplementedError("UpSampling2D: only nearest-neighbour interpolation is on the B200 path")
        for key, value in kwargs.items():
            setattr(self, key, value)

    def __str__(self) -> str:
        return f"UpSampling2D(size={self.size}, interpolation={self.interpolation})"

    def _out_shape(self, in_shape):
        return (in_shape[0], in_shape[1] * self.size[0], in_shape[2] * self.size[1], in_shape[3])

    def _plan_forward(self, ctx, st, x, act_layer, training):
        assert x.dim() == 4, "UpSampling2D input must be 4D (batch_size, height, width, channels)"
        fh, fw = (int(v) for v in self.size)
        y = ctx.buf(self._out_shape(tuple(x.shape)))
        st.update(in_shape=tuple(x.shape))
        ctx.emit(lambda: ops.upsample2d_fwd(x, y, fh, fw), "upsample2d_fwd", (x, y))
        return y

    def _plan_backward(self, ctx, st, err, mask_layer, need_dx):
        fh, fw = (int(v) for v in self.size)
        dx = ctx.buf(st["in_shape"])
        s = ctx.slot()
        ss, chain, et = ctx.ss, err.chain, err.t
        ctx.emit(lambda: ops.upsample2d_bwd(et, ss, chain, dx, s, fh, fw), "upsample2d_bwd", (et, dx))
        return _Err(dx, () if s is None else (s,))

    def forward_pass(self, input_data):
        x, host = _in(input_data)
        self.input = input_data
        self._st = {}
        return _out(self._plan_forward(_EagerCtx(), self._st, x, None, True), host)

    def backward_pass(self, output_error):
        e, host = _in(output_error)
        return _out(self._plan_backward(_EagerCtx(), self._st, _Err(e), None, True).t, host)

    def get_config(self) -> dict:
        return {"name": self.__class__.__name__, "size": self.size, "interpolation": self.interpolation}

    @staticmethod
    def from_config(config: dict):
        return UpSampling2D(size=config["size"], interpolation=config["interpolation"])


class MaxPooling1D(Layer):
    """layers.py:819-905 on the 2-D pooling kernels (an image of height 1)."""

    def __init__(self, pool_size: int, strides: int = None, padding: str = "valid"):
        self.pool_size = pool_size
        self.strides = strides if strides is not None else pool_size
        self.padding = padding

    def __str__(self) -> str:
        return f"MaxPooling1D(pool_size={self.pool_size}, strides={self.strides}, padding={self.padding})"

    def _geometry(self, shape):
        N, L, Cc = (int(v) for v in shape)
        pad = ((L - 1) * self.strides + self.pool_size - L) // 2 if self.padding == "same" else 0
        if self.padding == "same" and pad <= 0:
            raise ValueError("MaxPooling1D(padding='same') needs a positive pad; the reference backward breaks otherwise")
        out_l = (L + 2 * pad - self.pool_size) // self.strides + 1
        return B.PoolGeom(N, 1, L, Cc, 1, int(self.pool_size), 1, int(self.strides), 0, pad, 1, out_l)

    def _out_shape(self, in_shape):
        return (in_shape[0], self._geometry(in_shape).OW, in_shape[2])

    def _plan_forward(self, ctx, st, x, act_layer, training):
        assert x.dim() == 3, f"MaxPooling1D input must be 3D (batch_size, length, channels), got {tuple(x.shape)}"
        g = self._geometry(x.shape)
        y = ctx.buf((x.shape[0], g.OW, x.shape[2]))
        st.update(x=x, y=y, g=g)
        ctx.emit(lambda: ops.maxpool_fwd(g, x, y), "maxpool1d_fwd", (x, y))
        return y

    def _plan_backward(self, ctx, st, err, mask_layer, need_dx):
        x, y, g = st["x"], st["y"], st["g"]
        dx = ctx.buf(x.shape)
        s = ctx.slot()
        ss, chain, et = ctx.ss, err.chain, err.t
        ctx.emit(lambda: ops.maxpool_bwd(g, x, y, et, ss, chain, dx, s), "maxpool1d_bwd", (x, y, et, dx))
        return _Err(dx, () if s is None else (s,))

    def forward_pass(self, input_data):
        x, host = _in(input_data)
        self.input = input_data
        self._st = {}
        return _out(self._plan_forward(_EagerCtx(), self._st, x, None, True), host)

    def backward_pass(self, output_error):
        e, host = _in(output_error)
        return _out(self._plan_backward(_EagerCtx(), self._st, _Err(e), None, True).t, host)

    def get_config(self) -> dict:
        return {"name": self.__class__.__name__, "pool_size": self.pool_size, "strides": self.strides, "padding": self.padding}

    @staticmethod
    def from_config(config: dict):
        return MaxPooling1D(config["pool_size"], config["strides"], config["padding"])


class Conv1D(Layer, _ParamMixin):
    """layers.py:667-816 as im2col_1d (preprocessing.py:128-156: columns ordered (tap, c), which is also the order of
    weights.reshape(-1, F) -- no K-order quirk in 1-D) + the Dense kernels (tcgen05 / FFMA by math mode and shape)."""
    _fuses_act = True
    weights = _param_property("weights")
    bias = _param_property("bias")
    d_weights = _param_property("d_weights")
    d_bias = _param_property("d_bias")

    def __init__(self, filters: int, kernel_size: int, strides: int = 1, padding: str = "valid", weights_init: str = "default",
                 bias_init: str = "default", random_state: int = None, **kwargs):
        self.filters, self.kernel_size, self.strides, self.padding = filters, kernel_size, strides, padding
        self._weights = self._bias = self._d_weights = self._d_bias = None
        self._shapes = {}
        self.weights_init, self.bias_init, self.random_state = weights_init, bias_init, random_state
        for key, value in kwargs.items():
            setattr(self, key, value)

    def __str__(self) -> str:
        return (f"Conv1D(num_filters={self.filters}, kernel_size={self.kernel_size}, strides={self.strides}, "
                f"padding={self.padding})")

    def initialize_weights(self, input_shape: tuple):
        _, _, in_channels = input_shape
        self.rng = np.random.default_rng(self.random_state if self.random_state is not None else int(time.time_ns()))
        shape = (self.kernel_size, in_channels, self.filters)
        if self.weights_init == "xavier":
            w = self.rng.normal(0, np.sqrt(2 / (self.kernel_size * in_channels)), shape)
        elif self.weights_init == "he":
            w = self.rng.normal(0, np.sqrt(2 / (in_channels * self.kernel_size)), shape)
        elif self.weights_init == "default":
            w = self.rng.normal(0, 0.01, shape)
        else:
            raise ValueError("Invalid weights_init value. Possible values are 'xavier', 'he', and 'default'.")
        if self.bias_init == "default":
            b = np.zeros((1, self.filters))
        elif self.bias_init == "normal":
            b = self.rng.normal(0, 0.01, (1, self.filters))
        elif self.bias_init == "uniform":
            b = self.rng.uniform(-0.1, 0.1, (1, self.filters))
        elif self.bias_init == "small":
            b = np.full((1, self.filters), 0.01)
        else:
            raise ValueError("Invalid bias_init value. Possible values are 'normal', 'uniform', 'small' and 'default'.")
        self.weights, self.bias = w, b
        self.d_weights, self.d_bias = np.zeros_like(w), np.zeros((self.filters,))   # d_bias is (F,) after backward (layers.py:801)

    def _geometry(self, shape):
        N, L, Cin = (int(v) for v in shape)
        k, s = int(self.kernel_size), int(self.strides)
        pad = ((L - 1) * s + k - L) // 2 if self.padding == "same" else 0
        out_l = (L + 2 * pad - k) // s + 1
        return B.ConvGeom(N, 1, L, Cin, self.filters, 1, k, 1, s, 0, pad, 1, out_l)

    def _out_shape(self, in_shape):
        return (in_shape[0], self._geometry(in_shape).OW, self.filters)

    def _plan_forward(self, ctx, st, x, act_layer, training):
        assert x.dim() == 3, f"Conv1D input must be 3D (batch_size, length, channels), got {tuple(x.shape)}"
        if self._weights is None:
            self.initialize_weights(tuple(x.shape))
        if self._d_weights is None:
            self.d_weights = np.zeros(self._shapes["weights"]); self.d_bias = np.zeros((self.filters,))
        assert x.shape[2] == self._weights.shape[1], "Number of input channels must match"
        g = self._geometry(x.shape)
        M, K = x.shape[0] * g.OW, int(self.kernel_size) * x.shape[2]
        math = _math_code()
        ctx.need_ws(ops.dense_ws_bytes(M, self.filters, K, math))
        col = ctx.buf((M, K))
        y = ctx.buf((x.shape[0], g.OW, self.filters))
        kind, alpha = _act_of(act_layer)
        st.update(x=x, g=g, col=col, math=math)
        w2, b = self._weights.view(K, self.filters), self._bias
        ctx.emit(lambda: ops.im2col(g, x, col), "im2col_1d", (x, col))
        ctx.emit(lambda: ops.dense_fwd(math, col, w2, b, y, kind, alpha, ctx.ws), "conv1d_fwd", (col, w2, b, y))
        return y

    def _plan_backward(self, ctx, st, err, mask_layer, need_dx):
        x, g, col, math = st["x"], st["g"], st["col"], st["math"]
        K = col.shape[1]
        dcol = ctx.buf(col.shape) if need_dx else None
        w2, dw2, db = self._weights.view(K, self.filters), self._d_weights.view(K, self.filters), self._d_bias
        ss, chain, et = ctx.ss, err.chain, err.t.reshape(col.shape[0], self.filters)
        ctx.emit(lambda: ops.dense_bwd(math, col, w2, et, ss, chain, dcol, dw2, db, ops.ACT_NONE, 0.0, None, None, ctx.ws), "conv1d_bwd",
                 (col, et, dw2, db) + ((w2, dcol) if need_dx else ()))
        if not need_dx:
            return None
        dx = ctx.buf(x.shape)
        s = ctx.slot()
        ctx.emit(lambda: ops.col2im(g, dcol, None, (), dx, s), "col2im_1d", (dcol, dx))   # the pending clips were applied to dcol
        return _Err(dx, () if s is None else (s,))

    def forward_pass(self, input_data):
        x, host = _in(input_data)
        self.input = input_data
        self._st = {}
        return _out(self._plan_forward(_EagerCtx(), self._st, x, None, True), host)

    def backward_pass(self, output_error):
        e, host = _in(output_error)
        return _out(self._plan_backward(_EagerCtx(), self._st, _Err(e), None, True).t, host)

    def get_config(self) -> dict:
        w, b = self.weights, self.bias
        return {"name": self.__class__.__name__, "weights": w.tolist() if w is not None else None,
                "bias": b.tolist() if b is not None else None, "filters": self.filters, "kernel_size": self.kernel_size,
                "strides": self.strides, "padding": self.padding, "weights_init": self.weights_init, "bias_init": self.bias_init,
                "random_state": self.random_state}

    @staticmethod
    def from_config(config: dict):
        layer = Conv1D(config["filters"], config["kernel_size"], config["strides"], config["padding"], config["weights_init"],
                       config["bias_init"], config["random_state"])
        if config["weights"] is not None:
            layer.weights = np.array(config["weights"])
            layer.bias = np.array(config["bias"])
        return layer


# =======================================================================================================
# layer-ordering validation used by Sequential.add (layers.py:3916-3980, restricted to the layers that exist here)
incompatibility_dict = {
    Input: [], Dense: [Conv1D, Conv2D, UpSampling2D], Activation: [], Conv2D: [Conv1D], UpSampling2D: [Conv1D],
    Conv2DTranspose: [Conv1D], MaxPooling2D: [Conv1D, MaxPooling1D], AveragePooling2D: [Conv1D, MaxPooling1D],
    GlobalAveragePooling2D: [Conv1D, MaxPooling1D],
    Conv1D: [Conv2D, UpSampling2D, Conv2DTranspose, MaxPooling2D, AveragePooling2D, GlobalAveragePooling2D],
    MaxPooling1D: [Conv2D, UpSampling2D, Conv2DTranspose, MaxPooling2D, AveragePooling2D, GlobalAveragePooling2D],
    Flatten: [], Dropout: [], BatchNormalization: [], Reshape: [],
}
