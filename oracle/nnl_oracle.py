"""TEST INFRASTRUCTURE ONLY — float64 NumPy restatement of neuralnetlib's `Sequential.fit` hot path.

This file is the *oracle* for the CUDA backend in `neuralnetlib_b200/`: a from-scratch restatement
of what the upstream reference computes on the path
`Sequential.train_on_batch -> forward_pass -> loss -> backward_pass (+clip) -> Optimizer.update`
for Dense / Conv2D / Conv2DTranspose / BatchNormalization / MaxPooling2D / Flatten / Reshape /
activations / CCE+BCE+MSE / SGD+Adam, and `GAN.train_on_batch`.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import it. The product (`neuralnetlib_b200`) never does.

Pinning: the upstream reference holds no golden vectors for Conv2D/BN/MaxPool/Conv2DTranspose/clip
(SURVEY.md §8c). The oracle is therefore pinned by *executing the unmodified reference* in the
build container (`oracle/make_golden.py`, fixtures in `tests/golden/*.npz`) and by the reference's
own closed-form tests (tests/test_optimizers.py:16-52, tests/test_losses.py:11-33,
tests/test_activations.py:10-59), which `tests/test_oracle_golden.py` replays against this file.

Every function cites the reference lines (relative to /root/reference/neuralnetlib/) it restates.
All arithmetic is float64, as in the reference.
"""
from __future__ import annotations

import numpy as np

# --------------------------------------------------------------------------------------------
# activations  (activations.py:33-109)
# --------------------------------------------------------------------------------------------


def act_fwd(name: str, x: np.ndarray, alpha: float = 0.01) -> np.ndarray:
    """activations.py:34-36 (sigmoid, clip +-50), :47-48 (relu), :58-59 (tanh), :69-71 (softmax, axis 1),
    :82-83 (linear), :96-97 (leakyrelu = max(x, alpha*x))."""
    if name == "sigmoid":
        return 1.0 / (1.0 + np.exp(-np.clip(x, -50.0, 50.0)))
    if name == "relu":
        return np.maximum(0.0, x)
    if name == "tanh":
        return np.tanh(x)
    if name == "softmax":
        e = np.exp(x - x.max(axis=1, keepdims=True))
        return e / e.sum(axis=1, keepdims=True)
    if name == "linear":
        return x
    if name == "leakyrelu":
        return np.maximum(x, x * alpha)
    raise ValueError(name)


def act_bwd(name: str, x: np.ndarray, err: np.ndarray, alpha: float = 0.01) -> np.ndarray:
    """layers.py:236-237 (err * f'(cached pre-activation input)); derivatives activations.py:38-40,
    :50-51, :61-62, :85-86, :99-100. Softmax has no derivative (:73-75)."""
    if name == "sigmoid":
        s = act_fwd("sigmoid", x)
        return err * (s * (1.0 - s))
    if name == "relu":
        return err * (x > 0)
    if name == "tanh":
        return err * (1.0 - np.tanh(x) ** 2)
    if name == "linear":
        return err * 1.0
    if name == "leakyrelu":
        return err * np.where(x > 0, 1.0, alpha)
    raise NotImplementedError(name)


# --------------------------------------------------------------------------------------------
# losses  (losses.py:79-120)
# --------------------------------------------------------------------------------------------
_EPS = 1e-15


def loss_value(name: str, y: np.ndarray, p: np.ndarray) -> float:
    """losses.py:80-81 (mse), :91-94 (bce), :106-109 (cce)."""
    if name == "mse":
        return float(np.mean((y - p) ** 2))
    pc = np.clip(p, _EPS, 1.0 - _EPS)
    if name == "bce":
        return float(-np.mean(y * np.log(pc) + (1.0 - y) * np.log(1.0 - pc)))
    if name == "cce":
        return float(-np.sum(y * np.log(pc)) / y.shape[0])
    raise ValueError(name)


def loss_derivative(name: str, y: np.ndarray, p: np.ndarray) -> np.ndarray:
    """losses.py:83-84 (mse), :96-99 (bce), :111-115 (cce)."""
    if name == "mse":
        return 2.0 * (p - y) / y.size
    pc = np.clip(p, _EPS, 1.0 - _EPS)
    if name == "bce":
        return (pc - y) / (pc * (1.0 - pc))
    if name == "cce":
        return -y / pc
    raise ValueError(name)


# --------------------------------------------------------------------------------------------
# clip  (preprocessing.py:220-224) — whole-tensor L2, threshold fixed at 1.0 on this path
# --------------------------------------------------------------------------------------------


def clip_l2(g: np.ndarray, threshold: float = 1.0) -> np.ndarray:
    n = np.sqrt(np.sum(np.square(g)))
    return threshold * g / n if n > threshold else g


# --------------------------------------------------------------------------------------------
# Dense  (layers.py:152-198)
# --------------------------------------------------------------------------------------------


def dense_fwd(x, w, b):
    """layers.py:173."""
    return x @ w + b


def dense_bwd(x, w, err):
    """layers.py:189 (dX), :196 (dW), :197 (db keepdims -> (1, units))."""
    return err @ w.T, x.T @ err, err.sum(axis=0, keepdims=True)


# --------------------------------------------------------------------------------------------
# Conv2D  (layers.py:442-530, preprocessing.py:34-126)
# --------------------------------------------------------------------------------------------


def conv2d_geometry(H, W, kh, kw, sh, sw, padding):
    """layers.py:450-469: 'same' total pad = max(0,(ceil(H/s)-1)*s+k-H), split //2 symmetric,
    out recomputed from the symmetric pad (so k3 s2 on even H gives 28 -> 13)."""
    if padding == "same":
        oh = -(-H // sh)
        ow = -(-W // sw)
        ph = max(0, (oh - 1) * sh + kh - H) // 2
        pw = max(0, (ow - 1) * sw + kw - W) // 2
    else:
        ph = pw = 0
    oh = (H + 2 * ph - kh) // sh + 1
    ow = (W + 2 * pw - kw) // sw + 1
    return oh, ow, ph, pw


def _patches(x, kh, kw, sh, sw, ph, pw, oh, ow):
    """(N, oh, ow, kh, kw, C) window view of the zero-padded input (preprocessing.py:64-75)."""
    xp = np.pad(x, ((0, 0), (ph, ph), (pw, pw), (0, 0)))
    N, Hp, Wp, C = xp.shape
    s = xp.strides
    return np.lib.stride_tricks.as_strided(
        xp, (N, oh, ow, kh, kw, C), (s[0], s[1] * sh, s[2] * sw, s[1], s[2], s[3]), writeable=False)


def conv2d_effective_kernel(w):
    """Q1: im2col's K order is (c, ky, kx) (preprocessing.py:77) while `weights.reshape(-1, F)`
    (layers.py:474) enumerates (ky, kx, c): the flat buffer is *read* as (C, kh, kw, F)."""
    kh, kw, C, F = w.shape
    return w.reshape(C, kh, kw, F)


def conv2d_fwd(x, w, b, strides, padding):
    """layers.py:471-488."""
    N, H, W, C = x.shape
    kh, kw, _, F = w.shape
    sh, sw = strides
    oh, ow, ph, pw = conv2d_geometry(H, W, kh, kw, sh, sw, padding)
    pat = _patches(x, kh, kw, sh, sw, ph, pw, oh, ow)
    weff = conv2d_effective_kernel(w)  # [c, ky, kx, f]
    out = np.einsum("nyxijc,cijf->nyxf", pat, weff, optimize=True)
    return out + b.reshape(1, 1, 1, F)


def conv2d_bwd(x, w, err, strides, padding):
    """layers.py:513-530: d_bias (F,), d_weights via the same K-order reinterpretation, d_input = col2im."""
    N, H, W, C = x.shape
    kh, kw, _, F = w.shape
    sh, sw = strides
    oh, ow, ph, pw = conv2d_geometry(H, W, kh, kw, sh, sw, padding)
    assert err.shape == (N, oh, ow, F)
    pat = _patches(x, kh, kw, sh, sw, ph, pw, oh, ow)
    weff = conv2d_effective_kernel(w)
    d_bias = err.reshape(-1, F).sum(axis=0)
    dweff = np.einsum("nyxijc,nyxf->cijf", pat, err, optimize=True)
    d_weights = dweff.reshape(kh, kw, C, F)  # flat order (c,ky,kx,f) reinterpreted, layers.py:523-524
    dpat = np.einsum("nyxf,cijf->nyxijc", err, weff, optimize=True)
    # col2im_2d (preprocessing.py:113-126): scatter-add into the padded image, then crop the pad.
    img = np.zeros((N, H + 2 * ph + sh - 1, W + 2 * pw + sw - 1, C))
    for i in range(kh):
        for j in range(kw):
            img[:, i:i + sh * oh:sh, j:j + sw * ow:sw, :] += dpat[:, :, :, i, j, :]
    d_input = img[:, ph:ph + H, pw:pw + W, :]
    return d_input, d_weights, d_bias


# --------------------------------------------------------------------------------------------
# MaxPooling2D  (layers.py:570-643)
# --------------------------------------------------------------------------------------------


def _pool_geometry(H, W, ph_, pw_, sh, sw, padding):
    """layers.py:574-589."""
    if padding == "same":
        pad_h = ((H - 1) * sh + ph_ - H) // 2
        pad_w = ((W - 1) * sw + pw_ - W) // 2
    else:
        pad_h = pad_w = 0
    oh = (H + 2 * pad_h - ph_) // sh + 1
    ow = (W + 2 * pad_w - pw_) // sw + 1
    return oh, ow, pad_h, pad_w


def maxpool_fwd(x, pool, strides, padding="valid"):
    """layers.py:582-600: zero padding (value 0, not -inf), window max."""
    N, H, W, C = x.shape
    oh, ow, pad_h, pad_w = _pool_geometry(H, W, pool[0], pool[1], strides[0], strides[1], padding)
    pat = _patches(x, pool[0], pool[1], strides[0], strides[1], pad_h, pad_w, oh, ow)
    return pat.max(axis=(3, 4))


def maxpool_bwd(x, err, pool, strides, padding="valid"):
    """layers.py:618-641: every element equal to the window max receives the full gradient (ties!),
    overlapping windows accumulate."""
    N, H, W, C = x.shape
    sh, sw = strides
    oh, ow, pad_h, pad_w = _pool_geometry(H, W, pool[0], pool[1], sh, sw, padding)
    xp = np.pad(x, ((0, 0), (pad_h, pad_h), (pad_w, pad_w), (0, 0)))
    pat = _patches(x, pool[0], pool[1], sh, sw, pad_h, pad_w, oh, ow)
    mx = pat.max(axis=(3, 4), keepdims=True)
    contrib = (pat == mx) * err[:, :, :, None, None, :]
    d = np.zeros_like(xp)
    for i in range(pool[0]):
        for j in range(pool[1]):
            d[:, i:i + sh * oh:sh, j:j + sw * ow:sw, :] += contrib[:, :, :, i, j, :]
    if padding == "same":
        # layers.py:639-641 — the reference slices [pad:-pad]; with pad == 0 that yields an empty
        # array (upstream bug). Restated literally.
        d = d[:, pad_h:-pad_h, pad_w:-pad_w, :]
    return d


# --------------------------------------------------------------------------------------------
# BatchNormalization  (layers.py:1230-1276) — statistics per (h, w, c) POSITION over axis 0
# --------------------------------------------------------------------------------------------


def bn_fwd(x, gamma, beta, running_mean, running_var, momentum, eps, training=True):
    """layers.py:1234 (clip +-10), :1237-1238 (batch stats, var + eps), :1240-1243 (running),
    :1252 (std = sqrt(var + eps) => raw var + 2 eps in training), :1253-1256."""
    xc_in = np.clip(x, -10.0, 10.0)
    cache = {}
    if training:
        mean = xc_in.mean(axis=0)
        var = xc_in.var(axis=0) + eps
        running_mean = momentum * running_mean + (1.0 - momentum) * mean
        running_var = momentum * running_var + (1.0 - momentum) * var
        cache["batch_var"] = var
    else:
        mean, var = running_mean, running_var
    std = np.sqrt(var + eps)
    centered = xc_in - mean
    xhat = centered / std
    cache.update(std=std, centered=centered, xhat=xhat, mean=mean)
    return gamma * xhat + beta, running_mean, running_var, cache


def bn_bwd(err, gamma, cache, eps):
    """layers.py:1259-1276 (the +-10 clip is not differentiated)."""
    N = err.shape[0]
    d_gamma = np.sum(err * cache["xhat"], axis=0)
    d_beta = np.sum(err, axis=0)
    dn = err * gamma
    d_var = np.sum(dn * cache["centered"] * -0.5 * (cache["batch_var"] + eps) ** (-1.5), axis=0)
    d_mean = np.sum(dn * -1.0 / cache["std"], axis=0) + d_var * np.mean(-2.0 * cache["centered"], axis=0)
    d_input = dn / cache["std"] + d_var * 2.0 * cache["centered"] / N + d_mean / N
    return d_input, d_gamma, d_beta


# --------------------------------------------------------------------------------------------
# Conv2DTranspose  (layers.py:2588-2659) — weights (kh, kw, F, C_in), bias (1,1,1,F)
# --------------------------------------------------------------------------------------------


def convT_geometry(H, W, kh, kw, sh, sw, padding):
    """layers.py:2597-2609."""
    if padding == "same":
        oh, ow = H * sh, W * sw
        pad_h = max((H - 1) * sh + kh - oh, 0)
        pad_w = max((W - 1) * sw + kw - ow, 0)
        pt, pl = pad_h // 2, pad_w // 2
        pb, pr = pad_h - pt, pad_w - pl
    else:
        oh, ow = (H - 1) * sh + kh, (W - 1) * sw + kw
        pt = pb = pl = pr = 0
    return oh, ow, pt, pb, pl, pr


def convT_fwd(x, w, b, strides, padding):
    """layers.py:2611-2633: scatter-add W[i,j,f,c]*x[n,h,w,c] at (h*s+i, w*s+j), crop, add bias."""
    N, H, W, C = x.shape
    kh, kw, F, _ = w.shape
    sh, sw = strides
    oh, ow, pt, pb, pl, pr = convT_geometry(H, W, kh, kw, sh, sw, padding)
    full = np.zeros((N, oh + pt + pb, ow + pl + pr, F))
    for i in range(kh):
        for j in range(kw):
            # full extent of strided slice may overrun when padded_out is smaller than (H-1)*s+k
            hs = full[:, i:i + sh * H:sh, j:j + sw * W:sw, :]
            contrib = np.einsum("nhwc,fc->nhwf", x, w[i, j], optimize=True)
            hs += contrib[:, :hs.shape[1], :hs.shape[2], :]
    out = full if padding == "valid" else full[:, pt:pt + oh, pl:pl + ow, :]
    return out + b.reshape(1, 1, 1, F)


def convT_bwd(x, w, err, strides):
    """layers.py:2635-2659: error windows are taken from the (cropped) error WITHOUT pad offset and a
    window truncated by the border is skipped entirely (Q12)."""
    N, H, W, C = x.shape
    kh, kw, F, _ = w.shape
    sh, sw = strides
    EH, EW = err.shape[1], err.shape[2]
    d_input = np.zeros_like(x, dtype=np.float64)
    d_weights = np.zeros_like(w, dtype=np.float64)
    d_bias = err.sum(axis=(0, 1, 2), keepdims=True)
    # valid input pixels: full kh x kw window inside the error tensor
    hv = [h for h in range(H) if h * sh + kh <= EH]
    wv = [q for q in range(W) if q * sw + kw <= EW]
    if hv and wv:
        nh, nw = len(hv), len(wv)  # contiguous prefixes
        for i in range(kh):
            for j in range(kw):
                e = err[:, i:i + sh * nh:sh, j:j + sw * nw:sw, :]  # (N, nh, nw, F)
                d_input[:, :nh, :nw, :] += np.einsum("nhwf,fc->nhwc", e, w[i, j], optimize=True)
                d_weights[i, j] = np.einsum("nhwf,nhwc->fc", e, x[:, :nh, :nw, :], optimize=True)
    return d_input, d_weights, d_bias


# --------------------------------------------------------------------------------------------
# optimizers  (optimizers.py:47-50, 204-245)
# --------------------------------------------------------------------------------------------


class OracleSGD:
    def __init__(self, learning_rate=0.01):
        self.learning_rate = learning_rate

    def update(self, idx, w, dw, b=None, db=None):
        w -= self.learning_rate * dw
        if b is not None and db is not None:
            b -= self.learning_rate * db


class OracleAdam:
    """optimizers.py:167-245. `t` is incremented once per update() call, i.e. per LAYER per step (Q3)."""

    def __init__(self, learning_rate=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-8):
        self.learning_rate, self.beta_1, self.beta_2, self.epsilon = learning_rate, beta_1, beta_2, epsilon
        self.t = 0
        self.m_w, self.v_w, self.m_b, self.v_b = {}, {}, {}, {}

    def _step(self, p, g, m, v):
        m = self.beta_1 * m + (1 - self.beta_1) * g
        v = self.beta_2 * v + (1 - self.beta_2) * np.square(g)
        m_hat = m / (1 - self.beta_1 ** self.t)
        v_hat = v / (1 - self.beta_2 ** self.t)
        upd = self.learning_rate * m_hat / np.maximum(np.sqrt(v_hat) + self.epsilon, 1e-16)
        upd = np.nan_to_num(upd, nan=0.0, posinf=0.0, neginf=0.0)
        p -= upd
        return m, v

    def update(self, idx, w, dw, b=None, db=None):
        if idx not in self.m_w:
            self.m_w[idx], self.v_w[idx] = np.zeros_like(w), np.zeros_like(w)
            if b is not None:
                self.m_b[idx], self.v_b[idx] = np.zeros_like(b), np.zeros_like(b)
        self.t += 1
        self.m_w[idx], self.v_w[idx] = self._step(w, dw, self.m_w[idx], self.v_w[idx])
        if b is not None and db is not None:
            self.m_b[idx], self.v_b[idx] = self._step(b, db, self.m_b[idx], self.v_b[idx])


# --------------------------------------------------------------------------------------------
# weight initialisation (host side, identical generators) — layers.py:89-150, 367-398, 2555-2586
# --------------------------------------------------------------------------------------------


def init_dense(in_dim, units, seed, weights_init="glorot_uniform"):
    rng = np.random.default_rng(seed)
    if weights_init == "glorot_uniform":
        lim = np.sqrt(6 / (in_dim + units))
        w = rng.uniform(-lim, lim, (in_dim, units))
    elif weights_init == "he_normal":
        w = rng.normal(0, np.sqrt(2 / in_dim), (in_dim, units))
    else:
        raise ValueError(weights_init)
    return w, np.zeros((1, units))


def init_conv2d(in_ch, filters, ksize, seed):
    rng = np.random.default_rng(seed)
    return rng.normal(0, 0.01, (ksize[0], ksize[1], in_ch, filters)), np.zeros((1, filters))


def init_convT(in_ch, filters, ksize, seed):
    rng = np.random.default_rng(seed)
    return rng.normal(0, 0.01, (ksize[0], ksize[1], filters, in_ch)), np.zeros((1, 1, 1, filters))


# --------------------------------------------------------------------------------------------
# Sequential  (models.py:159-264)
# --------------------------------------------------------------------------------------------


def _pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


class OracleSequential:
    """Layer specs are dicts: {'type': 'dense', 'units': 64}, {'type': 'act', 'fn': 'relu'},
    {'type': 'conv2d', 'filters': 32, 'kernel': 3, 'strides': 1, 'padding': 'same'},
    {'type': 'convT', ...}, {'type': 'bn'}, {'type': 'maxpool', 'pool': 2}, {'type': 'flatten'},
    {'type': 'reshape', 'shape': (7, 7, 128)}, {'type': 'input'}.
    Weights are created lazily at first forward from `seed` (models.py:116-117: every layer gets the
    model's random_state => same-shaped layers start identical, Q8)."""

    def __init__(self, specs, loss="cce", optimizer=None, seed=42):
        self.layers = [dict(s) for s in specs]
        self.loss = loss
        self.optimizer = optimizer if optimizer is not None else OracleAdam()
        self.seed = seed
        self.predictions = None
        self.y_true = None

    # -- forward: models.py:183-193
    def forward(self, x, training=True):
        x = np.asarray(x, dtype=np.float64)
        for L in self.layers:
            t = L["type"]
            if t == "input":
                pass
            elif t == "dense":
                if "w" not in L:
                    L["w"], L["b"] = init_dense(x.shape[1], L["units"], self.seed, L.get("init", "glorot_uniform"))
                L["x"] = x
                x = dense_fwd(x, L["w"], L["b"])
            elif t == "act":
                L["x"] = x
                x = act_fwd(L["fn"], x, L.get("alpha", 0.01))
            elif t == "conv2d":
                k, s = _pair(L["kernel"]), _pair(L.get("strides", 1))
                if "w" not in L:
                    L["w"], L["b"] = init_conv2d(x.shape[3], L["filters"], k, self.seed)
                L["x"] = x
                x = conv2d_fwd(x, L["w"], L["b"], s, L.get("padding", "valid"))
            elif t == "convT":
                k, s = _pair(L["kernel"]), _pair(L.get("strides", 1))
                if "w" not in L:
                    L["w"], L["b"] = init_convT(x.shape[3], L["filters"], k, self.seed)
                L["x"] = x
                x = convT_fwd(x, L["w"], L["b"], s, L.get("padding", "valid"))
            elif t == "bn":
                if "gamma" not in L:
                    shp = x.shape[1:]
                    L["gamma"], L["beta"] = np.ones(shp), np.zeros(shp)
                    L["running_mean"], L["running_var"] = np.zeros(shp), np.ones(shp)
                x, L["running_mean"], L["running_var"], L["cache"] = bn_fwd(
                    x, L["gamma"], L["beta"], L["running_mean"], L["running_var"],
                    L.get("momentum", 0.9), L.get("eps", 1e-5), training)
            elif t == "maxpool":
                p = _pair(L["pool"])
                s = _pair(L["strides"]) if L.get("strides") is not None else p
                L["x"] = x
                x = maxpool_fwd(x, p, s, L.get("padding", "valid"))
            elif t == "flatten":
                L["shape"] = x.shape
                x = x.reshape(x.shape[0], -1)
            elif t == "reshape":
                L["in_shape"] = x.shape
                x = x.reshape((x.shape[0],) + tuple(L["shape"]))
            else:
                raise ValueError(t)
        self.predictions = x
        return x

    def _layer_bwd(self, L, err):
        t = L["type"]
        if t == "input":
            return err
        if t == "dense":
            dx, L["dw"], L["db"] = dense_bwd(L["x"], L["w"], err)
            return dx
        if t == "act":
            return act_bwd(L["fn"], L["x"], err, L.get("alpha", 0.01))
        if t == "conv2d":
            dx, L["dw"], L["db"] = conv2d_bwd(L["x"], L["w"], err, _pair(L.get("strides", 1)), L.get("padding", "valid"))
            return dx
        if t == "convT":
            dx, L["dw"], L["db"] = convT_bwd(L["x"], L["w"], err, _pair(L.get("strides", 1)))
            return dx
        if t == "bn":
            dx, L["dgamma"], L["dbeta"] = bn_bwd(err, L["gamma"], L["cache"], L.get("eps", 1e-5))
            return dx
        if t == "maxpool":
            p = _pair(L["pool"])
            s = _pair(L["strides"]) if L.get("strides") is not None else p
            return maxpool_bwd(L["x"], err, p, s, L.get("padding", "valid"))
        if t == "flatten":
            return err.reshape(L["shape"])
        if t == "reshape":
            return err.reshape(L["in_shape"])
        raise ValueError(t)

    # -- backward: models.py:195-249
    def backward(self, error, gan=False, compute_only=False, trace=None):
        n = len(self.layers)
        for i, L in enumerate(reversed(self.layers)):
            if i == 0 and L["type"] == "act":
                if not gan:
                    if (L["fn"] == "softmax" and self.loss == "cce") or (L["fn"] == "sigmoid" and self.loss == "bce"):
                        error = self.predictions - self.y_true  # models.py:200-205 (no /B, Q4)
                    # else: activation derivative skipped, raw loss derivative flows on (Q7)
                else:
                    error = self._layer_bwd(L, error)  # models.py:211-212
            else:
                error = clip_l2(error)  # models.py:214
                error = self._layer_bwd(L, error)
            if trace is not None:
                trace.append(error)
            if compute_only:
                continue
            if "w" in L:  # models.py:239-244 — BN has no `weights` attribute: gamma/beta frozen (Q5)
                self.optimizer.update(n - 1 - i, L["w"], clip_l2(L["dw"]), L["b"], clip_l2(L["db"]))
        return error

    # -- models.py:251-264
    def train_on_batch(self, x, y, trace=None):
        y = np.asarray(y, dtype=np.float64)
        self.y_true = y
        p = self.forward(x, training=True)
        loss = loss_value(self.loss, y, p.copy())
        err = loss_derivative(self.loss, y, p.copy())
        if err.ndim == 1:
            err = err[:, None]
        self.backward(err, trace=trace)
        return loss


# --------------------------------------------------------------------------------------------
# GAN.train_on_batch  (models.py:2609-2667)
# --------------------------------------------------------------------------------------------


def gan_train_on_batch(G: OracleSequential, D: OracleSequential, real, batch_size, latent_dim, seed,
                       n_critic=1, label_smoothing=0.9):
    """models.py:2616 (rng re-seeded identically every call), :2620 (choice without replacement),
    :2522-2565 (latent noise re-seeded identically every call), :2634-2636 (labels 0.9 / 0.1),
    :2639-2642 (D step), :2648-2665 (G step through D with compute_only)."""
    rng = np.random.default_rng(seed)
    d_total = 0.0
    for _ in range(n_critic):
        idx = rng.choice(len(real), batch_size, replace=False)
        real_b = real[idx]
        z = np.random.default_rng(seed).normal(0, 1, (batch_size, latent_dim))
        fake = G.forward(z, training=False)
        xb = np.concatenate([real_b, fake])
        yb = np.zeros((2 * batch_size, 1))
        yb[:batch_size] = label_smoothing
        yb[batch_size:] = 1 - label_smoothing
        D.y_true = yb
        p = D.forward(xb, training=True)
        d_total += loss_value(D.loss, yb, p)
        D.backward(loss_derivative(D.loss, yb, p))
    z = np.random.default_rng(seed).normal(0, 1, (batch_size, latent_dim))
    fake = G.forward(z, training=True)
    ones = np.ones((batch_size, 1))
    p = D.forward(fake, training=False)
    D.y_true = ones
    g_loss = loss_value(G.loss, ones, p)
    g_grad = loss_derivative(G.loss, ones, p)
    d_grad = D.backward(g_grad, compute_only=True)
    G.backward(d_grad, gan=True)
    return d_total / n_critic, g_loss


# --------------------------------------------------------------------------------------------
# the five BASELINE.json configurations (SURVEY.md §8d "Synthetic inputs")
# --------------------------------------------------------------------------------------------


def config_specs(cfg: int):
    A = lambda fn: {"type": "act", "fn": fn}  # noqa: E731
    if cfg == 1:
        return [{"type": "input"}, {"type": "dense", "units": 128}, A("relu"), {"type": "dense", "units": 64}, A("relu"),
                {"type": "dense", "units": 10}, A("softmax")], (784,)
    if cfg == 2:
        return [{"type": "input"}, {"type": "conv2d", "filters": 32, "kernel": 3, "padding": "same"}, A("relu"),
                {"type": "bn"}, {"type": "maxpool", "pool": 2}, {"type": "flatten"},
                {"type": "dense", "units": 64}, A("relu"), {"type": "dense", "units": 10}, A("softmax")], (28, 28, 1)
    if cfg == 3:
        s = [{"type": "input"}]
        for f in (64, 128, 256):
            s += [{"type": "conv2d", "filters": f, "kernel": 3, "padding": "same"}, A("relu"), {"type": "bn"},
                  {"type": "maxpool", "pool": 2}]
        s += [{"type": "flatten"}, {"type": "dense", "units": 10}, A("softmax")]
        return s, (32, 32, 3)
    if cfg == 4:
        s = [{"type": "input"}]
        for _ in range(4):
            s += [{"type": "dense", "units": 4096}, A("relu")]
        s += [{"type": "dense", "units": 10}, A("softmax")]
        return s, (4096,)
    raise ValueError(cfg)


def gan_specs():
    A = lambda fn: {"type": "act", "fn": fn}  # noqa: E731
    G = [{"type": "input"}, {"type": "dense", "units": 7 * 7 * 128}, {"type": "reshape", "shape": (7, 7, 128)},
         {"type": "convT", "filters": 64, "kernel": 3, "strides": 2, "padding": "same"}, A("relu"),
         {"type": "convT", "filters": 32, "kernel": 3, "strides": 2, "padding": "same"}, A("relu"),
         {"type": "conv2d", "filters": 1, "kernel": 3, "padding": "same"}, A("sigmoid")]
    D = [{"type": "input"}, {"type": "conv2d", "filters": 32, "kernel": 3, "strides": 2, "padding": "same"}, A("relu"),
         {"type": "conv2d", "filters": 64, "kernel": 3, "strides": 2, "padding": "same"}, A("relu"), {"type": "flatten"},
         {"type": "dense", "units": 128}, A("relu"), {"type": "dense", "units": 1}, A("sigmoid")]
    return G, D


def synthetic_batch(cfg: int, batch: int, seed: int = 0):
    _, shp = config_specs(cfg)
    rng = np.random.default_rng(seed)
    x = rng.random((batch,) + shp).astype(np.float32)
    lab = rng.integers(0, 10, batch)
    y = np.zeros((batch, 10), dtype=np.float32)
    y[np.arange(batch), lab] = 1.0
    return x, y
