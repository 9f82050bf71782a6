"""TEST INFRASTRUCTURE ONLY — generates `tests/golden/*.npz` by EXECUTING THE UNMODIFIED UPSTREAM
REFERENCE (imported from /root/reference through `oracle/ref_loader.py`).

Run in the build container only:   python oracle/make_golden.py
The fixtures are what travels to the GPU box; nothing at test/bench time reads /root/reference.

What is recorded
  * op-level vectors: Conv2D / MaxPooling2D / BatchNormalization / Conv2DTranspose / Dense /
    activations / losses forward + backward on seeded random inputs, several geometries each;
  * `clip_gradients`, `Adam.update` (3 calls => the per-layer `t` quirk), `SGD.update`;
  * model-level traces of `Sequential.train_on_batch` for BASELINE configs 1, 2, 3 (reduced batch),
    a narrow variant of config 4, and `GAN.train_on_batch` on the notebook architecture: inputs,
    per-step loss and predictions, step-1 per-layer `input_error`, per-layer d_weights/d_bias,
    weights after every step, Adam m/v/t at the end, BN running stats.
Large tensors are stored as a fixed strided subsample plus their L2 norm and sum (see `pack`).
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from ref_loader import load_reference  # noqa: E402

load_reference()
from neuralnetlib.activations import LeakyReLU, Linear, ReLU, Sigmoid, Softmax, Tanh  # noqa: E402
from neuralnetlib.layers import (Activation, BatchNormalization, Conv2D, Conv2DTranspose, Dense, Flatten,  # noqa: E402
                                 Input, MaxPooling2D, Reshape)
from neuralnetlib.losses import BinaryCrossentropy, CategoricalCrossentropy, MeanSquaredError  # noqa: E402
from neuralnetlib.models import GAN, Sequential  # noqa: E402
from neuralnetlib.optimizers import SGD, Adam  # noqa: E402
from neuralnetlib.preprocessing import clip_gradients  # noqa: E402
from neuralnetlib.utils import shuffle  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
MAX_FULL = 12288
SUB = 4096


def pack(d: dict, name: str, arr):
    arr = np.array(arr, dtype=np.float64, copy=True)
    if arr.size <= MAX_FULL:
        d[name] = arr
    else:
        stride = arr.size // SUB
        d[name + "__sub"] = arr.ravel()[::stride].copy()
        d[name + "__stride"] = np.array(stride)
        d[name + "__shape"] = np.array(arr.shape)
        d[name + "__norm"] = np.array(np.sqrt(np.sum(arr * arr)))
        d[name + "__sum"] = np.array(arr.sum())


def save(name, d):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **d)
    print(f"{name}: {len(d)} arrays, {os.path.getsize(path) / 1024:.0f} KiB")


# ------------------------------------------------------------------------------------------------
def ops_golden():
    d = {}
    rng = np.random.default_rng(1234)

    # Dense
    x = rng.normal(size=(5, 7)); err = rng.normal(size=(5, 3))
    L = Dense(3, random_state=3)
    d["dense_x"], d["dense_err"] = x, err
    d["dense_y"] = L.forward_pass(x)
    d["dense_w"], d["dense_b"] = L.weights.copy(), L.bias.copy()
    d["dense_dx"] = L.backward_pass(err)
    d["dense_dw"], d["dense_db"] = L.d_weights, L.d_bias

    # activations
    xa = rng.normal(size=(6, 9)) * 3
    ea = rng.normal(size=(6, 9))
    d["act_x"], d["act_err"] = xa, ea
    for nm, fn in (("sigmoid", Sigmoid()), ("relu", ReLU()), ("tanh", Tanh()), ("linear", Linear()),
                   ("leakyrelu", LeakyReLU(0.2)), ("softmax", Softmax())):
        A = Activation(fn)
        d[f"act_{nm}_y"] = A.forward_pass(xa)
        if nm != "softmax":
            d[f"act_{nm}_dx"] = A.backward_pass(ea)

    # losses
    p = Softmax()(rng.normal(size=(8, 10)))
    y = np.eye(10)[rng.integers(0, 10, 8)]
    d["loss_p"], d["loss_y"] = p, y
    d["loss_cce"] = np.array(CategoricalCrossentropy()(y, p)); d["loss_cce_d"] = CategoricalCrossentropy().derivative(y, p)
    pb = Sigmoid()(rng.normal(size=(8, 1))); yb = rng.integers(0, 2, (8, 1)).astype(float)
    d["loss_pb"], d["loss_yb"] = pb, yb
    d["loss_bce"] = np.array(BinaryCrossentropy()(yb, pb)); d["loss_bce_d"] = BinaryCrossentropy().derivative(yb, pb)
    d["loss_mse"] = np.array(MeanSquaredError()(y, p)); d["loss_mse_d"] = MeanSquaredError().derivative(y, p)

    # clip
    g = rng.normal(size=(4, 5))
    d["clip_in"] = g
    d["clip_big"] = clip_gradients(g.copy())
    d["clip_small"] = clip_gradients(g.copy() * 1e-2)

    # Conv2D geometries: (H, W, C, F, k, s, padding)
    for gi, (H, W, C, F, k, s, pad) in enumerate([(6, 6, 3, 4, 3, 1, "same"), (7, 6, 2, 5, 3, 2, "same"),
                                                  (8, 8, 3, 2, 3, 2, "same"), (6, 7, 2, 3, 2, 1, "valid"),
                                                  (9, 9, 1, 4, (3, 2), (2, 1), "valid"), (5, 5, 4, 3, 1, 1, "same")]):
        x = rng.normal(size=(3, H, W, C))
        L = Conv2D(F, k, strides=s, padding=pad, random_state=5 + gi, bias_init="uniform")
        yv = L.forward_pass(x)
        err = rng.normal(size=yv.shape)
        dx = L.backward_pass(err)
        d[f"conv{gi}_cfg"] = np.array([H, W, C, F, L.kernel_size[0], L.kernel_size[1], L.strides[0], L.strides[1], pad == "same"])
        d[f"conv{gi}_x"], d[f"conv{gi}_w"], d[f"conv{gi}_b"], d[f"conv{gi}_y"] = x, L.weights, L.bias, yv
        d[f"conv{gi}_err"], d[f"conv{gi}_dx"], d[f"conv{gi}_dw"], d[f"conv{gi}_db"] = err, dx, L.d_weights, L.d_bias

    # MaxPool: with ties (quantised input) and overlapping windows
    for gi, (H, W, C, p, s, q) in enumerate([(6, 6, 3, 2, None, False), (6, 6, 2, 2, None, True), (7, 7, 2, 3, (2, 2), True),
                                             (8, 6, 2, (2, 3), None, False)]):
        x = rng.normal(size=(3, H, W, C))
        if q:
            x = np.round(x)  # many ties, incl. negative maxima and zeros
        L = MaxPooling2D(p, strides=s)
        yv = L.forward_pass(x)
        err = rng.normal(size=yv.shape)
        d[f"pool{gi}_cfg"] = np.array([H, W, C, L.pool_size[0], L.pool_size[1], L.strides[0], L.strides[1]])
        d[f"pool{gi}_x"], d[f"pool{gi}_y"], d[f"pool{gi}_err"], d[f"pool{gi}_dx"] = x, yv, err, L.backward_pass(err)
    x = np.zeros((1, 2, 2, 1)); L = MaxPooling2D(2); L.forward_pass(x)
    d["pool_allzero_dx"] = L.backward_pass(np.ones((1, 1, 1, 1)))

    # BatchNorm: two training steps (running stats), then inference; values beyond +-10 to hit the clip
    x1 = rng.normal(size=(8, 4, 4, 3)) * 6; x2 = rng.normal(size=(8, 4, 4, 3)) * 2
    L = BatchNormalization()
    d["bn_x1"], d["bn_x2"] = x1, x2
    d["bn_y1"] = L.forward_pass(x1, True)
    e1 = rng.normal(size=x1.shape); d["bn_e1"] = e1
    d["bn_dx1"] = L.backward_pass(e1); d["bn_dgamma1"], d["bn_dbeta1"] = L.d_gamma, L.d_beta
    d["bn_rm1"], d["bn_rv1"] = L.running_mean.copy(), L.running_var.copy()
    d["bn_y2"] = L.forward_pass(x2, True)
    d["bn_rm2"], d["bn_rv2"] = L.running_mean.copy(), L.running_var.copy()
    d["bn_yinf"] = L.forward_pass(x1, False)
    xb2 = rng.normal(size=(6, 5)); L2 = BatchNormalization(momentum=0.8, epsilon=1e-3)
    d["bn2d_x"] = xb2; d["bn2d_y"] = L2.forward_pass(xb2, True)
    e2 = rng.normal(size=xb2.shape); d["bn2d_e"] = e2; d["bn2d_dx"] = L2.backward_pass(e2)

    # Conv2DTranspose geometries
    for gi, (H, W, C, F, k, s, pad) in enumerate([(4, 4, 3, 2, 3, 2, "same"), (7, 7, 4, 3, 3, 2, "same"), (3, 4, 2, 3, 3, 2, "valid"),
                                                  (4, 4, 2, 2, 4, 2, "same"), (5, 5, 2, 3, 3, 1, "same"), (4, 3, 2, 2, 2, 2, "valid")]):
        x = rng.normal(size=(3, H, W, C))
        L = Conv2DTranspose(F, k, strides=s, padding=pad, random_state=9 + gi, bias_init="uniform")
        yv = L.forward_pass(x)
        err = rng.normal(size=yv.shape)
        dx = L.backward_pass(err)
        d[f"convT{gi}_cfg"] = np.array([H, W, C, F, L.kernel_size[0], L.kernel_size[1], L.strides[0], L.strides[1], pad == "same"])
        d[f"convT{gi}_x"], d[f"convT{gi}_w"], d[f"convT{gi}_b"], d[f"convT{gi}_y"] = x, L.weights, L.bias, yv
        d[f"convT{gi}_err"], d[f"convT{gi}_dx"], d[f"convT{gi}_dw"], d[f"convT{gi}_db"] = err, dx, L.d_weights, L.d_bias

    # Adam: three update() calls on two "layers" => t = 1, 2, 3 (Q3); SGD one call
    w0 = rng.normal(size=(4, 3)); b0 = rng.normal(size=(1, 3)); w1 = rng.normal(size=(3, 2)); b1 = rng.normal(size=(1, 2))
    gs = [rng.normal(size=s_) * sc for s_, sc in (((4, 3), 1.0), ((1, 3), 1e-9), ((3, 2), 1e-3), ((1, 2), 1.0), ((4, 3), 0.1), ((1, 3), 0.0))]
    for i, a in enumerate((w0, b0, w1, b1)):
        d[f"adam_p{i}_init"] = a.copy()
    for i, a in enumerate(gs):
        d[f"adam_g{i}"] = a
    opt = Adam()
    opt.update(5, w0, gs[0], b0, gs[1])
    opt.update(3, w1, gs[2], b1, gs[3])
    opt.update(5, w0, gs[4], b0, gs[5])
    for i, a in enumerate((w0, b0, w1, b1)):
        d[f"adam_p{i}_final"] = a.copy()
    d["adam_t"] = np.array(opt.t)
    d["adam_m_w5"], d["adam_v_w5"], d["adam_m_b5"], d["adam_v_b5"] = opt.m_w[5], opt.v_w[5], opt.m_b[5], opt.v_b[5]
    ws = rng.normal(size=(3, 3)); d["sgd_w_init"] = ws.copy(); gsg = rng.normal(size=(3, 3)); d["sgd_g"] = gsg
    SGD(0.05).update(0, ws, gsg); d["sgd_w_final"] = ws

    # shuffle permutation (utils.py:34-50)
    xs = np.arange(20).reshape(10, 2); ys = np.arange(10)
    a, b = shuffle(xs, ys, random_state=42)
    d["shuffle_x"], d["shuffle_y"] = a, b
    save("ops", d)


# ------------------------------------------------------------------------------------------------
def build_cfg(cfg: int, width: int | None = None):
    m = Sequential(random_state=42)
    if cfg == 1:
        m.add(Input(784)); m.add(Dense(128, activation="relu")); m.add(Dense(64, activation="relu")); m.add(Dense(10, activation="softmax"))
        shp = (784,)
    elif cfg == 2:
        m.add(Input((28, 28, 1))); m.add(Conv2D(32, 3, padding="same", activation="relu")); m.add(BatchNormalization())
        m.add(MaxPooling2D(2)); m.add(Flatten()); m.add(Dense(64, activation="relu")); m.add(Dense(10, activation="softmax"))
        shp = (28, 28, 1)
    elif cfg == 3:
        m.add(Input((32, 32, 3)))
        for f in (64, 128, 256):
            m.add(Conv2D(f, 3, padding="same", activation="relu")); m.add(BatchNormalization()); m.add(MaxPooling2D(2))
        m.add(Flatten()); m.add(Dense(10, activation="softmax"))
        shp = (32, 32, 3)
    elif cfg == 4:
        w = width or 4096
        m.add(Input(w))
        for _ in range(4):
            m.add(Dense(w, activation="relu"))
        m.add(Dense(10, activation="softmax"))
        shp = (w,)
    m.compile(loss_function="categorical_crossentropy", optimizer=Adam())
    return m, shp


def model_golden(name, cfg, batch, steps=3, width=None):
    m, shp = build_cfg(cfg, width)
    rng = np.random.default_rng(0)
    x = rng.random((batch,) + shp).astype(np.float32)
    lab = rng.integers(0, 10, batch)
    y = np.zeros((batch, 10), dtype=np.float32); y[np.arange(batch), lab] = 1.0
    d = {"x": x, "y": y, "batch": np.array(batch), "steps": np.array(steps)}
    if width:
        d["width"] = np.array(width)
    # step-1 backward trace: wrap every layer's backward_pass to record its returned input_error
    trace = []
    for li, L in enumerate(m.layers):
        orig = L.backward_pass

        def wrapped(err, _orig=orig, _li=li):
            out = _orig(err)
            trace.append((_li, np.array(out, dtype=np.float64)))
            return out
        L.backward_pass = wrapped
    for s in range(steps):
        trace.clear()
        loss = m.train_on_batch(x, y)
        d[f"s{s}_loss"] = np.array(loss)
        pack(d, f"s{s}_pred", m.predictions)
        for li, L in enumerate(m.layers):
            if hasattr(L, "weights") and L.weights is not None:
                pack(d, f"s{s}_L{li}_w", L.weights); pack(d, f"s{s}_L{li}_b", L.bias)
                pack(d, f"s{s}_L{li}_dw", L.d_weights); pack(d, f"s{s}_L{li}_db", L.d_bias)
            if isinstance(L, BatchNormalization):
                pack(d, f"s{s}_L{li}_rm", L.running_mean); pack(d, f"s{s}_L{li}_rv", L.running_var)
                pack(d, f"s{s}_L{li}_dgamma", L.d_gamma); pack(d, f"s{s}_L{li}_dbeta", L.d_beta)
        if s == 0:
            for li, e in trace:
                pack(d, f"s0_L{li}_inerr", e)
    opt = m.optimizer
    d["adam_t"] = np.array(opt.t)
    d["adam_keys"] = np.array(sorted(opt.m_w.keys()))
    for k in opt.m_w:
        pack(d, f"adam_m_w{k}", opt.m_w[k]); pack(d, f"adam_v_w{k}", opt.v_w[k])
        pack(d, f"adam_m_b{k}", opt.m_b[k]); pack(d, f"adam_v_b{k}", opt.v_b[k])
    # inference forward after training (BN running stats path)
    pack(d, "final_pred_eval", m.forward_pass(x, training=False))
    save(name, d)


def gan_golden(batch=4, steps=2):
    G = Sequential(random_state=7)
    G.add(Input(32)); G.add(Dense(7 * 7 * 128)); G.add(Reshape((7, 7, 128)))
    G.add(Conv2DTranspose(64, kernel_size=3, strides=2, padding="same", activation="relu"))
    G.add(Conv2DTranspose(32, kernel_size=3, strides=2, padding="same", activation="relu"))
    G.add(Conv2D(1, kernel_size=3, padding="same", activation="sigmoid"))
    D = Sequential(random_state=7)
    D.add(Input((28, 28, 1)))
    D.add(Conv2D(32, kernel_size=3, strides=2, padding="same", activation="relu"))
    D.add(Conv2D(64, kernel_size=3, strides=2, padding="same", activation="relu"))
    D.add(Flatten()); D.add(Dense(128, activation="relu")); D.add(Dense(1, activation="sigmoid"))
    gan = GAN(latent_dim=32, random_state=7)
    gan.compile(G, D, generator_optimizer="adam", discriminator_optimizer="adam", loss_function="bce")
    rng = np.random.default_rng(0)
    real = rng.random((batch * 2, 28, 28, 1)).astype(np.float32)
    d = {"real": real, "batch": np.array(batch), "steps": np.array(steps)}
    for s in range(steps):
        dl, gl = gan.train_on_batch(real, batch_size=batch, n_critic=1)
        d[f"s{s}_d_loss"], d[f"s{s}_g_loss"] = np.array(dl), np.array(gl)
        for nm, M in (("G", G), ("D", D)):
            for li, L in enumerate(M.layers):
                if hasattr(L, "weights") and L.weights is not None:
                    pack(d, f"s{s}_{nm}{li}_w", L.weights); pack(d, f"s{s}_{nm}{li}_b", L.bias)
                    pack(d, f"s{s}_{nm}{li}_dw", L.d_weights); pack(d, f"s{s}_{nm}{li}_db", L.d_bias)
        pack(d, f"s{s}_fake", G.predictions)
    d["G_t"], d["D_t"] = np.array(gan.generator_optimizer.t), np.array(gan.discriminator_optimizer.t)
    save("gan_b4", d)


def fit_golden():
    """Sequential.fit: 2 epochs, batch 16 over 40 samples (short last batch), accuracy metric, validation."""
    m, shp = build_cfg(1)
    rng = np.random.default_rng(0)
    x = rng.random((40, 784)).astype(np.float32)
    y = np.eye(10, dtype=np.float32)[rng.integers(0, 10, 40)]
    xv = rng.random((12, 784)).astype(np.float32); yv = np.eye(10, dtype=np.float32)[rng.integers(0, 10, 12)]
    h = m.fit(x, y, epochs=2, batch_size=16, verbose=False, metrics=["accuracy"], random_state=42, validation_data=(xv, yv))
    d = {"x": x, "y": y, "xv": xv, "yv": yv}
    for k, v in h.items():
        d["hist_" + k] = np.array(v, dtype=np.float64)
    for li, L in enumerate(m.layers):
        if hasattr(L, "weights") and L.weights is not None:
            pack(d, f"L{li}_w", L.weights); pack(d, f"L{li}_b", L.bias)
    d["adam_t"] = np.array(m.optimizer.t)
    pack(d, "pred", m.predict(xv))
    save("fit_cfg1", d)


if __name__ == "__main__":
    ops_golden()
    model_golden("cfg1_b32", 1, 32)
    model_golden("cfg2_b8", 2, 8)
    model_golden("cfg3_b4", 3, 4, steps=2)
    model_golden("cfg4w256_b16", 4, 16, width=256)
    gan_golden()
    fit_golden()
