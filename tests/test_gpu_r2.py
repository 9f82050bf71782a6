"""GPU parity of the round-2 additions against fixtures recorded from the unmodified reference (oracle/make_golden_r2.py):
the remaining optimizers (Momentum, RMSprop, AdaBelief, RAdam, Adam with clip_norm / clip_value), the remaining losses
(MeanAbsoluteError, Huber, SparseCategoricalCrossentropy) — op level through `Optimizer.update` / `LossFunction.__call__`, model
level through `Sequential.train_on_batch` — and one training step of the BASELINE configs AT THEIR FULL SIZE (config 3 at batch
1024, config 4 at batch 8192, the GAN at batch 64). Tolerances: norm-relative 1e-5 (fp32 against the reference's fp64)."""
import numpy as np
import pytest

import nnl_oracle as O
from helpers import check_packed, nrel

pytestmark = pytest.mark.gpu

OPT_NAMES = ["momentum", "rmsprop", "adabelief", "radam", "radam_default", "adam_clip"]


def _make_opt(name):
    from neuralnetlib_b200 import optimizers as Opt
    return {"momentum": lambda: Opt.Momentum(learning_rate=0.05, momentum=0.8),
            "rmsprop": lambda: Opt.RMSprop(learning_rate=0.01, rho=0.85, epsilon=1e-7),
            "adabelief": lambda: Opt.AdaBelief(learning_rate=0.002),
            "radam": lambda: Opt.RAdam(learning_rate=0.003, beta_2=0.9),
            "radam_default": lambda: Opt.RAdam(),
            "adam_clip": lambda: Opt.Adam(learning_rate=0.01, clip_norm=0.5, clip_value=0.02)}[name]()


@pytest.mark.parametrize("name", OPT_NAMES)
def test_optimizer_update_golden(golden, name):
    """`update()` mutates the caller's arrays in place (optimizers.py:10-12); four steps over two layers exercise the per-layer
    `t` of the Adam family, RAdam's rectified and un-rectified branches, and the optimizer-level clips."""
    g = golden("opt_loss")
    opt = _make_opt(name)
    params = {li: (g[f"{name}_L{li}_w0"].copy(), g[f"{name}_L{li}_b0"].copy()) for li in (3, 1)}
    for s in range(4):
        for li in (3, 1):
            w, b = params[li]
            opt.update(li, w, g[f"{name}_s{s}_L{li}_gw"], b, g[f"{name}_s{s}_L{li}_gb"])
            # teacher forcing is not needed: every rule is well conditioned on these inputs except RMSprop / Adam at |g| ~ eps
            assert nrel(w, g[f"{name}_s{s}_L{li}_w"]) < 2e-6, (name, s, li)
            assert nrel(b, g[f"{name}_s{s}_L{li}_b"]) < 2e-6, (name, s, li)
    if f"{name}_t" in g:
        assert opt.t == int(g[f"{name}_t"])
    cfg = opt.get_config()
    assert cfg["name"] == type(opt).__name__
    from neuralnetlib_b200.optimizers import Optimizer
    clone = Optimizer.from_config(cfg)
    assert type(clone) is type(opt) and clone.learning_rate == opt.learning_rate


def test_losses_golden(golden):
    from neuralnetlib_b200 import losses as Ls
    g = golden("opt_loss")
    p, y, yi = g["loss_p"], g["loss_y"], g["loss_yi"]
    for nm, lf, yy in (("mae", Ls.MeanAbsoluteError(), y), ("huber", Ls.Huber(delta=0.3), y),
                       ("scce", Ls.SparseCategoricalCrossentropy(), yi)):
        assert abs(lf(yy, p) - float(g[f"loss_{nm}_value"])) <= 1e-6 * abs(float(g[f"loss_{nm}_value"])), nm
        assert nrel(lf.derivative(yy, p), g[f"loss_{nm}_deriv"]) < 1e-6, nm
    assert isinstance(Ls.LossFunction.from_name("huber"), Ls.Huber)
    assert Ls.LossFunction.from_config({"name": "Huber", "delta": 0.3}).delta == 0.3


@pytest.mark.parametrize("name", OPT_NAMES + ["sgd_scce", "sgd_scce_tanh", "sgd_mae", "sgd_huber"])
def test_model_with_other_optimizers_and_losses(golden, name):
    from neuralnetlib_b200 import layers as L, losses as Ls, models as M, optimizers as Opt
    g = golden("opt_loss")
    last, loss, yy = {"sgd_scce": ("softmax", Ls.SparseCategoricalCrossentropy(), g["m_yi"]),
                      "sgd_scce_tanh": ("tanh", Ls.SparseCategoricalCrossentropy(), g["m_yi"]),
                      "sgd_mae": ("sigmoid", Ls.MeanAbsoluteError(), g["m_yr"]),
                      "sgd_huber": ("tanh", Ls.Huber(delta=0.3), g["m_yr"])}.get(name, ("softmax", "categorical_crossentropy", g["m_y1"]))
    m = M.Sequential(random_state=5)
    m.add(L.Input(12)); m.add(L.Dense(16, activation="relu")); m.add(L.Dense(5, activation=last))
    m.compile(loss_function=loss, optimizer=_make_opt(name) if name in OPT_NAMES else Opt.SGD(learning_rate=0.1))
    for s in range(2):
        lv = m.train_on_batch(g["m_x"], yy)
        ref = float(g[f"model_{name}_s{s}_loss"])
        assert abs(lv - ref) <= 1e-5 * max(1.0, abs(ref)), (name, s, lv, ref)
        for li, layer in enumerate(m.layers):
            if hasattr(layer, "weights"):
                assert nrel(layer.d_weights, g[f"model_{name}_s{s}_L{li}_dw"]) < 2e-5, (name, s, li)
                assert nrel(layer.weights, g[f"model_{name}_s{s}_L{li}_w"]) < 2e-5, (name, s, li)
                assert nrel(layer.bias, g[f"model_{name}_s{s}_L{li}_b"]) < 2e-5, (name, s, li)


# ---- BASELINE size ---------------------------------------------------------------------------------------------------------
def _counted_close(a, ref, stride, tol, what, frac=0.1, cap=1e-3, l2cap=2e-3):
    """Element-wise bound with a COUNTED allowance of outliers (discrete flips, see below): returns (number of elements beyond
    tol * max|ref|, worst norm-relative deviation); at most `frac` of the elements may exceed tol and none may exceed `cap`."""
    a = np.asarray(a, dtype=np.float64).ravel()[::stride][:ref.size]
    d = np.abs(a - ref.ravel())
    scale = max(np.abs(ref).max(), 1e-300)
    bad, worst = int(np.sum(d > tol * scale)), float(d.max() / scale)
    l2 = float(np.sqrt(np.sum(d * d)) / max(np.sqrt(np.sum(ref.ravel() ** 2)), 1e-300))
    ok = l2 <= l2cap and bad <= max(2, int(ref.size * frac)) and worst <= cap
    return ok, bad, d.size, float(f"{worst:.2e}"), float(f"{l2:.2e}")


@pytest.mark.parametrize("math", ["fp32", "tf32x3"])
@pytest.mark.parametrize("name,cfg,batch", [("cfg3_b1024", 3, 1024), ("cfg4_b8192", 4, 8192)])
def test_full_size_step_golden(golden, name, cfg, batch, math):
    """ONE train step at the size BASELINE.json quotes, against the reference's own trace. Loss / predictions 1e-5. Gradients:
    the LAST weighted layer (no discrete decision lies between it and the loss) at the strict norm-relative bound 2e-5; every
    other gradient at 2e-5 with a COUNTED number of outliers (fp32: <= 15 % of the elements, none beyond 5e-3 of the largest, L2-
    relative deviation <= 5e-4; the counts are printed). Two effects reach those tensors and only those: (1) discrete decisions:
    an fp32 max-pool window whose two largest BatchNorm outputs differ by less than fp32 resolution selects a different maximum
    than fp64, and a ReLU pre-activation within rounding of zero flips its mask (config 4 has 4 x 33 M of them); one flip moves
    one error element (3 windows of 1.6 M at config 2, see test_gpu_models.grads_close); (2) a weight / bias gradient of a convolution sums up to 1 M terms (1024 x 32 x 32
    pixels) per element, so fp32 accumulation alone leaves ~1e-5-level noise relative to the largest element. For scale: the
    reference's OWN algorithm evaluated in float32 NumPy deviates from its float64 evaluation by 2e-3 .. 1.5e-2 on these tensors
    with 6 .. 95 % of the elements beyond 2e-5 (scripts/fp32_emulation_cfg3.py, profiles/r2_fp32_emulation_cfg3_b1024.txt).
    Under tf32x3 the conv2 / conv3 contractions run on the tensor cores at ~1e-6 per contraction (ten times the FFMA kernels' error),
    which flips ten times as many near-tied pool maxima / ReLU masks: against the FFMA path on the same GPU the deviations are
    confined to a few filters of conv3's gradient (median 3e-7, 1.5 % of the elements beyond 2e-5, worst 7e-3) and spread below it
    (median 5e-5 .. 9e-5; scripts/diag_x3_cfg3.py). Those tensors are held to the north star's tensor-core bound in the L2 sense
    (||a - b|| <= 2e-3 ||b||) with every element within 1e-2 of the largest; everything above the last pool keeps the strict 2e-5."""
    import neuralnetlib_b200 as nn
    from test_gpu_models import build_cfg
    nn.set_math(math)
    g = golden(name)
    x, y = O.synthetic_batch(cfg, batch)
    assert abs(float(x.astype(np.float64).sum()) - float(g["x_sum"])) < 1e-6 * float(g["x_sum"])      # same inputs as the fixture
    m = build_cfg(cfg)
    loss = m.train_on_batch(x, y)
    assert abs(loss - float(g["s0_loss"])) <= 1e-5 * max(1.0, abs(float(g["s0_loss"]))), (loss, float(g["s0_loss"]))
    check_packed(g, "s0_pred", m.predictions, 2e-5, what=name + " ")
    weighted = [(li, L) for li, L in enumerate(m.layers) if hasattr(L, "weights")]
    # discrete decisions (max-pool selections, ReLU masks) sit between every weighted layer and the loss except the LAST one
    last_pool = weighted[-1][0] - 1
    report = {}
    # (fraction of elements allowed beyond 2e-5, cap on every element, cap on the L2-relative deviation)
    # config 4 (4 x 33 M ReLU decisions): a NumPy float32 evaluation of the reference's own algorithm already deviates by 1.7e-3
    # (L2) from its float64 evaluation at width 4096 (scripts/fp32_emulation_cfg3.py's method on config 4; the deviation appears
    # at ONE layer boundary and persists below it: ReLU-mask flips), so those tensors are bounded in the L2 sense only
    tc_bound = {(3, "fp32"): (0.15, 5e-3, 5e-4), (3, "tf32x3"): (1.0, 1e-2, 2e-3),
                (4, "fp32"): (1.0, 2e-2, 5e-3), (4, "tf32x3"): (1.0, 5e-2, 1e-2)}[(cfg, math)]
    strict = (0.0, 2e-5, 2e-5)
    for li, L in weighted:
        for nm, got in (("dw", L.d_weights), ("db", L.d_bias)):
            key = f"s0_L{li}_{nm}"
            bound = strict if li > last_pool else tc_bound
            if key in g:
                report[key] = _counted_close(got, g[key], 1, 2e-5, key, *bound)
            else:
                r = _counted_close(got, g[key + "__sub"], int(g[key + "__stride"]), 2e-5, key, *bound)
                nr = float(np.sqrt(np.sum(np.asarray(got, np.float64) ** 2)))   # and the tensor as a whole: its norm
                nrm = abs(nr - float(g[key + "__norm"])) / float(g[key + "__norm"])
                report[key] = (r[0] and nrm <= max(bound[2], 1e-4),) + r[1:] + (float(f"{nrm:.2e}"),)
    print(f"{name} [{math}]: (ok, elements beyond 2e-5, of, worst norm-relative, L2-relative[, whole-tensor norm error]): {report}")
    assert all(v[0] for v in report.values()), {k: v for k, v in report.items() if not v[0]}
    assert m.optimizer.t == int(g["adam_t"])


@pytest.mark.parametrize("math", ["fp32", "tf32x3"])
def test_gan_batch64_golden(golden, math):
    """Two GAN.train_on_batch steps of the notebook GAN at batch 64 (the round-1 trace was batch 4). fp32: every gradient within
    5e-5 (the discriminator's 128 -> 1 head: 2e-4). tf32x3 (Conv2DTranspose and the strided convolutions on the tensor cores): losses 2e-5, every gradient with no ReLU
    between it and the loss (the generator's output convolution, the whole discriminator) within 1e-4; the generator's layers BELOW
    its ReLUs are bounded at 5e-2 — a ~1e-6 forward difference flips a few of the 1.6 M ReLU masks and each flip moves one element
    of an error tensor whose norm is tiny at this point of training (the same effect as in test_full_size_step_golden)."""
    import neuralnetlib_b200 as nn
    from bench import build_gan
    nn.set_math(math)
    g = golden("gan_b64")
    gan = build_gan()
    real = g["real"].astype(np.float32)
    for s in range(int(g["steps"])):
        dl, gl = gan.train_on_batch(real, batch_size=int(g["batch"]), n_critic=1)
        assert abs(dl - float(g[f"s{s}_d_loss"])) <= 2e-5 * max(1.0, abs(float(g[f"s{s}_d_loss"]))), (s, dl)
        assert abs(gl - float(g[f"s{s}_g_loss"])) <= 2e-5 * max(1.0, abs(float(g[f"s{s}_g_loss"]))), (s, gl)
        if s == 0:
            for nm, M in (("G", gan.generator), ("D", gan.discriminator)):
                last = max(li for li, L in enumerate(M.layers) if hasattr(L, "weights"))
                for li, L in enumerate(M.layers):
                    if hasattr(L, "weights"):
                        below_relu = nm == "G" and li < last
                        # the discriminator's 128 -> 1 head gradient sums 64 near-cancelling terms per element: 9e-5 in fp32
                        tol = (2e-4 if (nm == "D" and li == last) else 5e-5) if math == "fp32" else (5e-2 if below_relu else 2e-4)
                        check_packed(g, f"s0_{nm}{li}_dw", L.d_weights, tol, what=f"{nm}{li} ")


# ---- streaming BatchNorm / pool kernels (large batches) -------------------------------------------------------------------
@pytest.mark.parametrize("Bn,H,W,C", [(1024, 8, 8, 32), (600, 4, 6, 64), (129, 2, 2, 32), (2048, 4, 4, 32)])
def test_streaming_bn_pool_kernels(Bn, H, W, C):
    """The streaming forms (csrc/bn_pool.cu: one-sweep shifted statistics merged in fp64; partial / reduce / apply backward)
    against the unfused oracle chain incl. the three clip points between MaxPool-bwd, BatchNorm-bwd and ReLU'; ragged last batch
    chunk, 4-way pool ties, and the guard bands around every output."""
    import torch
    from neuralnetlib_b200 import _backend as B, _ops as ops
    from test_gpu_ops import Guarded

    import types
    G = Guarded(types.SimpleNamespace(empty=B.empty, B=B))
    rng = np.random.default_rng(Bn + H)
    x = np.maximum(rng.normal(size=(Bn, H, W, C)) * 0.05 - 0.01, 0).astype(np.float32)
    x[:, :2, :2, 0] = 0.0                       # an all-dead window: 4-way ties in the pool (Q6)
    x[:, :, :, 1] += 3.0                        # a channel with mean >> std: the one-pass variance must not cancel
    P = H * W * C
    gamma = (1 + 0.1 * rng.normal(size=P)).astype(np.float32); beta = (0.1 * rng.normal(size=P)).astype(np.float32)
    beta.reshape(H, W, C)[:2, :2, 0] = 0.25; gamma.reshape(H, W, C)[:2, :2, 0] = 1.0
    dev = lambda a: B.to_device(np.asarray(a, dtype=np.float32))   # noqa: E731
    host = lambda t: (B.synchronize(), t.detach().cpu().numpy().astype(np.float64))[1]   # noqa: E731
    xd, gd, bd = dev(x), dev(gamma), dev(beta)
    rm, rv = B.zeros((P,)), B.to_device(np.ones(P))
    sm, si = G((P,)), G((P,))
    stats = B.zeros((2 * P,), dtype=torch.float64)
    ws = G((max(ops.bn_stats_stream_ws_bytes(Bn, P), ops.pool_bn_bwd_stream_ws_bytes(Bn, H, W, C)) // 4,))
    ops.bn_stats_stream(xd, stats, ws)
    x64 = x.astype(np.float64)
    xc = np.clip(x64, -10, 10).reshape(Bn, P)
    st = host(stats)
    assert np.abs(st[:P] - xc.sum(0)).max() <= 1e-6 * np.abs(xc.sum(0)).max()
    var_ref = xc.var(0)
    var_got = st[P:] / Bn - (st[:P] / Bn) ** 2
    assert np.abs(var_got - var_ref).max() <= 1e-5 * var_ref.max() + 1e-12, np.abs(var_got - var_ref).max() / var_ref.max()
    yp = G((Bn, H // 2, W // 2, C))
    ops.bn_pool_fwd_apply(xd, gd, bd, rm, rv, sm, si, stats, yp, Bn, 0.9, 1e-5)
    shp = (H, W, C)
    g64, b64 = gamma.astype(np.float64).reshape(shp), beta.astype(np.float64).reshape(shp)
    ybn, rrm, rrv, cache = O.bn_fwd(x64, g64, b64, np.zeros(shp), np.ones(shp), 0.9, 1e-5, True)
    ryp = O.maxpool_fwd(ybn, (2, 2), (2, 2))
    assert nrel(host(yp), ryp) < 2e-5
    assert nrel(host(rm).reshape(shp), rrm) < 2e-5 and nrel(host(rv).reshape(shp), rrv) < 2e-5
    e = rng.normal(size=ryp.shape).astype(np.float32)
    ed = dev(e); ss = B.zeros((16,), dtype=torch.float64); ops.sumsq(ed, ss, 0)
    r0, sums = G(x.shape), G((2 * P,))
    ops.pool_bn_bwd_stream_partial(xd, ed, ss, (0,), gd, bd, sm, si, sums, ws, 1)
    ops.pool_bn_act_bwd_stream_apply(xd, ed, ss, (0,), gd, bd, sm, si, r0, sums, Bn, 1, 0.0, 2, 3)
    G.check()
    U = O.maxpool_bwd(ybn, O.clip_l2(e.astype(np.float64)), (2, 2), (2, 2))
    D, rdg, rdb = O.bn_bwd(U, g64, cache, 1e-5)
    R = D * (x64 > 0)
    h = host(ss)
    for slot, ref in ((1, U), (2, D), (3, R)):
        assert abs(h[slot] - np.sum(ref ** 2)) < 1e-4 * np.sum(ref ** 2), slot
    assert nrel(host(r0), R) < 2e-5
    hs = host(sums)
    assert nrel(hs[:P].reshape(shp), rdb) < 2e-5 and nrel(hs[P:].reshape(shp), rdg) < 2e-5


# ---- Conv2DTranspose on the tensor cores -------------------------------------------------------------------------------------
@pytest.mark.parametrize("math,tol", [(2, 1e-5), (1, 2e-3)])
@pytest.mark.parametrize("N,H,W,C,F,k,s,pad", [(64, 7, 7, 128, 64, 3, 2, "same"), (64, 14, 14, 64, 32, 3, 2, "same"),
                                                 (37, 5, 6, 32, 16, 3, 2, "valid"), (16, 6, 6, 64, 32, 4, 2, "same"),
                                                 (16, 8, 8, 32, 32, 3, 1, "same")])
def test_convt_tensor_cores(N, H, W, C, F, k, s, pad, math, tol):
    """Conv2DTranspose as tcgen05 contractions on column buffers (csrc/gemm_simt.cu: Ycol = x . W^T + overlap-add forward; window
    gather + Ecol . W and Ecol^T . x backward) against the oracle: strides 1 / 2, 'same' / 'valid', 3x3 / 4x4 kernels, the Q12 edge
    skip, bias + ReLU epilogue, pending clip + ReLU' mask + both sum-of-squares slots, guard bands around every output."""
    import torch
    import types
    from neuralnetlib_b200 import _backend as B, _ops as ops
    from test_gpu_ops import Guarded
    G = Guarded(types.SimpleNamespace(empty=B.empty, B=B))
    rng = np.random.default_rng(H + C + k)
    x = np.maximum(rng.normal(size=(N, H, W, C)), 0).astype(np.float32)
    w = (rng.normal(size=(k, k, F, C)) / np.sqrt(k * k * C)).astype(np.float32)
    b = rng.normal(size=(F,)).astype(np.float32)
    oh, ow, pt, _pb, pl, _pr = O.convT_geometry(H, W, k, k, s, s, pad)
    geom = B.ConvTGeom(N, H, W, C, F, k, k, s, s, pt, pl, oh, ow)
    err = rng.normal(size=(N, oh, ow, F)).astype(np.float32)
    dev = lambda a: B.to_device(np.asarray(a, dtype=np.float32))   # noqa: E731
    host = lambda t: (B.synchronize(), t.detach().cpu().numpy().astype(np.float64))[1]   # noqa: E731
    xd, wd, bd, ed = dev(x), dev(w), dev(b), dev(err)
    ws = G((max(ops.convt_ws_bytes(geom, math) // 4, 1),))
    y = G((N, oh, ow, F))
    ops.convt_fwd(math, geom, xd, wd, bd, y, 1, 0.0, ws)
    ref = np.maximum(O.convT_fwd(x.astype(np.float64), w.astype(np.float64), b.astype(np.float64), (s, s), pad), 0)
    assert nrel(host(y), ref) < tol
    ss = B.zeros((16,), dtype=torch.float64); ops.sumsq(ed, ss, 0)
    dx, dw, db = G(x.shape), G(w.shape), G((F,))
    ops.convt_bwd(math, geom, xd, wd, ed, ss, (0,), dx, dw, db, 1, 0.0, 1, 2, ws)
    G.check()
    rdx, rdw, rdb = O.convT_bwd(x.astype(np.float64), w.astype(np.float64), O.clip_l2(err.astype(np.float64)), (s, s))
    assert nrel(host(dx), rdx * (x > 0)) < tol
    assert nrel(host(dw), rdw) < tol
    assert nrel(host(db), rdb.reshape(-1)) < 1e-5
    h = host(ss)
    assert abs(h[1] - np.sum(rdx ** 2)) < max(10 * tol, 1e-4) * np.sum(rdx ** 2)


# ---- strided / 'valid' / small-F convolutions in the tensor-core math modes: im2col + the Dense kernels -------------------------
@pytest.mark.parametrize("math,tol", [(2, 1e-5), (1, 2e-3)])
@pytest.mark.parametrize("N,H,W,C,F,k,s,pad", [(64, 28, 28, 32, 1, 3, 1, "same"), (64, 13, 13, 32, 64, 3, 2, "same"),
                                                 (16, 12, 10, 8, 16, 5, 1, "valid"), (37, 9, 9, 64, 48, 3, 2, "valid"),
                                                 (8, 28, 28, 32, 64, 3, 2, "same")])
def test_conv2d_column_path(N, H, W, C, F, k, s, pad, math, tol):
    """Geometries outside the TMA implicit-GEMM path (non-power-of-two images, stride 2 incl. the Q11 'same' padding, 'valid',
    F = 1) run as im2col + Dense contractions in the tensor-core math modes; forward (bias + ReLU), data gradient (pending clip,
    ReLU' mask of the layer below, both sum-of-squares slots), weight / bias gradients in the reference's (c,ky,kx) order (Q1)."""
    import torch
    import types
    from neuralnetlib_b200 import _backend as B, _ops as ops
    from test_gpu_ops import Guarded
    G = Guarded(types.SimpleNamespace(empty=B.empty, B=B))
    rng = np.random.default_rng(H * 31 + C + F)
    x = np.maximum(rng.normal(size=(N, H, W, C)), 0).astype(np.float32)
    w = (rng.normal(size=(k, k, C, F)) / np.sqrt(k * k * C)).astype(np.float32)
    b = rng.normal(size=(F,)).astype(np.float32)
    oh, ow, ph, pw = O.conv2d_geometry(H, W, k, k, s, s, pad)
    geom = B.ConvGeom(N, H, W, C, F, k, k, s, s, ph, pw, oh, ow)
    err = rng.normal(size=(N, oh, ow, F)).astype(np.float32)
    dev = lambda a: B.to_device(np.asarray(a, dtype=np.float32))   # noqa: E731
    host = lambda t: (B.synchronize(), t.detach().cpu().numpy().astype(np.float64))[1]   # noqa: E731
    xd, wd, bd, ed = dev(x), dev(w), dev(b), dev(err)
    ws = G((max(ops.conv2d_ws_bytes(geom, math) // 4, 1),))
    y = G((N, oh, ow, F))
    ops.conv2d_fwd(math, geom, xd, wd, bd, y, 1, 0.0, ws)
    ref = np.maximum(O.conv2d_fwd(x.astype(np.float64), w.astype(np.float64), b.astype(np.float64), (s, s), pad), 0)
    assert nrel(host(y), ref) < tol
    ss = B.zeros((16,), dtype=torch.float64); ops.sumsq(ed, ss, 0)
    dx, dw, db = G(x.shape), G(w.shape), G((F,))
    ops.conv2d_bwd(math, geom, xd, wd, ed, ss, (0,), dx, dw, db, 1, 0.0, 1, 2, ws)
    G.check()
    rdx, rdw, rdb = O.conv2d_bwd(x.astype(np.float64), w.astype(np.float64), O.clip_l2(err.astype(np.float64)), (s, s), pad)
    assert nrel(host(dx), rdx * (x > 0)) < tol
    assert nrel(host(dw), rdw) < tol
    assert nrel(host(db), rdb) < 1e-5
    h = host(ss)
    assert abs(h[1] - np.sum(rdx ** 2)) < max(10 * tol, 1e-4) * np.sum(rdx ** 2)
    assert abs(h[2] - np.sum((rdx * (x > 0)) ** 2)) < max(10 * tol, 1e-4) * np.sum((rdx * (x > 0)) ** 2)


def test_evaluate_with_enable_padding_counts_real_rows_only():
    """evaluate() under enable_padding (models.py:159-181, 565-601): a batch that is not a multiple of padding_size runs on
    zero-padded rows, but loss and predictions cover the real rows only (round-1 advisor finding: it raised)."""
    from neuralnetlib_b200 import layers as L, models as M, optimizers as Opt
    rng = np.random.default_rng(2)
    x = rng.normal(size=(13, 9)).astype(np.float32)
    y = np.eye(4, dtype=np.float32)[rng.integers(0, 4, 13)]
    out = []
    for pad in (True, False):
        m = M.Sequential(random_state=8, enable_padding=pad, padding_size=8)
        m.add(L.Input(9)); m.add(L.Dense(6, activation="tanh")); m.add(L.Dense(4, activation="softmax"))
        m.compile(loss_function="categorical_crossentropy", optimizer=Opt.SGD())
        out.append(m.evaluate(x, y, batch_size=13))
    assert abs(out[0][0] - out[1][0]) <= 1e-6 * max(1.0, abs(out[1][0]))
    assert out[0][1].shape == out[1][1].shape == (13, 4)
    assert nrel(out[0][1], out[1][1]) < 1e-6
